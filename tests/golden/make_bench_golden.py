"""Golden slices of the BENCH configuration itself, from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libqref.so):

    python tests/golden/make_bench_golden.py

bench.py's workload "T2D" (SURVEY.md section 8(d)): m = 2048, s = 1, synthetic d, r from
random.seed(20482048), dimension 128 with the Richardson pass at 256. The committed
tests/golden/slices.npz holds this configuration at D = 32 / 64 only; this file adds, at the
dimensions generate_distribution really uses (src/main_generate_distribution.cpp:1222-1343),

  * 13 slices at D = 128 spanning the fused kernel's three slice classes (main path, |u| < 1/16
    polynomial, |u| < 2^-9 polynomial), both signs of alpha_d (the same-sign and the
    opposite-sign tile variants) and the corners of the coordinate range,
  * one slice upgraded to `required_dimension` 512 (D = 256, stored as it is),
  * one slice upgraded to `required_dimension` 1024 (D = 512) scaled back to
    MAX_SLICE_DIMENSION = 256 with the reference's distribution_slice_copy_scale
    (src/distribution_slice.cpp:230-264),
  * one slice at D = 128 with the sigma-optimal method (-sigma-optimal),

each computed by the reference's own distribution_slice_compute_richardson (192-bit MPFR), one
slice per worker process. Output: tests/golden/bench_slices.npz (long doubles as (hi, lo) float64
pairs) + bench_slices_meta.json. About 4 minutes on 8 cores.
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

M, S, SEED = 2048, 1, 20482048

CASES = [
    # (tag, D, method, scale_to, a_d, a_r)
    ("c0", 128, 0, 0, 2048, 2048), ("c0", 128, 0, 0, -2049, 2048), ("c0", 128, 0, 0, 2058, 2058),
    ("c0", 128, 0, 0, -2058, 2050), ("c0", 128, 0, 0, 2047, 2052), ("c0", 128, 0, 0, -2045, 2046),
    ("c0", 128, 0, 0, 2052, 2038),
    ("c1", 128, 0, 0, 2040, 2041), ("c1", 128, 0, 0, -2041, 2039), ("c1", 128, 0, 0, 2042, 2030),
    ("c2", 128, 0, 0, 2030, 2033), ("c2", 128, 0, 0, -2025, 2036), ("c2", 128, 0, 0, 2018, 2018),
    ("up512", 256, 0, 0, 2049, 2048),
    ("up1024", 512, 0, 256, 2051, 2051),
    ("so", 128, 1, 0, 2048, 2048),
]


def synthetic_d_r(m, seed):
    import random
    rnd = random.Random(seed)
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    return d, r


def split(x):
    x = np.asarray(x, dtype=np.longdouble)
    hi = x.astype(np.float64)
    lo = (x - hi.astype(np.longdouble)).astype(np.float64)
    return hi, lo


def work(case):
    tag, D, method, scale_to, a_d, a_r = case
    d, r = synthetic_d_r(M, SEED)
    P = ref.RefParameters(M, S, d, r)
    sl = ref.distribution_slice_compute(P, D, a_d, a_r, method=method, richardson=True)
    cells, flags = sl.cells, sl.flags
    if scale_to:
        cells, flags = ref.distribution_slice_copy_scale(sl.cells, D, sl.flags, scale_to)
    return case, np.asarray(cells, dtype=np.longdouble), sl.total_probability, sl.total_error, int(flags)


def main():
    ref.build()
    out, meta = {}, []
    # the most expensive first so that the pool stays busy
    order = sorted(CASES, key=lambda c: -c[1] * (3 if c[2] == 1 else 1))
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        for case, cells, tp, te, flags in pool.imap_unordered(work, order):
            tag, D, method, scale_to, a_d, a_r = case
            name = f"{tag}/{a_d}_{a_r}"
            hi, lo = split(cells)
            out[name + "/cells_hi"], out[name + "/cells_lo"] = hi, lo
            tph, tpl = split([tp])
            mant, exp = np.frexp(np.longdouble(te))
            meta.append(dict(name=name, tag=tag, D=D, method=method, scale_to=scale_to, a_d=a_d, a_r=a_r,
                             tp_hi=float(tph[0]), tp_lo=float(tpl[0]), te_mant=float(mant), te_exp=int(exp),
                             flags=flags))
            print(name, float(tp), flush=True)
    meta.sort(key=lambda q: [c[:1] + c[4:] for c in CASES].index((q["tag"], q["a_d"], q["a_r"])))
    np.savez_compressed(os.path.join(HERE, "bench_slices.npz"), **out)
    json.dump(dict(m=M, s=S, seed=SEED, slices=meta), open(os.path.join(HERE, "bench_slices_meta.json"), "w"),
              indent=0)


if __name__ == "__main__":
    main()
