"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libqref.so):

    python tests/golden/make_golden.py

Writes
  tests/golden/slices.npz      full slices (every cell) computed by the reference's own
                               *_slice_compute_richardson / *_slice_compute entry points
  tests/golden/kat/*.txt       a sample of the reference's known-answer vector files
                               (res/test-vectors), copied record-for-record
  tests/golden/mathematica_totals.json
                               the Mathematica NIntegrate slice totals quoted in the
                               reference's src/test/test_linear_distribution.cpp and
                               src/test/test_diagonal_distribution.cpp (parsed, not copied code)
Long doubles are stored as (hi, lo) float64 pairs.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF = "/root/reference"


def split(x):
    x = np.asarray(x, dtype=np.longdouble)
    hi = x.astype(np.float64)
    lo = (x - hi.astype(np.longdouble)).astype(np.float64)
    return hi, lo


def synthetic_d_r(m, seed):
    import random
    rnd = random.Random(seed)
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    return d, r


def main():
    ref.build()
    out = {}
    meta = []

    def add(name, sl, **kw):
        hi, lo = split(sl.cells)
        out[name + "/cells_hi"] = hi
        out[name + "/cells_lo"] = lo
        tph, tpl = split([sl.total_probability])
        teh, tel = split([sl.total_error])
        # total_error underflows float64: keep mantissa/exponent
        te = np.longdouble(sl.total_error)
        mant, exp = np.frexp(te)
        meta.append(dict(name=name, tp_hi=float(tph[0]), tp_lo=float(tpl[0]),
                         te_mant=float(mant), te_exp=int(exp), flags=int(sl.flags), **kw))
        print(name, float(sl.total_probability), flush=True)

    # ---- two-dimensional -----------------------------------------------------
    cases2d = [
        # (tag, m, s, (d, r) source, D, richardson, method, coords)
        ("c2", 2048, 1, "syn", 32, 1, 0, [(2048, 2048), (-2049, 2048), (2047, 2048), (2058, 2058),
                                           (2030, 2040), (2018, 2018), (-2058, 2050), (2052, 2051)]),
        ("c2q", 2048, 1, "syn", 32, 1, 2, [(2048, 2048), (-2050, 2049)]),
        ("c2s", 2048, 1, "det", 16, 0, 0, [(2048, 2048), (2040, 2041)]),
        ("c1", 128, 2, "det", 32, 1, 0, [(130, 129), (128, 128), (-127, 126), (138, 138), (100, 101)]),
        ("c1q", 128, 2, "det", 32, 1, 2, [(128, 128)]),
        ("c4", 3072, 4, "det", 32, 1, 0, [(3075, 3074), (-3060, 3082), (3072, 3072)]),
        ("odd", 256, 3, "det", 12, 1, 0, [(256, 257), (-250, 255)]),
        ("c2d64", 2048, 1, "syn", 64, 1, 0, [(2049, 2049)]),
        # sigma-optimal method (-sigma-optimal; docs/pages/info-distribution.md uses it at s = 30)
        ("c1o", 128, 2, "det", 32, 1, 1, [(130, 129), (128, 128), (-127, 126), (136, 135)]),
        ("s30o", 2048, 30, "det", 32, 1, 1, [(2049, 2047), (-2048, 2048), (2040, 2044)]),
        ("c2o", 2048, 1, "syn", 8, 1, 1, [(2048, 2048), (-2050, 2049)]),
        ("c1os", 128, 2, "det", 16, 0, 1, [(129, 128)]),
    ]
    for tag, m, s, src, D, rich, method, coords in cases2d:
        d, r = synthetic_d_r(m, 20482048) if src == "syn" else ref.deterministic_d_r(m)
        P = ref.RefParameters(m, s, d, r)
        for (a, b) in coords:
            sl = ref.distribution_slice_compute(P, D, a, b, method=method, richardson=bool(rich))
            add(f"2d/{tag}/{a}_{b}", sl, kind="2d", m=m, s=s, l=P.l, d=str(d), r=str(r), D=D,
                richardson=rich, method=method, a_d=a, a_r=b)

    # ---- linear ------------------------------------------------------------------
    cases_lin = [
        ("c1", 128, 2, "det", 256, [128, -130, 100, 138, 157, 98]),
        ("c3", 1023, 8, "det", 128, [1023, -1030, 1000, 1033]),
        ("m2048", 2048, 1, "syn", 64, [2048, -2040]),
    ]
    for tag, m, s, src, D, coords in cases_lin:
        d, r = synthetic_d_r(m, 20482048) if src == "syn" else ref.deterministic_d_r(m)
        P = ref.RefParameters(m, s, d, r)
        for target in (0, 1):
            for a in coords:
                if tag == "m2048" and target == 0 and a != 2048:
                    continue  # 6144-bit MPFR: 6.4 ms per evaluation
                sl = ref.linear_distribution_slice_compute(P, D, a, target)
                add(f"lin/{tag}/t{target}/{a}", sl, kind="lin", m=m, s=s, l=P.l, d=str(d), r=str(r),
                    D=D, richardson=1, target=target, a=a)
    # single pass
    d, r = ref.deterministic_d_r(128)
    P = ref.RefParameters(128, 2, d, r)
    for target in (0, 1):
        sl = ref.linear_distribution_slice_compute(P, 64, 127, target, richardson=False)
        add(f"lin/single/t{target}/127", sl, kind="lin", m=128, s=2, l=P.l, d=str(d), r=str(r), D=64,
            richardson=0, target=target, a=127)

    # ---- diagonal ----------------------------------------------------------------
    cases_diag = [
        ("m128", 128, 5, 1, "det", 256, [(128, 0), (-130, 0), (126, -1), (131, 2), (125, 25), (100, -25)]),
        ("m2048", 2048, 12, 1, "syn", 64, [(2048, 0), (-2050, 1), (2058, -2), (2070, 0)]),
        ("m2048s0", 2048, 0, 1, "syn", 64, [(2040, 0), (2040, 1)]),
    ]
    for tag, m, sigma, s, src, D, coords in cases_diag:
        d, r = synthetic_d_r(m, 20482048) if src == "syn" else ref.deterministic_d_r(m)
        P = ref.RefDiagonalParameters(m, sigma, s, d, r, eta_bound=25)
        for (a, eta) in coords:
            sl = ref.diagonal_distribution_slice_compute(P, D, a, eta)
            add(f"diag/{tag}/{a}_{eta}", sl, kind="diag", m=m, s=s, sigma=sigma, l=int(np.ceil(m / s)),
                d=str(d), r=str(r), D=D, richardson=1, a=a, eta=eta)

    np.savez_compressed(os.path.join(HERE, "slices.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "slices_meta.json"), "w"), indent=0)

    # ---- KAT sample ----------------------------------------------------------------
    kat_dir = os.path.join(HERE, "kat")
    os.makedirs(kat_dir, exist_ok=True)
    tv = os.path.join(REF, "res", "test-vectors")

    def sample(fname, rec_lines, keep):
        lines = open(os.path.join(tv, fname)).read().split("\n")
        recs = [lines[i:i + rec_lines] for i in range(0, len(lines) - rec_lines + 1, rec_lines)]
        idx = sorted(set(np.linspace(0, len(recs) - 1, keep).astype(int).tolist()))
        with open(os.path.join(kat_dir, fname), "w") as f:
            for i in idx:
                f.write("\n".join(recs[i]) + "\n")
        return len(idx)

    for m, s in [(128, 1), (128, 2), (256, 4), (512, 8), (1024, 10), (2048, 1), (2048, 30), (4096, 50), (8192, 80)]:
        sample(f"probabilities-det-m-{m}-s-{s}.txt", 4, 64)
    for m, s in [(128, 1), (128, 2), (512, 8), (2048, 1), (2048, 30), (8192, 80)]:
        for t in ("d", "r"):
            sample(f"linear-probabilities-det-{t}-m-{m}-s-{s}.txt", 2, 61)
    for m, sigma, s in [(128, 0, 1), (128, 5, 10), (512, 3, 8), (2048, 0, 1), (2048, 4, 50)]:
        fname = f"diagonal-probabilities-f-eta-det-m-{m}-sigma-{sigma}-s-{s}.txt"
        lines = open(os.path.join(tv, fname)).read().split("\n")
        keep = 4 if m < 512 else 1   # records of 52 lines (2400-digit values at m = 2048)
        with open(os.path.join(kat_dir, fname), "w") as f:
            f.write("\n".join(lines[:52 * keep]) + "\n")

    # ---- Mathematica totals quoted in the reference's distribution tests ------------
    # Parsed from the numeric literals of the `expected_*probabilities` arrays
    # (src/test/test_linear_distribution.cpp:93-131, 353-388;
    #  src/test/test_diagonal_distribution.cpp:87-389). m = 128, s = 1, t = 30,
    # deterministic d, r; dimension 2048 with Richardson; tolerance 1e-6.
    totals = {}
    for which in ("linear", "diagonal"):
        src_txt = open(os.path.join(REF, "src", "test", f"test_{which}_distribution.cpp")).read()
        for mm in re.finditer(r"const long double (expected_\w+)\[[^\]]*\]\s*=\s*\{(.*?)\};", src_txt, re.S):
            body = re.sub(r"/\*.*?\*/", "", mm.group(2), flags=re.S)
            vals = re.findall(r"[0-9]\.[0-9]{40,}(?:e-?[0-9]+)?", body)
            totals.setdefault(which, []).append(dict(array=mm.group(1), values=vals))
    json.dump(totals, open(os.path.join(HERE, "mathematica_totals.json"), "w"), indent=0)
    print({k: [(a["array"], len(a["values"])) for a in v] for k, v in totals.items()})


if __name__ == "__main__":
    main()
