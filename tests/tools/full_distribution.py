"""Generate a FULL m = 2048 two-dimensional distribution with the reference's own generator
executable + the drop-in (BASELINE.json configs[1]) and check a random sample of its slices
against the reference itself (oracle/_ref) on the host cores.

    python tests/tools/full_distribution.py [--clients 1] [--dim 256] [--sample 32]

Writes gpurun_out/full_distribution_report.json.
"""
import argparse
import json
import multiprocessing as mp
import os
import random
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from integration import distfile  # noqa: E402

B = os.path.join(ROOT, "integration", "_build")


def _ref_slice(job):
    """What a reference client sends for one coordinate pair: fixed dimension, or the
    dimension heuristic of main_client (src/main_generate_distribution.cpp:1222-1342)."""
    from oracle import ref
    m, s, d, r, D, a_d, a_r = job
    P = ref.RefParameters(m, s, d, r)
    if D:
        sl = ref.distribution_slice_compute(P, D, a_d, a_r)
        return (a_d, a_r, D, np.asarray(sl.cells, dtype=np.longdouble), sl.total_error, sl.flags)
    max_alpha = max(abs(a_d), abs(a_r))
    required = 256
    if abs(a_d - a_r) <= 1 and (a_d > 0) == (a_r > 0):
        if max_alpha >= m + 3:
            required = 1024
        elif max_alpha >= m:
            required = 512
    sl = ref.distribution_slice_compute(P, required // 2, a_d, a_r)
    updated = False
    if sl.total_probability >= 1e-7 and max_alpha >= m and required < 512:
        required, updated = 512, True
    if sl.total_probability >= 1e-10 and max_alpha >= m + 10 and required < 1024:
        required, updated = 1024, True
    if updated:
        sl = ref.distribution_slice_compute(P, required // 2, a_d, a_r)
    D, cells, flags = required // 2, np.asarray(sl.cells, dtype=np.longdouble), sl.flags
    if D > 256:   # MAX_SLICE_DIMENSION
        cells, flags = ref.distribution_slice_copy_scale(cells, D, flags, 256)
        D = 256
    return (a_d, a_r, D, cells, sl.total_error, flags)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=1)
    ap.add_argument("--dim", type=int, default=256,
                    help="-dim value; 0 = the generator's default dimension heuristic")
    ap.add_argument("--sample", type=int, default=32)
    ap.add_argument("--m", type=int, default=2048)
    ap.add_argument("--s", type=int, default=1)
    args = ap.parse_args()
    m = args.m
    rnd = random.Random(20482048)
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    t = tempfile.mkdtemp()
    os.makedirs(os.path.join(t, "distributions"))
    cmd = [os.path.join(B, "minimpirun"), "-np", str(args.clients + 1),
           os.path.join(B, "gpu", "generate_distribution"), "-exp", str(d), str(r),
           *(["-dim", str(args.dim)] if args.dim else []), str(m), str(args.s)]
    env = dict(os.environ)
    t0 = time.time()
    p = subprocess.run(cmd, cwd=t, env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit(1)
    lines = p.stdout.splitlines()
    t_compute = None
    files = sorted(os.listdir(os.path.join(t, "distributions")))
    main_file = [f for f in files if f.startswith("distribution-") and f.endswith(".txt")][0]
    size = os.path.getsize(os.path.join(t, "distributions", main_file))
    t1 = time.time()
    dist = distfile.read(os.path.join(t, "distributions", main_file), "2d")
    t_parse = time.time() - t1
    keys = [k for k in dist.slices if not (dist.slices[k]["flags"] & 0x100)]   # computed, not mirrored
    rnd2 = random.Random(1)
    sample = rnd2.sample(keys, min(args.sample, len(keys)))
    if not args.dim:
        # make sure the upgraded dimensions are in the sample: near the diagonal, top of the range
        near = [k for k in keys if abs(k[0] - k[1]) <= 1 and k[0] > 0 and max(k) >= m]
        sample = sample[:max(1, args.sample - 4)] + rnd2.sample(near, min(4, len(near)))
    D = dist.slices[sample[0]]["dimension"] if args.dim else 0
    dims = sorted({s["dimension"] for s in dist.slices.values()})
    jobs = [(m, args.s, d, r, D, k[0], k[1]) for k in sample]
    t2 = time.time()
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        refs = pool.map(_ref_slice, jobs)
    t_ref = time.time() - t2
    worst_cell = worst_mass = worst_err = 0.0
    for (a_d, a_r, Dr, cells, te, fl) in refs:
        s = dist.slices[(a_d, a_r)]
        assert s["dimension"] == Dr, (a_d, a_r, s["dimension"], Dr)
        floor = np.longdouble(1e-15) * np.max(np.abs(cells))
        worst_cell = max(worst_cell, float(np.max(np.abs(s["cells"] - cells) / (np.abs(cells) + floor))))
        worst_mass = max(worst_mass, abs(float(s["cells"].sum() - cells.sum())))
        worst_err = max(worst_err, abs(float((s["total_error"] - te) / te)))
        assert s["flags"] == fl, (a_d, a_r, s["flags"], fl)
    total_mass = float(sum(s["cells"].sum() for s in dist.slices.values()))
    rep = dict(command=" ".join(os.path.basename(c) for c in cmd[:4]) + " -exp <d> <r> " + " ".join(cmd[7:]),
               clients=args.clients, wall_s=wall, slices_in_file=len(dist.slices),
               dimension=D or "heuristic", stored_dimensions=dims,
               sampled_dimensions=sorted({r_[2] for r_ in refs}),
               file_bytes=size, files=files, total_mass=total_mass, parse_s=t_parse,
               sample=len(sample), ref_cpu_s=t_ref, worst_cell=worst_cell, worst_mass=worst_mass,
               worst_total_error_rel=worst_err,
               last_lines=lines[-6:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "full_distribution_report.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
    assert worst_cell <= 1e-9 and worst_mass <= 1e-12 and worst_err <= 1e-9


if __name__ == "__main__":
    main()
