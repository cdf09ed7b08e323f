"""estimate_runs_distribution on a FULL m = 2048, s = 1 two-dimensional distribution (BASELINE
configs[1]: 6404 slices of dimension 128, 3.1 GB of text), the reference's own executable in both
flavours of integration/build.py:

  ref   tau_estimate.cpp of the reference, one client rank per host core
  gpu   qunundrum_b200/dropin/dropin_tau.cpp (+ the integrator and text drop-ins), a few client
        ranks sharing one B200

    python tests/tools/estimate_runs_timing.py [--ref-clients 16] [--gpu-clients 8] [--skip-ref]

Writes gpurun_out/estimate_runs_report_s<s>.json (wall clocks, the log lines of both runs).
"""
import argparse
import json
import os
import random
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = os.path.join(ROOT, "integration", "_build")


def run(flavour, exe, args, np_, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.time()
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(np_), os.path.join(B, flavour, exe), *args],
                       cwd=cwd, env=e, capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit(1)
    marks = {}
    return p.stdout, p.stderr, wall, marks


def log_lines(cwd):
    out = []
    d = os.path.join(cwd, "logs")
    for f in sorted(os.listdir(d)):
        out += [l.strip() for l in open(os.path.join(d, f)) if l.startswith("m:")]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-clients", type=int, default=min(16, os.cpu_count() or 2))
    ap.add_argument("--gpu-clients", type=int, default=8)
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--m", type=int, default=2048)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--s", type=int, default=1, help="tradeoff factor: l = ceil(m / s); large s = many runs n")
    args = ap.parse_args()
    m = args.m
    rnd = random.Random(20482048)
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    t = tempfile.mkdtemp()
    os.makedirs(os.path.join(t, "distributions"))
    env = {"QB200_DEVICE": "0", "QB200_TEXT_DEVICE": "0", "QB200_DROPIN_STATS": "1"}
    _, _, gen_wall, _ = run("gpu", "generate_distribution",
                            ["-exp", str(d), str(r), "-dim", str(args.dim), str(m), str(args.s)], 3, t, env)
    name = [f for f in os.listdir(os.path.join(t, "distributions"))
            if f.startswith("distribution-") and f.endswith(".txt")][0]
    path = os.path.join("distributions", name)
    size = os.path.getsize(os.path.join(t, path))
    rep = {"distribution": name, "file_bytes": size, "generate_wall_s": gen_wall, "m": m, "s": args.s, "runs": {}}
    for flavour, clients in (("gpu", args.gpu_clients), ("ref", args.ref_clients)):
        if flavour == "ref" and args.skip_ref:
            continue
        cwd = os.path.join(t, flavour)
        os.makedirs(cwd)
        os.symlink(os.path.join(t, "distributions"), os.path.join(cwd, "distributions"))
        out, err, wall, _ = run(flavour, "estimate_runs_distribution", [path], clients + 1, cwd, env)
        lines = log_lines(cwd)
        rep["runs"][flavour] = {"clients": clients, "wall_s": wall, "log": lines,
                                "stderr_tail": err.strip().splitlines()[-12:]}
        print(flavour, f"{wall:.1f} s with {clients} clients")
        for l in lines:
            print("   ", l)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", f"estimate_runs_report_s{args.s}.json"), "w"), indent=1)
    shutil.rmtree(t, ignore_errors=True)


if __name__ == "__main__":
    main()
