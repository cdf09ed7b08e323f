"""tau_estimate_diagonal: the drop-in against the reference's own function at m = 2048, in one process
on identically seeded generators (integration/tools/tau_diagonal_check.cpp), with timings:

    python tests/tools/tau_diagonal_timing.py [--estimates 2000] [--n 8] > gpurun_out/tau_diagonal_m2048.json

Generates the distribution with the reference's generate_diagonal_distribution (drop-in flavour, on the
GPU), then runs `estimates` estimates of n samples through both functions: flags, tau and the generator
state at every batch boundary must agree; reports the seconds each side took (the reference: one host
core) and the drop-in's own split (QB200_DROPIN_STATS).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
IB = os.path.join(ROOT, "integration", "_build")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--estimates", type=int, default=2000)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--m", type=int, default=2048)
    ap.add_argument("--sigma", type=int, default=5)
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=1000)
    a = ap.parse_args()
    env = dict(os.environ, QB200_DEVICE="0", QB200_TEXT_DEVICE="0")
    out = {"config": vars(a)}
    with tempfile.TemporaryDirectory() as t:
        os.makedirs(os.path.join(t, "distributions"))
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(IB, "minimpirun"), "-np", "3",
                            os.path.join(IB, "gpu", "generate_diagonal_distribution"), "-dim", str(a.dim),
                            "-eta-bound", "2", "-det", str(a.m), str(a.sigma), "1"],
                           cwd=t, env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-1000:]
        out["generate_s"] = time.perf_counter() - t0
        dist = os.path.join(t, "distributions", os.listdir(os.path.join(t, "distributions"))[0])
        out["distribution"] = os.path.basename(dist)
        for label, extra in (("cold", []), ("second_run_same_process_caches_cold_again", [])):
            env2 = dict(env, QB200_TAU_BATCH=str(a.batch), QB200_DROPIN_STATS="1")
            p = subprocess.run([os.path.join(IB, "gpu", "tau_diagonal_check"), dist, str(a.n), str(a.estimates),
                                "1000000", "2", "1", str(a.batch)], env=env2, capture_output=True, text=True)
            res = json.loads(p.stdout.strip().splitlines()[-1])
            res["returncode"] = p.returncode
            m = re.search(r"diagonal tau drop-in: (.*)", p.stderr)
            res["dropin_stats"] = m.group(1) if m else p.stderr[-500:]
            res["speedup_vs_one_core"] = res["reference_s"] / res["dropin_s"]
            res["samples_per_s_dropin"] = a.estimates * a.n / res["dropin_s"]
            res["samples_per_s_reference_one_core"] = a.estimates * a.n / res["reference_s"]
            out[label] = res
            break
    print(json.dumps(out))
    sys.exit(0 if out["cold"]["ok"] else 1)


if __name__ == "__main__":
    main()
