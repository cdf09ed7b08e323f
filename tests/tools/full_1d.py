"""The one-dimensional BASELINE configurations at FULL size with the reference's own generator
executables + both drop-ins, checked against the reference itself (oracle/_ref) on a sample of
slices on the host cores.

    python tests/tools/full_1d.py rsa       # config 3: generate_linear_distribution_rsa, n = 2048
                                            #   (m = 1023, l = 1003), synthetic p and q
    python tests/tools/full_1d.py sweep     # config 3: generate_linear_distribution -d -exp <d_rsa>
                                            #   1023 s for s = 1 .. 8 in ONE run
    python tests/tools/full_1d.py diagonal  # config 5: generate_diagonal_distribution, m = 2048,
                                            #   sigma in {0, 5, 12}, eta-bound 2, in ONE run
    python tests/tools/full_1d.py linear    # config 1 at full dimension: m = 128, s = 2, -d and -r

Appends a report to gpurun_out/full_1d_report.json.
"""
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


def _probable_prime(n, rnd, rounds=24):
    if n < 4:
        return n in (2, 3)
    for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % sp == 0:
            return n == sp
    d, s = n - 1, 0
    while d % 2 == 0:
        d, s = d // 2, s + 1
    for _ in range(rounds):
        x = pow(rnd.randrange(2, n - 1), d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def synthetic_p_q(bits, seed):
    """Two random primes of `bits` bits (the generator checks primality, src/
    main_generate_linear_distribution_rsa.cpp:270-290): a synthetic RSA modulus."""
    rnd = random.Random(seed)
    out = []
    while len(out) < 2:
        c = ((1 << (bits - 1)) + rnd.randrange(1 << (bits - 1))) | 1
        while not _probable_prime(c, rnd):
            c += 2
        if c.bit_length() == bits and c not in out:
            out.append(c)
    return out[0], out[1]


def _ref_job(job):
    from oracle import ref
    kind, h, key, D = job
    if kind == "diagonal":
        P = ref.RefDiagonalParameters(h["m"], h["sigma"], h["s"], h["d"], h["r"],
                                      eta_bound=h["eta_bound"], t=h["t"],
                                      l=0 if h["s"] else h["l"])
        sl = ref.diagonal_distribution_slice_compute(P, D, key[0], key[1])
    else:
        P = ref.RefParameters(h["m"], h["s"], h["d"], h["r"], t=h["t"], l=0 if h["s"] else h["l"])
        sl = ref.linear_distribution_slice_compute(P, D, key[0], 0 if kind == "linear_d" else 1)
    return key, np.asarray(sl.cells, dtype=np.longdouble), sl.flags


def run(exe, args, ranks, cwd):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    env = dict(os.environ, QB200_DEVICE="0")
    t0 = time.time()
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(ranks), os.path.join(B, "gpu", exe), *args],
                       cwd=cwd, env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit(1)
    d = os.path.join(cwd, "distributions")
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".txt"))
    if not files:      # the generators return 0 after printing an argument error
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit("no distribution was written")
    return wall, files


def check(path, kind, n_sample, pool, seed=1):
    fmt = "diagonal" if kind == "diagonal" else "linear"
    dist = distfile.read(path, fmt)
    keys = [k for k in dist.slices if not (dist.slices[k]["flags"] & 0x100)]   # computed, not mirrored
    sample = random.Random(seed).sample(keys, min(n_sample, len(keys)))
    # always include the heaviest slice
    heavy = max(keys, key=lambda k: float(dist.slices[k]["cells"].sum()))
    if heavy not in sample:
        sample[0] = heavy
    jobs = [(kind, dist.header, k, dist.slices[k]["dimension"]) for k in sample]
    t0 = time.time()
    refs = pool.map(_ref_job, jobs)
    t_ref = time.time() - t0
    worst_cell = worst_mass = 0.0
    for key, cells, fl in refs:
        s = dist.slices[key]
        floor = np.longdouble(1e-15) * np.max(np.abs(cells))
        worst_cell = max(worst_cell, float(np.max(np.abs(s["cells"] - cells) / (np.abs(cells) + floor))))
        worst_mass = max(worst_mass, abs(float(s["cells"].sum() - cells.sum())))
        assert s["flags"] == fl, (key, s["flags"], fl)
    h = dist.header
    rep = dict(file=os.path.basename(path), kind=kind, m=h["m"], s=h["s"], l=h["l"],
               sigma=h.get("sigma"), slices=len(dist.slices), computed_slices=len(keys),
               dimension=jobs[0][3],
               total_mass=float(sum(s["cells"].sum() for s in dist.slices.values())),
               sampled=len(sample), ref_cpu_s=round(t_ref, 1), worst_cell=worst_cell,
               worst_mass=worst_mass)
    assert worst_cell <= 1e-9 and worst_mass <= 1e-12, rep
    return rep


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "rsa"
    n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    t = tempfile.mkdtemp()
    p, q = synthetic_p_q(1024, 20482048)
    m = 1023
    d_rsa = (p - 1) // 2 + (q - 1) // 2 - 2 ** m
    reports = []
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        if which == "rsa":
            wall, files = run("generate_linear_distribution_rsa", ["-exp", str(p), str(q), "2048"], 3, t)
            reports = [check(f, "linear_d", n_sample, pool) for f in files]
        elif which == "sweep":
            args = ["-d", "-exp", str(d_rsa)]
            for s in range(1, 9):
                args += ["1023", str(s)]
            wall, files = run("generate_linear_distribution", args, 3, t)
            reports = [check(f, "linear_d", max(2, n_sample // 4), pool) for f in files]
        elif which == "diagonal":
            rnd = random.Random(20482048)
            r = 2 ** 2047 + 1 + rnd.randrange(2 ** 2047 - 1)
            d = r // 2 + rnd.randrange(r // 2)
            args = ["-eta-bound", "2", "-exp", str(d), str(r)]
            for sigma in (0, 5, 12):
                args += ["2048", str(sigma), "1"]
            wall, files = run("generate_diagonal_distribution", args, 3, t)
            reports = [check(f, "diagonal", n_sample, pool) for f in files]
        elif which == "linear":
            wall, files = run("generate_linear_distribution", ["-d", "-det", "128", "2"], 2, t)
            reports = [check(f, "linear_d", n_sample, pool) for f in files]
            w2, files = run("generate_linear_distribution", ["-r", "-det", "128", "2"], 2,
                            os.path.join(t, "r"))
            wall += w2
            reports += [check(f, "linear_r", n_sample, pool) for f in files]
        else:
            raise SystemExit("unknown configuration")
    rep = dict(configuration=which, generate_wall_s=round(wall, 2), distributions=reports)
    print(json.dumps(rep, indent=1))
    out = os.path.join(ROOT, "gpurun_out", "full_1d_report.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    prev = json.load(open(out)) if os.path.exists(out) else []
    json.dump(prev + [rep], open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
