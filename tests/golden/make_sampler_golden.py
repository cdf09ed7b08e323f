"""Golden fixtures of the reference's SAMPLING path, generated from the
UNMODIFIED reference (oracle/_ref/libqref.so, built from /root/reference/src in place):

    python tests/golden/make_sampler_golden.py        ->  tests/golden/sampler.npz

For a two-dimensional distribution (m = 64, s = 2, 392 Richardson slices of dimensions 8, 16 and 32 and their 392 mirror images computed
by the reference's own distribution_slice_compute_richardson, sorted by
distribution_sort_slices) and a linear one (m = 128, s = 2, 38 slices of dimension 64):

  * the distribution itself (slice order, coordinates, cells and totals as raw x87 bytes);
  * a stream of 64-bit words drawn from the reference's Keccak generator (random_generate,
    src/random.c:88) seeded by keccak_random_init_seed (src/keccak_random.c:52);
  * what the reference returns on that stream: distribution_sample_region /
    linear_distribution_sample_region (regions + success flags),
    distribution_sample_approximate_alpha_d_r / linear_..._alpha (alpha / 2^m as long double),
    and tau_estimate / tau_estimate_linear for several n (each from a fresh generator with the
    same seed), also with the distribution's total_probability forced above one
    (src/distribution.cpp:373-381).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

LD = np.longdouble
SEED = bytes((7 * i + 3) & 0xff for i in range(32))


def raw(x):
    return np.ascontiguousarray(x, dtype=LD).view(np.uint8).copy()


def describe_sorted(dist, slices, dims):
    dim, a, b, tot, total = dist.describe()
    key = {(int(s[1]), int(s[2])): s for s in slices}
    ordered = [key[(int(a[i]), int(b[i]))] for i in range(len(dim))]
    cells = np.concatenate([s[3] for s in ordered])
    return dim, a, b, tot, total, cells


def record(out, name, dims, P, slices, taus, n_samples):
    dist = ref.RefDistribution(dims, P, [s[0] for s in slices], [s[1] for s in slices],
                               [s[2] for s in slices], np.concatenate([s[3] for s in slices]))
    dist.sort()
    dim, a, b, tot, total, cells = describe_sorted(dist, slices, dims)
    out[name + "/m"] = np.array([P.m])
    out[name + "/dims"] = np.array([dims])
    out[name + "/dimension"] = dim
    out[name + "/c0"] = a
    out[name + "/c1"] = b
    out[name + "/cells"] = raw(cells)
    out[name + "/totals"] = raw(tot)
    out[name + "/total"] = raw([total])
    wps = dims + 2
    n_words = max(n_samples * wps, max(n * c for n, c in taus) * wps)
    out[name + "/words"] = ref.RefRandom(SEED).words(n_words)
    reg, ok = dist.sample_region(ref.RefRandom(SEED), n_samples)
    out[name + "/region"] = reg
    out[name + "/region_ok"] = ok
    a0, a1, ok = dist.sample_alpha(ref.RefRandom(SEED), n_samples)
    out[name + "/alpha0"] = raw(np.where(ok, a0, 0))
    out[name + "/alpha1"] = raw(np.where(ok, a1, 0))
    out[name + "/alpha_ok"] = ok
    for n, count in taus:
        t0, t1, ok = dist.tau_estimate(ref.RefRandom(SEED), n, count)
        out[f"{name}/tau/{n}/t0"] = raw(t0)
        out[f"{name}/tau/{n}/t1"] = raw(t1)
        out[f"{name}/tau/{n}/ok"] = ok
    # total_probability > 1: the slice pivot is scaled (the walk itself is unchanged)
    big = LD(1.25)
    dist.set_total(big)
    out[name + "/big_total"] = raw([big])
    t0, t1, ok = dist.tau_estimate(ref.RefRandom(SEED), 4, 200)
    out[name + "/big/t0"] = raw(t0)
    out[name + "/big/t1"] = raw(t1)
    out[name + "/big/ok"] = ok
    print(name, "slices", len(dim), "mass", float(total), "negative cells", int((cells < 0).sum()),
          "region ok", float(out[name + "/region_ok"].mean()))


def main():
    ref.build()
    out = {}
    # two-dimensional: m = 64, s = 2
    m = 64
    d, r = ref.deterministic_d_r(m)
    P = ref.RefParameters(m, 2, d, r)
    sl = []
    for a in range(m - 11, m + 3):
        for b in range(m - 11, m + 3):
            for sg in (1, -1):
                D = 32 if (abs(a - m) <= 1 and abs(b - m) <= 1) else (16 if min(a, b) >= m - 4 else 8)
                x = ref.distribution_slice_compute(P, D, sg * a, b)
                sl.append((D, sg * a, b, np.asarray(x.cells, dtype=LD)))
                # the mirrored slice the generator's server adds (same cells at (-alpha_d, -alpha_r),
                # src/main_generate_distribution.cpp:1006-1016): brings the mass to ~0.97
                sl.append((D, -sg * a, -b, np.asarray(x.cells, dtype=LD)))
    record(out, "2d", 2, P, sl, [(1, 300), (3, 300), (10, 200), (64, 50)], 4000)
    # linear (target d): m = 128, s = 2
    m = 128
    d, r = ref.deterministic_d_r(m)
    P = ref.RefParameters(m, 2, d, r)
    sl = []
    for a in range(m - 12, m + 7):
        for sg in (1, -1):
            x = ref.linear_distribution_slice_compute(P, 64, sg * a, 0)
            sl.append((64, sg * a, 0, np.asarray(x.cells, dtype=LD)))
    record(out, "lin", 1, P, sl, [(1, 300), (5, 300), (32, 100)], 4000)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("wrote", os.path.join(HERE, "sampler.npz"), os.path.getsize(os.path.join(HERE, "sampler.npz")), "bytes")


if __name__ == "__main__":
    main()
