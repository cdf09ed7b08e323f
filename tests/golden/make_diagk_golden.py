"""Generates tests/golden/diagk.npz: inputs and outputs of the UNMODIFIED reference's
sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412-646) through oracle/_ref, on seeded random
(j, eta, pivot) and on the edge cases the domain has (j = 0 and tiny j so that q + eta < 0, the
largest j, |eta| = 25, pivot 0 and 1, a delta_bound that the pivot outruns, l small enough for k
to wrap around, m not a multiple of 32, r shorter than m bits).

Run in the build container (needs /root/reference for oracle/_ref):
    python tests/golden/make_diagk_golden.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.set_int_max_str_digits(0)

from oracle import ref as R  # noqa: E402

CASES = [  # name, m, sigma, l, delta_bound
    ("m128_s5_l13", 128, 5, 13, 0xffffffff),
    ("m128_s0_l128", 128, 0, 128, 0xffffffff),
    ("m1023_s3_l512", 1023, 3, 512, 0xffffffff),
    ("m2048_s0_l2048", 2048, 0, 2048, 1000),
    ("m2048_s12_l1030", 2048, 12, 1030, 2),
    ("m2050_short_r", 2050, 4, 100, 0xffffffff),
    ("m4096_s2_l59", 4096, 2, 59, 0xffffffff),
]


def draw_d_r(rng, m, short_r=False):
    bits = m - 37 if short_r else m
    r = (1 << (bits - 1)) + 1 + rng.randrange((1 << (bits - 1)) - 1)
    d = r // 2 + rng.randrange(r // 2)
    return d, r


def main():
    rng = random.Random(20261017)
    out = {}
    names = []
    for name, m, sigma, l, delta_bound in CASES:
        d, r = draw_d_r(rng, m, short_r="short_r" in name)
        P = R.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
        n = m + sigma
        js, etas, pivots = [], [], []
        for _ in range(48):
            js.append(rng.randrange(1 << n))
            etas.append(rng.randrange(-25, 26))
            pivots.append(np.longdouble(rng.random()))
        # edge cases
        tiny = [0, 1, 3, (1 << n) - 1, (1 << (n - 1)), ((10 << n) // r), ((10 << n) // r) + 1]
        for j in tiny:
            for eta in (-25, 0, 25):
                js.append(j)
                etas.append(eta)
                pivots.append(np.longdouble(rng.random()))
        for p in (np.longdouble(0), np.longdouble(1), np.longdouble(1) - np.longdouble(2) ** -20):
            js.append(rng.randrange(1 << n))
            etas.append(rng.randrange(-25, 26))
            pivots.append(p)
        db = delta_bound
        ks, oks, alphas = [], [], []
        for j, eta, p in zip(js, etas, pivots):
            bound = db
            if p > np.longdouble(0.99) and db == 0xffffffff and l > 24:
                bound = 5000  # keep the reference's walk finite
            ok, k, a, _ = R.sample_k_from_diagonal_j_eta_pivot(P, p, j, eta, bound, precision=256)
            ks.append(k)
            oks.append(ok)
            alphas.append(a)
        bounds = [5000 if (p > np.longdouble(0.99) and db == 0xffffffff and l > 24) else db
                  for p in pivots]
        wj, wl = (n + 31) // 32, (l + 31) // 32
        out[name + "_params"] = np.array([m, sigma, l], dtype=np.int64)
        out[name + "_d"] = np.frombuffer(d.to_bytes((m + 7) // 8, "big"), dtype=np.uint8)
        out[name + "_r"] = np.frombuffer(r.to_bytes((m + 7) // 8, "big"), dtype=np.uint8)
        out[name + "_j"] = np.stack([np.frombuffer(j.to_bytes(4 * wj, "little"), dtype=np.uint32)
                                     for j in js])
        out[name + "_eta"] = np.array(etas, dtype=np.int32)
        out[name + "_pivot"] = np.array(pivots, dtype=np.longdouble)
        out[name + "_bound"] = np.array(bounds, dtype=np.uint64)
        out[name + "_k"] = np.stack([np.frombuffer(k.to_bytes(4 * wl, "little"), dtype=np.uint32)
                                     for k in ks])
        out[name + "_ok"] = np.array(oks, dtype=np.uint8)
        out[name + "_alpha"] = np.array(alphas, dtype=np.longdouble)
        names.append(name)
        print(name, "ok", sum(oks), "of", len(oks))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "diagk.npz"), **out)


if __name__ == "__main__":
    main()
