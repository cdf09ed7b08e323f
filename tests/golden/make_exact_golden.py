"""Generates tests/golden/exact.json: inputs and outputs of the UNMODIFIED reference's exact samplers
(src/sample.cpp:78-410: sample_alpha_from_region, sample_j_from_alpha_r, sample_j_k_from_alpha_d,
sample_j_k_from_alpha_d_r, sample_j_from_diagonal_alpha_r) through oracle/_ref, each case on one seeded
Random_State whose bytes are stored beside the results: regions in the order they were sampled (so that
the stream layout is part of what is pinned), then the (j, k) functions on the sampled arguments.
Cases: m = 34 ... 2048, dimensions 16 ... 2048, odd and even d / r (kappa up to 40, kappa_t_r > 0),
both signs, first and last regions of a slice, |log alpha| from 8 (where the reference's 3 (e + 1)-bit
rounding of the bounds is visible) to m + 60.

Run in the build container (needs /root/reference for oracle/_ref):
    python tests/golden/make_exact_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.set_int_max_str_digits(0)

from oracle import ref as R  # noqa: E402
from tests.test_exact import d_r_with_kappa, mpz_bytes  # noqa: E402

# name, kind (0 two-dimensional / 1 diagonal), m, s, sigma, kappa_d, kappa_r, dimension, e range, samples
CASES = [
    ("2d_m128_s1", 0, 128, 1, 0, 0, 0, 256, (98, 180), 40),
    ("2d_m128_s2_even", 0, 128, 2, 0, 3, 5, 64, (98, 180), 40),
    ("2d_m160_s3_kd33", 0, 160, 3, 0, 33, 0, 16, (130, 210), 30),
    ("2d_m128_s4_ktr", 0, 128, 4, 0, 1, 40, 32, (98, 150), 30),
    ("2d_m2048_s1", 0, 2048, 1, 0, 0, 2, 256, (2018, 2100), 12),
    ("diag_m64_sigma6", 1, 64, 2, 6, 0, 0, 128, (8, 69), 60),
    ("diag_m34_sigma3", 1, 34, 1, 3, 0, 4, 1024, (12, 36), 60),
    ("diag_m2048_sigma5", 1, 2048, 1, 5, 0, 0, 2048, (2018, 2052), 12),
]


def hx(v):
    return format(v, "x")


def main():
    out = []
    for name, kind, m, s, sigma, kd, kr, D, (e_lo, e_hi), count in CASES:
        g = np.random.default_rng(sum(map(ord, name)))
        d, r = d_r_with_kappa(g, m, kd, kr)
        if kind == 1:
            P = R.RefDiagonalParameters(m, sigma, s, d, r)
            l = -(-m // s)
        else:
            P = R.RefParameters(m, s, d, r)
            l = P.l
        seed = bytes(g.integers(0, 256, 32, dtype=np.uint8))
        rng, twin = R.RefRandom(seed), R.RefRandom(seed)
        stream = twin.bytes(count * 2 * ((e_hi + 80) // 8 + 64) + 4096)
        off = 0
        samples = []
        for it in range(count):
            rec = {}
            args = []
            for which, kap in (("d", kd), ("r", kr)):
                e = int(g.integers(e_lo, e_hi))
                sign = -1 if g.integers(2) else 1
                reg = (0, D - 1)[it % 2] if it % 5 == 0 else int(g.integers(0, D))
                with_lo, with_hi = sign * (e + reg / D), sign * (e + (reg + 1) / D)
                a = R.sample_alpha_from_region(with_lo, with_hi, kap, rng)
                args.append(a)
                rec[which] = {"min_log_alpha": sign * e, "region": reg, "kappa": kap, "alpha": hx(a)}
                # the bytes the reference read follow from its modulus (src/random.c:163-164): the bounds at
                # its precision of 3 (e + 1) bits; the check after the loop confirms the whole layout
                with mp.workprec(3 * (e + 1)):
                    lo_b = int(mp.nint(mp.mpf(2) ** mp.mpf(abs(with_lo))))
                    hi_b = int(mp.nint(mp.mpf(2) ** mp.mpf(abs(with_hi))))
                ln = mpz_bytes((hi_b - lo_b).bit_length())
                rec[which]["offset"], rec[which]["length"] = off, ln
                off += ln
            if kind == 1:
                # sample_j_from_diagonal_alpha_r on the r argument
                t = 0
                if kr > 0:
                    ln = mpz_bytes(kr + 1)
                    t = int.from_bytes(stream[off:off + ln], "big") % (1 << kr)
                    off += ln
                j, _ = R.sample_j_k(3, P, None, args[1], rng)
                rec["t_r"], rec["j_diagonal"] = hx(t), hx(j)
            else:
                # mode 0, then mode 1, then mode 2, each drawing what the reference draws
                t0 = 0
                if kr > 0:
                    ln = mpz_bytes(kr + 1)
                    t0 = int.from_bytes(stream[off:off + ln], "big") % (1 << kr)
                    off += ln
                j0, _ = R.sample_j_k(0, P, None, args[1], rng)
                kt = max(0, kr - kd - l)
                t1 = 0
                if kr > 0:
                    ln = mpz_bytes(kr - kt + 1)
                    t1 = (int.from_bytes(stream[off:off + ln], "big") % (1 << (kr - kt))) << kt
                    off += ln
                j1, k1 = R.sample_j_k(1, P, args[0], args[1], rng)
                t2 = 0
                if kd > 0:
                    ln = mpz_bytes(kd + 1)
                    t2 = int.from_bytes(stream[off:off + ln], "big") % (1 << kd)
                    off += ln
                ln = mpz_bytes(l + 1)
                k2 = int.from_bytes(stream[off:off + ln], "big") % (1 << l)
                off += ln
                j2, k2r = R.sample_j_k(2, P, args[0], None, rng)
                assert k2r == k2
                rec.update({"t_r": hx(t0), "j_from_alpha_r": hx(j0), "t_r_scaled": hx(t1), "j_k_from_alpha_d_r": [hx(j1), hx(k1)],
                            "t_d": hx(t2), "k_drawn": hx(k2), "j_from_alpha_d_k": hx(j2)})
            samples.append(rec)
        # the layout above is the reference's: the next bytes of its generator are the stream's
        assert rng.bytes(16) == stream[off:off + 16], name
        out.append({"name": name, "kind": kind, "m": m, "l": l, "sigma": sigma, "d": hx(d), "r": hx(r),
                    "kappa_d": kd, "kappa_r": kr, "dimension": D, "stream": stream[:off].hex(), "samples": samples})
        print(name, len(samples), "samples,", off, "bytes")
    path = os.path.join(ROOT, "tests", "golden", "exact.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
