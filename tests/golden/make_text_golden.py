"""Generate the committed golden fixtures of the slice TEXT format from the UNMODIFIED
reference exporters (oracle/_ref/libqref.so: distribution_slice_export,
linear_distribution_slice_export, diagonal_distribution_slice_export writing to a
memory stream) and from the libc calls they make.

    python tests/golden/make_text_golden.py        (build container: needs /root/reference)

Writes under tests/golden/text/:
  slice_2d.txt / slice_linear.txt / slice_diagonal.txt
        the reference's export of three golden slices (tests/golden/slices.npz), with
        slice_*.npz holding the exact x87 bit patterns (mantissa, sign|exponent) of the
        values that were exported (cells, then total_error) and the header fields
  adversarial.npz + adversarial.txt
        hand-picked x87 bit patterns (ties, decade and style boundaries, extremes,
        denormals, specials) and fprintf("%.24Lg\\n") of each
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import text as ot  # noqa: E402
from tests.conftest import golden_slices  # noqa: E402

OUT = os.path.join(HERE, "text")


def adversarial():
    L = np.longdouble
    two = L(2)
    v = []
    # exact short decimals, powers of two, powers of ten around the style boundaries
    for e in range(-70, 70):
        v += [two ** e, L(3) * two ** e, -(L(5) * two ** e)]
    for e in range(-30, 30):
        v.append(L(10) ** e)
    v += [L(x) for x in (0.1, 0.25, 0.5, 1.5, 1e-4, 9.9999e-5, 1e-5, 123456.789, 1e23, 1e24,
                         0.000123, 999999.5, 1e22)]
    # true ties at the 24th digit: (odd integer with 19 digits) / 2^6 etc.
    for k in range(1, 30):
        base = (1 << 63) + 2 * k * 982451653 + 1
        for q in range(-12, 0):
            v.append(L(base) * two ** q)
            v.append(L(base + 2) * two ** q)
    v += [L(1) + two ** -24, L(1) + two ** -23, L(3) + two ** -24]
    a = np.array(v, dtype=np.longdouble)
    mant, se = ot.ld_fields(a)
    mant, se = list(mant), list(se)
    # just below / above powers of ten, extremes, denormals, specials as bit patterns
    for x in (-4931, -4000, -1000, -310, -100, -60, -5, -4, 0, 17, 23, 24, 100, 1000, 4931):
        m, s = ot.parse_ld_exact(f"1e{x}".encode())
        for dm in (-2, -1, 0, 1, 2):
            mm = m + dm
            if 0 < mm < 2 ** 64 and (mm >> 63) == 1:
                mant.append(np.uint64(mm)); se.append(np.uint16(s))
    for m, s in ((1 << 63, 1), (2 ** 64 - 1, 0x7FFE), (1, 0), (2 ** 63 - 1, 0), (1 << 63, 0),
                 (0, 0), (0, 0x8000), (1 << 63, 0x7FFF), (1 << 63, 0xFFFF), (3 << 62, 0x7FFF),
                 (3 << 62, 0xFFFF), (12345, 0), (2 ** 64 - 1, 1), (2 ** 64 - 1, 0x8001)):
        mant.append(np.uint64(m)); se.append(np.uint16(s))
    return np.array(mant, dtype=np.uint64), np.array(se, dtype=np.uint16)


def main():
    os.makedirs(OUT, exist_ok=True)
    G = golden_slices()
    picks = (("2d", 0, next(g for g in G if g.meta["name"].startswith("2d/c2/"))),
             ("linear", 1, next(g for g in G if g.meta["name"].startswith("lin/c1/t0/"))),
             ("diagonal", 2, next(g for g in G if g.meta["name"].startswith("diag/"))))
    for name, kind, g in picks:
        q = g.meta
        D = q["D"]
        if kind == 0:
            c0, c1 = q["a_d"], q["a_r"]
        elif kind == 1:
            c0, c1 = q["a"], 0
        else:
            c0, c1 = q["a"], q["eta"]
        text = ot.ref_slice_export(kind, D, c0, c1, g.flags, g.cells, g.total_error)
        open(os.path.join(OUT, f"slice_{name}.txt"), "wb").write(text)
        vals = np.concatenate([np.asarray(g.cells, dtype=np.longdouble),
                               np.array([g.total_error], dtype=np.longdouble)])
        mant, se = ot.ld_fields(vals)
        np.savez_compressed(os.path.join(OUT, f"slice_{name}.npz"), mant=mant, se=se,
                            head=np.array([D, c0, c1, g.flags], dtype=np.int64))
        print(name, q["name"], len(text), "bytes")
    mant, se = adversarial()
    text = ot.format_ld24(ot.ld_from_fields(mant, se))
    np.savez_compressed(os.path.join(OUT, "adversarial.npz"), mant=mant, se=se)
    open(os.path.join(OUT, "adversarial.txt"), "wb").write(text)
    print("adversarial", mant.size, "values,", len(text), "bytes,", ot.libc_version())


if __name__ == "__main__":
    main()
