"""Oracle of the slice text format.  TEST INFRASTRUCTURE ONLY (oracle/README.md).

Three checkers, strongest first:

* the UNMODIFIED reference exporters / importers (oracle/_ref/libqref.so:
  distribution_slice_export, src/distribution_slice_import_export.cpp:89-103, and the
  linear / diagonal twins) writing to / reading from memory streams;
* the libc calls they make, fprintf("%.24Lg\\n") and fscanf("%Lg\\n")
  (oracle/text_oracle.c -> oracle/_build/libtextoracle.so).  The arithmetic is a
  third-party dependency absent from /root/reference: GNU libc 2.39
  (stdio-common/printf_fp.c, stdlib/strtod_l.c);
* a restatement in exact Python integers of what those calls are specified to
  do (ISO C 7.21.6.1 %g with precision 24; correct rounding, ties to even):
  `format_ld24_exact`, `parse_ld_exact`.  Pinned against libc in
  tests/test_text_format.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libtextoracle.so")
_lib = None


def build() -> str:
    src = os.path.join(_HERE, "text_oracle.c")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", src, "-o", LIB_PATH])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.text_oracle_format_ld.restype = C.c_size_t
        L.text_oracle_format_ld.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.text_oracle_parse_ld.restype = C.c_size_t
        L.text_oracle_parse_ld.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.text_oracle_libc.restype = C.c_char_p
        _lib = L
    return _lib


def libc_version() -> str:
    return lib().text_oracle_libc().decode()


def format_ld24(values) -> bytes:
    """fprintf("%.24Lg\\n") per value (libc)."""
    v = np.ascontiguousarray(values, dtype=np.longdouble)
    cap = 40 * v.size + 64
    out = C.create_string_buffer(cap)
    n = lib().text_oracle_format_ld(v.ctypes.data, v.size, out, cap)
    assert n != C.c_size_t(-1).value
    return out.raw[:n]


def parse_ld(text: bytes, n: int) -> np.ndarray:
    """fscanf("%Lg\\n") n times (libc)."""
    v = np.zeros(n, dtype=np.longdouble)
    got = lib().text_oracle_parse_ld(text, len(text), v.ctypes.data, n)
    if got != n:
        raise ValueError(f"parsed {got} of {n} values")
    return v


# ---- x87 bit patterns ------------------------------------------------------------

def ld_fields(values):
    """(mantissa uint64, sign|exponent uint16) of each long double."""
    v = np.ascontiguousarray(values, dtype=np.longdouble).reshape(-1)
    b = v.view(np.uint8).reshape(-1, 16)
    mant = b[:, :8].copy().view(np.uint64).reshape(-1)
    se = b[:, 8:10].copy().view(np.uint16).reshape(-1)
    return mant, se


def ld_from_fields(mant, se) -> np.ndarray:
    mant = np.ascontiguousarray(mant, dtype=np.uint64).reshape(-1)
    se = np.ascontiguousarray(se, dtype=np.uint16).reshape(-1)
    v = np.zeros(mant.size, dtype=np.longdouble)
    b = v.view(np.uint8).reshape(-1, 16)
    b[:, :8] = mant.view(np.uint8).reshape(-1, 8)
    b[:, 8:10] = se.view(np.uint8).reshape(-1, 2)
    return v


# ---- exact restatement -----------------------------------------------------------

def _render(neg: bool, digits: int, x: int) -> bytes:
    """digits: 24-digit integer (10^23 <= digits < 10^24), x: exponent of its first digit."""
    s = str(digits).rstrip("0") or "0"
    sign = "-" if neg else ""
    if x < -4 or x >= 24:                      # %e style, precision 23, zeros removed
        body = s[0] + ("." + s[1:] if len(s) > 1 else "")
        return f"{sign}{body}e{'-' if x < 0 else '+'}{abs(x):02d}\n".encode()
    if x >= 0:                                 # %f style, precision 23 - x
        s = s.ljust(x + 1, "0")
        return (sign + s[:x + 1] + ("." + s[x + 1:] if len(s) > x + 1 else "") + "\n").encode()
    return (sign + "0." + "0" * (-x - 1) + s + "\n").encode()


def format_ld24_exact(mant: int, se: int) -> bytes:
    """'%.24Lg\\n' of the x87 value (mant, se) by exact rational arithmetic."""
    neg = bool(se >> 15)
    e = se & 0x7FFF
    sign = "-" if neg else ""
    if e == 0x7FFF:
        return (sign + ("inf" if (mant << 1) & (2 ** 64 - 1) == 0 else "nan") + "\n").encode()
    if mant == 0:
        return (sign + "0\n").encode()
    q = (e if e else 1) - 16383 - 63
    val = Fraction(mant) * (Fraction(2) ** q)
    # x = floor(log10(val))
    x = int((mant.bit_length() + q - 1) * 0.30102999566398) - 2
    while Fraction(10) ** (x + 1) <= val:
        x += 1
    scaled = val / (Fraction(10) ** (x - 23))   # in [10^23, 10^24)
    d, rem = divmod(scaled.numerator, scaled.denominator)
    twice = 2 * rem
    if twice > scaled.denominator or (twice == scaled.denominator and (d & 1)):
        d += 1
    if d == 10 ** 24:
        d = 10 ** 23
        x += 1
    return _render(neg, d, x)


def format_ld24_exact_array(values) -> bytes:
    mant, se = ld_fields(values)
    return b"".join(format_ld24_exact(int(m), int(s)) for m, s in zip(mant, se))


def parse_ld_exact(token: bytes):
    """(mant, se) of strtold(token): decimal -> nearest x87 value, ties to even."""
    t = token.strip().decode().lower()
    neg = t.startswith("-")
    t = t.lstrip("+-")
    if t.startswith("inf"):
        return 1 << 63, 0x7FFF | (0x8000 if neg else 0)
    if t.startswith("nan"):
        return 3 << 62, 0x7FFF | (0x8000 if neg else 0)
    val = Fraction(t)
    sb = 0x8000 if neg else 0
    if val == 0:
        return 0, sb
    # binary exponent e2 with 2^e2 <= val < 2^(e2+1)
    e2 = val.numerator.bit_length() - val.denominator.bit_length()
    if Fraction(2) ** e2 > val:
        e2 -= 1
    elif Fraction(2) ** (e2 + 1) <= val:
        e2 += 1
    q = max(e2 - 63, -16382 - 63)            # exponent of the last mantissa bit
    scaled = val / (Fraction(2) ** q)
    m, rem = divmod(scaled.numerator, scaled.denominator)
    twice = 2 * rem
    if twice > scaled.denominator or (twice == scaled.denominator and (m & 1)):
        m += 1
    if m == 1 << 64:
        m >>= 1
        q += 1
    if m < 1 << 63:                            # denormal (or zero after rounding)
        return m, sb
    e = q + 63 + 16383
    if e >= 0x7FFF:
        return 1 << 63, 0x7FFF | sb
    return m, e | sb


# ---- the reference's own exporters / importers (libqref) ---------------------------

def ref_slice_export(kind: int, dimension: int, c0: int, c1: int, flags: int, cells,
                     total_error) -> bytes:
    """kind 0: distribution_slice_export, 1: linear_..., 2: diagonal_... (unmodified reference)."""
    from oracle import ref
    L = ref.lib()
    L.qref_slice_export.restype = C.c_size_t
    L.qref_slice_export.argtypes = [C.c_int, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                    C.c_void_p, C.c_longdouble, C.c_void_p, C.c_size_t]
    v = np.ascontiguousarray(cells, dtype=np.longdouble)
    cap = 40 * (v.size + 8)
    out = C.create_string_buffer(cap)
    n = L.qref_slice_export(kind, dimension, c0, c1, flags, v.ctypes.data,
                            C.c_longdouble(total_error), out, cap)
    assert n != C.c_size_t(-1).value
    return out.raw[:n]


def ref_slice_import(kind: int, text: bytes, max_cells: int):
    """The reference's *_slice_init_import on `text`: (head4, cells, total_probability, total_error)."""
    from oracle import ref
    L = ref.lib()
    L.qref_slice_import.restype = C.c_int
    L.qref_slice_import.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_uint32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    head = np.zeros(4, dtype=np.uint32)
    cells = np.zeros(max_cells, dtype=np.longdouble)
    tp = np.zeros(1, dtype=np.longdouble)
    te = np.zeros(1, dtype=np.longdouble)
    rc = L.qref_slice_import(kind, text, len(text), max_cells, head.ctypes.data,
                             cells.ctypes.data, tp.ctypes.data, te.ctypes.data)
    if rc:
        raise ValueError(f"qref_slice_import rc={rc}")
    n = int(head[0]) ** 2 if kind == 0 else int(head[0])
    return head, cells[:n], tp[0], te[0]
