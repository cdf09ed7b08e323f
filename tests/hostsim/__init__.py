"""ctypes driver of tests/hostsim (TEST-ONLY CPU twin of the device functions).

Builds tests/hostsim/_build/libhostsim.so on first use with g++ (seconds).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "_build", "libhostsim.so")
_SRCS = [os.path.join(_HERE, "hostsim.cpp"),
         os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
         os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp")]
_DEPS = _SRCS + [os.path.join(_ROOT, "qunundrum_b200", "csrc", f) for f in
                 ("qmath.cuh", "integrands.cuh", "slice_cells.cuh", "sigma_opt.cuh", "plan.hpp",
                  "hostconst.hpp", "bigint.hpp", "textfmt.cuh", "textparse.cuh", "text_tables.hpp")]
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    stale = force or not os.path.exists(_LIB) or any(
        os.path.getmtime(d) > os.path.getmtime(_LIB) for d in _DEPS)
    if stale:
        subprocess.check_call(
            ["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++",
             *_SRCS, "-o", _LIB])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hostsim_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def be(x: int) -> bytes:
    return x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")


def host_consts(m, l, sigma, d, r):
    out = np.zeros(20)
    db, rb = be(d), be(r)
    rc = lib().hostsim_host_consts(
        C.c_uint32(m), C.c_uint32(l), C.c_uint32(sigma), db, C.c_size_t(len(db)), rb,
        C.c_size_t(len(rb)), out.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"host_consts rc={rc}")
    names = ["kappa", "kappa_q", "c_over_L", "n_over_L", "n1_over_L", "beta_m",
             "rbeta_m", "r_m", "d_m", "rho"]
    return {n: (out[2 * i], out[2 * i + 1]) for i, n in enumerate(names)}


def exp2_table(n):
    out = np.zeros(2 * (n + 1))
    lib().hostsim_exp2_table(C.c_uint32(n), out.ctypes.data_as(C.c_void_p))
    return out.reshape(-1, 2)


def heuristic_sigma(l):
    lib().hostsim_heuristic_sigma.restype = C.c_uint32
    return int(lib().hostsim_heuristic_sigma(C.c_uint32(l)))


def slice2d(m, l, d, r, method, richardson, D, a_d, a_r):
    a_d = np.ascontiguousarray(a_d, dtype=np.int32)
    a_r = np.ascontiguousarray(a_r, dtype=np.int32)
    n = len(a_d)
    cells = np.zeros((n, D * D))
    tp = np.zeros(n, dtype=np.longdouble)
    te = np.zeros(n, dtype=np.longdouble)
    fl = np.zeros(n, dtype=np.uint32)
    db, rb = be(d), be(r)
    rc = lib().hostsim_slice2d(
        C.c_uint32(m), C.c_uint32(l), db, C.c_size_t(len(db)), rb, C.c_size_t(len(rb)),
        C.c_int(method), C.c_int(richardson), C.c_uint32(D), C.c_uint32(n),
        a_d.ctypes.data_as(C.c_void_p), a_r.ctypes.data_as(C.c_void_p),
        cells.ctypes.data_as(C.c_void_p), tp.ctypes.data_as(C.c_void_p),
        te.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(lib().hostsim_last_error().decode())
    return cells, tp, te, fl


def slice1d(m, l, sigma, d, r, kind, richardson, D, a, eta=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    n = len(a)
    eta_arr = np.ascontiguousarray(eta if eta is not None else np.zeros(n), dtype=np.int32)
    cells = np.zeros((n, D))
    tp = np.zeros(n, dtype=np.longdouble)
    fl = np.zeros(n, dtype=np.uint32)
    db, rb = be(d), be(r)
    rc = lib().hostsim_slice1d(
        C.c_uint32(m), C.c_uint32(l), C.c_uint32(sigma), db, C.c_size_t(len(db)), rb,
        C.c_size_t(len(rb)), C.c_int(kind), C.c_int(richardson), C.c_uint32(D),
        C.c_uint32(n), a.ctypes.data_as(C.c_void_p), eta_arr.ctypes.data_as(C.c_void_p),
        cells.ctypes.data_as(C.c_void_p), tp.ctypes.data_as(C.c_void_p),
        fl.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(lib().hostsim_last_error().decode())
    return cells, tp, fl


# ---- text formatter ----------------------------------------------------------------

def pow10_entry(k):
    w = (C.c_uint32 * 6)()
    e2 = C.c_int32()
    ex = C.c_uint32()
    if lib().hostsim_pow10_entry(C.c_int(k), w, C.byref(e2), C.byref(ex)):
        raise ValueError(k)
    T = sum(int(w[i]) << (32 * i) for i in range(6))
    return T, e2.value, bool(ex.value)


def floor_log10_pow2(n):
    lib().hostsim_floor_log10_pow2.restype = C.c_int32
    return int(lib().hostsim_floor_log10_pow2(C.c_int32(n)))


def text_format_ld(values, force_band=False):
    """'%.24Lg\\n' per value through the CPU compile of textfmt.cuh; returns (bytes, n_exact)."""
    v = np.ascontiguousarray(values, dtype=np.longdouble)
    out = C.create_string_buffer(34 * max(1, v.size))
    nx = C.c_uint64()
    lib().hostsim_text_format_ld.restype = C.c_size_t
    n = lib().hostsim_text_format_ld(v.ctypes.data_as(C.c_void_p), C.c_size_t(v.size), out,
                                     C.c_int(1 if force_band else 0), C.byref(nx))
    return out.raw[:n], nx.value


def text_parse_ld(text: bytes, n: int, force_band=False):
    """The first n numbers of text through the CPU compile of textparse.cuh:
    (values, consumed, n_exact); raises ValueError(code) on a parse error."""
    v = np.zeros(max(n, 1), dtype=np.longdouble)
    used = C.c_size_t()
    nx = C.c_uint64()
    rc = lib().hostsim_text_parse_ld(text, C.c_size_t(len(text)), C.c_size_t(n),
                                     v.ctypes.data_as(C.c_void_p), C.byref(used),
                                     C.c_int(1 if force_band else 0), C.byref(nx))
    if rc:
        raise ValueError(rc)
    return v[:n], used.value, nx.value
