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
                  "hostconst.hpp", "bigint.hpp", "textfmt.cuh", "textparse.cuh", "text_tables.hpp",
                  "sampler.cuh", "x87soft.cuh", "diagk.cuh", "diagk_host.hpp", "client_math.cuh", "exact.cuh", "exact_host.hpp")]
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


def so_fast(m, l, d, r, richardson, D, a_d, a_r):
    """total_error, flags and (sigma_0 coarse, fine) of the sigma-optimal method by the closed-form
    walk (sigma_opt.cuh); fallback = True when a point left the range the closed form is proven in."""
    a_d = np.ascontiguousarray(a_d, dtype=np.int32)
    a_r = np.ascontiguousarray(a_r, dtype=np.int32)
    n = len(a_d)
    te = np.zeros(n, dtype=np.longdouble)
    fl = np.zeros(n, dtype=np.uint32)
    s0 = np.zeros((n, 2), dtype=np.int32)
    db, rb = be(d), be(r)
    rc = lib().hostsim_so_fast(
        C.c_uint32(m), C.c_uint32(l), db, C.c_size_t(len(db)), rb, C.c_size_t(len(rb)), C.c_int(richardson),
        C.c_uint32(D), C.c_uint32(n), a_d.ctypes.data_as(C.c_void_p), a_r.ctypes.data_as(C.c_void_p),
        te.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p), s0.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise ValueError(lib().hostsim_last_error().decode())
    return te, fl, s0, bool(rc)


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


# ---- sampler twin (sampler.cuh / x87soft.cuh) -------------------------------------------------

def x87_op(op: int, a, b):
    """RN64 result of a (+, -, *) b through x87soft.cuh, as np.longdouble; None if unsupported."""
    aa = np.array([a], dtype=np.longdouble)
    bb = np.array([b], dtype=np.longdouble)
    out = np.zeros(1, dtype=np.longdouble)
    ok = lib().hostsim_x87_op(op, aa.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p),
                              out.ctypes.data_as(C.c_void_p))
    return out[0] if ok else None


def x87_pivot(w: int):
    out = np.zeros(1, dtype=np.longdouble)
    hi, lo = C.c_double(0), C.c_double(0)
    lib().hostsim_x87_pivot(C.c_uint64(w), out.ctypes.data_as(C.c_void_p), C.byref(hi), C.byref(lo))
    return out[0], hi.value, lo.value


class HostSampler:
    """CPU twin of qb200_sampler: the same __host__ __device__ code, driven by plain loops."""

    def __init__(self, dims, m, dimension, c0, c1, cells, totals, total):
        self.dims = dims
        n = len(dimension)
        dim = np.ascontiguousarray(dimension, dtype=np.uint32)
        a = np.ascontiguousarray(c0, dtype=np.int32)
        b = np.ascontiguousarray(c1 if c1 is not None else np.zeros(n), dtype=np.int32)
        cl = np.ascontiguousarray(cells, dtype=np.longdouble)
        t = np.ascontiguousarray(totals, dtype=np.longdouble)
        tt = np.array([total], dtype=np.longdouble)
        L = lib()
        L.hostsim_sampler_new.restype = C.c_void_p
        self.h = C.c_void_p(L.hostsim_sampler_new(
            dims, C.c_uint32(m), C.c_uint32(n), dim.ctypes.data_as(C.c_void_p),
            a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p),
            t.ctypes.data_as(C.c_void_p), tt.ctypes.data_as(C.c_void_p)))

    def bad(self) -> bool:
        return bool(lib().hostsim_sampler_bad(self.h))

    def sample(self, words, force_exact=False):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        k = w.size // (self.dims + 2)
        out = np.zeros((k, 8))
        st = np.zeros(k, dtype=np.int32)
        ex = np.zeros(k, dtype=np.int32)
        lib().hostsim_sampler_sample(self.h, C.c_uint32(k), w.ctypes.data_as(C.c_void_p),
                                     int(force_exact), out.ctypes.data_as(C.c_void_p),
                                     st.ctypes.data_as(C.c_void_p), ex.ctypes.data_as(C.c_void_p))
        return out, st, ex

    def guide_stats(self, words):
        """(searches, brackets that held, mean bracket width in blocks) for the slice search and for
        the cell search, over the pivots in `words` (two per sample)."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.zeros(6)
        lib().hostsim_sampler_guide_stats(self.h, C.c_uint32(w.size // 2), w.ctypes.data_as(C.c_void_p),
                                          out.ctypes.data_as(C.c_void_p))
        return [(int(out[i]), int(out[i + 1]), out[i + 2] / max(1.0, out[i + 1])) for i in (0, 3)]

    def __del__(self):
        if getattr(self, "h", None):
            lib().hostsim_sampler_free(self.h)
            self.h = None


# ---- diagonal k sampler (diagk.cuh) -----------------------------------------------------

def int_to_limbs(x: int, n: int) -> np.ndarray:
    return np.frombuffer(x.to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def limbs_to_int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


class DiagK:
    """CPU twin of the diagonal k sampler: diagk_sample() of csrc/diagk.cuh in a plain loop."""

    def __init__(self, m, sigma, l, d, r):
        L = lib()
        L.hostsim_diagk_new.restype = C.c_void_p
        db, rb = be(d), be(r)
        self.h = L.hostsim_diagk_new(C.c_uint32(m), C.c_uint32(sigma), C.c_uint32(l), db,
                                     C.c_size_t(len(db)), rb, C.c_size_t(len(rb)))
        if not self.h:
            raise ValueError(L.hostsim_last_error().decode())
        dims = (C.c_uint32 * 3)()
        L.hostsim_diagk_dims(C.c_void_p(self.h), dims)
        self.k, self.wj, self.wl = [int(v) for v in dims]
        self.m, self.sigma, self.l = m, sigma, l

    def __del__(self):
        if getattr(self, "h", None):
            lib().hostsim_diagk_free(C.c_void_p(self.h))
            self.h = None

    def set_force_exact(self, on):
        lib().hostsim_diagk_force_exact(C.c_void_p(self.h), C.c_int(int(bool(on))))

    def sample(self, js, etas, pivots, delta_bound=0xffffffff):
        """-> (k as Python ints, x = alpha_phi / 2^(m+sigma-l) as (hi, lo) rows, delta, status)"""
        n = len(js)
        J = np.zeros((n, self.wj), dtype=np.uint32)
        for i, j in enumerate(js):
            J[i] = int_to_limbs(j, self.wj)
        eta = np.ascontiguousarray(etas, dtype=np.int32)
        piv = np.ascontiguousarray(pivots, dtype=np.longdouble)
        K = np.zeros((n, self.wl), dtype=np.uint32)
        x = np.zeros((n, 2))
        delta = np.zeros(n, dtype=np.int64)
        status = np.zeros(n, dtype=np.int32)
        rc = lib().hostsim_diagk_sample(
            C.c_void_p(self.h), C.c_uint32(n), J.ctypes.data_as(C.c_void_p),
            eta.ctypes.data_as(C.c_void_p), piv.ctypes.data_as(C.c_void_p),
            C.c_uint64(delta_bound), K.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
            delta.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p))
        if rc:
            raise ValueError("bad pivot")
        return [limbs_to_int(K[i]) for i in range(n)], x, delta, status


def sinpi_acc(hi, lo=0.0):
    out = np.zeros(2)
    lib().hostsim_sinpi_acc(C.c_double(hi), C.c_double(lo), out.ctypes.data_as(C.c_void_p))
    return out


def x87_from_dd(hi, lo):
    out = np.zeros(1, dtype=np.longdouble)
    lib().hostsim_x87_from_dd(C.c_double(hi), C.c_double(lo), out.ctypes.data_as(C.c_void_p))
    return out[0]


def diagk_h(l, x):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(len(x), dtype=np.longdouble)
    lib().hostsim_diagk_h(C.c_uint32(l), C.c_uint32(len(x)), x.ctypes.data_as(C.c_void_p),
                          out.ctypes.data_as(C.c_void_p))
    return out


# ---- generator client / server tail (client_math.cuh) -------------------------------------------

def x87_div_u32(a, q: int):
    a = np.array([a], dtype=np.longdouble)
    out = np.zeros(1, dtype=np.longdouble)
    ok = lib().hostsim_x87_div_u32(a.ctypes.data_as(C.c_void_p), C.c_uint32(q), out.ctypes.data_as(C.c_void_p))
    return out[0] if ok else None


def x87_from_double(v: float):
    out = np.zeros(1, dtype=np.longdouble)
    lib().hostsim_x87_from_double(C.c_double(v), out.ctypes.data_as(C.c_void_p))
    return out[0]


def copy_scale(cells, D: int, store: int):
    """distribution_slice_copy_scale as kernels_client.cuh computes it (D x D doubles in)."""
    src = np.ascontiguousarray(cells, dtype=np.float64)
    assert src.size == D * D
    out = np.zeros(store * store, dtype=np.longdouble)
    ok = lib().hostsim_copy_scale(D, store, src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert ok
    return out


def collapse(axis: int, max_dim: int, slices):
    """One destination vector: `slices` = the source slices' cells (long double, D x D) in order."""
    arrs = [np.ascontiguousarray(s, dtype=np.longdouble) for s in slices]
    dims = np.array([int(round(a.size ** 0.5)) for a in arrs], dtype=np.uint32)
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.zeros(max_dim, dtype=np.longdouble)
    ok = lib().hostsim_collapse(axis, C.c_uint32(max_dim), C.c_uint32(len(arrs)), dims.ctypes.data_as(C.c_void_p),
                                ptrs, out.ctypes.data_as(C.c_void_p))
    assert ok
    return out


# ---- exact samplers (exact.cuh) ---------------------------------------------------------------

REGION_DTYPE = np.dtype([("min_log_alpha", "<i4"), ("region", "<u4"), ("dimension", "<u4"),
                         ("length", "<u4"), ("offset", "<u8")])


def _rows(vals, w):
    out = np.zeros((len(vals), w), dtype=np.uint32)
    for i, v in enumerate(vals):
        out[i] = np.frombuffer(int(abs(v)).to_bytes(4 * w, "little"), dtype=np.uint32)
    return out


def _ints(rows):
    return [int.from_bytes(np.ascontiguousarray(r).tobytes(), "little") for r in rows]


class Exact:
    """The exact samplers through the CPU twin; the interface of qunundrum_b200.host.ExactSampler."""

    def __init__(self, kind, m, l, sigma, d, r, dimension_max, emax=0):
        db, rb = be(d), be(r)
        L = lib()
        L.hostsim_exact_new.restype = C.c_void_p
        L.hostsim_exact_table.restype = C.c_void_p
        L.hostsim_exact_inverse.restype = C.c_void_p
        L.hostsim_exact_region_bytes.restype = C.c_uint32
        self.h = C.c_void_p(L.hostsim_exact_new(
            C.c_int(kind), C.c_uint32(m), C.c_uint32(l), C.c_uint32(sigma), db, C.c_size_t(len(db)), rb,
            C.c_size_t(len(rb)), C.c_uint32(dimension_max), C.c_uint32(emax)))
        if not self.h:
            raise ValueError(L.hostsim_last_error().decode())
        out = (C.c_uint32 * 9)()
        L.hostsim_exact_dims(self.h, out)
        (self.wa, self.wn, self.wk, self.kappa_d, self.kappa_r, self.tw, self.P, self.table_dim,
         self.emax) = [int(x) for x in out]

    def __del__(self):
        if getattr(self, "h", None):
            lib().hostsim_exact_free(self.h)
            self.h = None

    def table(self):
        p = lib().hostsim_exact_table(self.h)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.table_dim, self.tw))
        return _ints(a)

    def inverse(self, which):
        p = lib().hostsim_exact_inverse(self.h, C.c_int(which))
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(1, self.wn))
        return _ints(a)[0]

    def region_bytes(self, min_log_alpha, region, dimension):
        st = C.c_int32(0)
        n = lib().hostsim_exact_region_bytes(self.h, C.c_int32(min_log_alpha), C.c_uint32(region),
                                             C.c_uint32(dimension), C.byref(st))
        return int(n), int(st.value)

    def alpha(self, regions, kappa, stream: bytes):
        """regions: (min_log_alpha, region, dimension, offset, length) per sample. Returns the signed
        integers and the status codes."""
        g = np.zeros(len(regions), dtype=REGION_DTYPE)
        for i, (a, reg, dim, off, ln) in enumerate(regions):
            g[i] = (a, reg, dim, ln, off)
        buf = np.frombuffer(stream, dtype=np.uint8)
        n = len(g)
        rows = np.zeros((n, self.wa), dtype=np.uint32)
        neg = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        lib().hostsim_exact_alpha(self.h, C.c_uint32(n), g.ctypes.data_as(C.c_void_p), C.c_uint32(kappa),
                                  buf.ctypes.data_as(C.c_void_p), C.c_uint64(len(buf)),
                                  rows.ctypes.data_as(C.c_void_p), neg.ctypes.data_as(C.c_void_p),
                                  st.ctypes.data_as(C.c_void_p))
        vals = [(-v if s else v) for v, s in zip(_ints(rows), neg)]
        return vals, st

    def _jk(self, mode, alpha_d, alpha_r, t, k_in=None):
        n = len(alpha_d if alpha_d is not None else alpha_r)
        kap = self.kappa_d if mode == 2 else self.kappa_r
        tl = max(1, (kap + 31) // 32)
        if kap and t is None:
            raise ValueError("t is needed when the divisor is even")
        ad = _rows(alpha_d, self.wa) if alpha_d is not None else None
        ar = _rows(alpha_r, self.wa) if alpha_r is not None else None
        nd = np.array([1 if v < 0 else 0 for v in alpha_d], dtype=np.int32) if alpha_d is not None else None
        nr = np.array([1 if v < 0 else 0 for v in alpha_r], dtype=np.int32) if alpha_r is not None else None
        tt = _rows(t, tl) if (t is not None and kap) else None
        j = np.zeros((n, self.wn), dtype=np.uint32)
        k = _rows(k_in, self.wk) if k_in is not None else np.zeros((n, max(1, self.wk)), dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        lib().hostsim_exact_jk(self.h, C.c_int(mode), C.c_uint32(n), p(ad), p(nd), p(ar), p(nr), p(tt), p(j), p(k))
        return _ints(j), _ints(k)

    def j_from_alpha_r(self, alpha_r, t=None):
        return self._jk(0, None, alpha_r, t)[0]

    def j_k_from_alpha_d_r(self, alpha_d, alpha_r, t=None):
        return self._jk(1, alpha_d, alpha_r, t)

    def j_from_alpha_d_k(self, alpha_d, k, t=None):
        return self._jk(2, alpha_d, None, t, k_in=k)[0]
