"""ctypes driver for oracle/_ref/libqref.so -- the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from qunundrum_b200/.

libqref.so is the reference's own hot-path translation units (built by
oracle/Makefile from /root/reference/src, see oracle/ref_capi.cpp for the list
and file:line citations) behind a plain-C handle API.  The built library
travels to the GPU box with the snapshot; /root/reference itself is not needed
at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libqref.so")

# Slice flag constants (reference: src/common.h:226-257).
SLICE_FLAGS_ERROR_BOUND_WARNING = 0x00000001
SLICE_FLAGS_METHOD_SIMPSON = 0x00020000
SLICE_FLAGS_METHOD_RICHARDSON = 0x00080000

METHOD_HEURISTIC_SIGMA = 0
METHOD_OPTIMAL_LOCAL_SIGMA = 1
METHOD_QUICK = 2
TARGET_D = 0
TARGET_R = 1

_lib = None


def build(reference_root: str = "/root/reference") -> bool:
    """Compile _ref/libqref.so if the reference sources are present."""
    if not os.path.isdir(os.path.join(reference_root, "src")):
        return os.path.exists(LIB_PATH)
    subprocess.check_call(
        ["make", "-s", "-j8", "-C", _HERE, f"REF={reference_root}", "ref"])
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `make -C oracle ref` where "
            "/root/reference is mounted")
    L = C.CDLL(LIB_PATH)
    u32, i32, vp, cp, sz = C.c_uint32, C.c_int32, C.c_void_p, C.c_char_p, C.c_size_t
    ldp = C.c_void_p  # long double * (numpy longdouble buffers)
    L.qref_version.restype = cp
    L.qref_deterministic_d_r.argtypes = [u32, cp, cp, sz]
    L.qref_deterministic_d_r.restype = C.c_int
    L.qref_parameters_new.argtypes = [u32, u32, u32, u32, cp, cp]
    L.qref_parameters_new.restype = vp
    L.qref_parameters_free.argtypes = [vp]
    L.qref_parameters_get.argtypes = [vp, C.POINTER(u32)]
    L.qref_diagonal_parameters_new.argtypes = [u32, u32, u32, u32, u32, u32, cp, cp]
    L.qref_diagonal_parameters_new.restype = vp
    L.qref_diagonal_parameters_free.argtypes = [vp]
    L.qref_distribution_slice_compute.argtypes = [
        vp, C.c_int, C.c_int, u32, i32, i32, ldp, ldp, ldp, C.POINTER(u32)]
    L.qref_linear_distribution_slice_compute.argtypes = [
        vp, C.c_int, C.c_int, u32, i32, ldp, ldp, ldp, C.POINTER(u32)]
    L.qref_diagonal_distribution_slice_compute.argtypes = [
        vp, C.c_int, u32, i32, i32, ldp, ldp, ldp, C.POINTER(u32)]
    L.qref_probability_approx.argtypes = [vp, u32, cp, cp, cp, cp, sz]
    L.qref_probability_approx.restype = C.c_int
    L.qref_probability_approx_quick.argtypes = [vp, cp, cp, cp, sz]
    L.qref_linear_probability.argtypes = [vp, C.c_int, cp, cp, sz]
    L.qref_diagonal_probability_f_eta.argtypes = [vp, cp, i32, u32, cp, sz]
    # sampling path (ref_capi_sample.cpp)
    L.qref_dist_new.argtypes = [C.c_int, vp, u32, vp, vp, vp, ldp, ldp]
    L.qref_dist_new.restype = vp
    L.qref_dist_free.argtypes = [vp]
    L.qref_dist_sort.argtypes = [vp]
    L.qref_dist_ptr.argtypes = [vp]
    L.qref_dist_ptr.restype = vp
    L.qref_dist_describe.argtypes = [vp, vp, vp, vp, ldp, ldp]
    L.qref_dist_set_total.argtypes = [vp, C.c_longdouble]
    L.qref_random_new.argtypes = [cp]
    L.qref_random_new.restype = vp
    L.qref_random_free.argtypes = [vp]
    L.qref_random_bytes.argtypes = [vp, vp, u32]
    L.qref_dist_sample_region.argtypes = [vp, vp, u32, vp, vp]
    L.qref_dist_sample_region.restype = u32
    L.qref_dist_sample_alpha.argtypes = [vp, vp, u32, ldp, ldp, vp]
    L.qref_dist_sample_alpha.restype = u32
    L.qref_tau_estimate.argtypes = [vp, vp, u32, u32, ldp, ldp, vp]
    L.qref_sample_k_from_diagonal.argtypes = [
        vp, C.c_longdouble, cp, i32, u32, u32, cp, sz, ldp, cp, sz]
    L.qref_sample_k_from_diagonal.restype = C.c_int
    L.qref_diagonal_probability_h.argtypes = [vp, cp, u32]
    L.qref_diagonal_probability_h.restype = C.c_longdouble
    L.qref_sample_alpha_from_region.argtypes = [C.c_double, C.c_double, u32, vp, cp, sz]
    L.qref_sample_alpha_from_region.restype = C.c_int
    L.qref_sample_j_k.argtypes = [C.c_int, vp, cp, cp, vp, cp, cp, sz]
    L.qref_sample_j_k.restype = C.c_int
    _lib = L
    return L


def deterministic_d_r(m: int) -> tuple[int, int]:
    """parameters_selection_deterministic_d_r (src/parameters_selection.cpp:21)."""
    cap = 4096
    d = C.create_string_buffer(cap)
    r = C.create_string_buffer(cap)
    if lib().qref_deterministic_d_r(m, d, r, cap) != 0:
        raise RuntimeError("buffer too small")
    return int(d.value), int(r.value)


class RefParameters:
    """Parameters (src/parameters.h:33-114) owned by the reference library."""

    def __init__(self, m: int, s: int, d: int, r: int, t: int = 30, l: int = 0):
        self.h = lib().qref_parameters_new(
            m, s, l, t, str(d).encode(), str(r).encode())
        if not self.h:
            raise ValueError("bad d/r")
        out = (C.c_uint32 * 8)()
        lib().qref_parameters_get(self.h, out)
        (self.m, self.l, self.s, self.t, self.min_alpha_d, self.max_alpha_d,
         self.min_alpha_r, self.max_alpha_r) = [int(x) for x in out]
        self.d, self.r = d, r

    def __del__(self):
        if getattr(self, "h", None):
            lib().qref_parameters_free(self.h)
            self.h = None


class RefDiagonalParameters:
    """Diagonal_Parameters (src/diagonal_parameters.h:32-106)."""

    def __init__(self, m: int, sigma: int, s: int, d: int, r: int,
                 eta_bound: int = 0, t: int = 30, l: int = 0):
        self.h = lib().qref_diagonal_parameters_new(
            m, sigma, s, l, eta_bound, t, str(d).encode(), str(r).encode())
        if not self.h:
            raise ValueError("bad d/r")
        self.m, self.sigma, self.s, self.t, self.eta_bound = m, sigma, s, t, eta_bound
        self.d, self.r = d, r

    def __del__(self):
        if getattr(self, "h", None):
            lib().qref_diagonal_parameters_free(self.h)
            self.h = None


class RefSlice:
    def __init__(self, cells, total_probability, total_error, flags):
        self.cells = cells  # np.longdouble
        self.total_probability = total_probability
        self.total_error = total_error
        self.flags = flags


def _ld1():
    return np.zeros(1, dtype=np.longdouble)


def distribution_slice_compute(params: RefParameters, dimension: int,
                               min_log_alpha_d: int, min_log_alpha_r: int,
                               method: int = METHOD_HEURISTIC_SIGMA,
                               richardson: bool = True) -> RefSlice:
    """cells[i_d + dimension * j_r] as the reference stores them."""
    cells = np.zeros(dimension * dimension, dtype=np.longdouble)
    tp, te = _ld1(), _ld1()
    fl = C.c_uint32(0)
    lib().qref_distribution_slice_compute(
        params.h, int(richardson), method, dimension, min_log_alpha_d,
        min_log_alpha_r, cells.ctypes.data, tp.ctypes.data, te.ctypes.data,
        C.byref(fl))
    return RefSlice(cells, tp[0], te[0], fl.value)


def distribution_slice_copy_scale(cells, src_dimension: int, flags: int, dst_dimension: int):
    """distribution_slice_copy_scale (src/distribution_slice.cpp:229-264): (cells, flags)."""
    L = lib()
    L.qref_distribution_slice_copy_scale.restype = C.c_uint32
    L.qref_distribution_slice_copy_scale.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32,
                                                     C.c_uint32, C.c_void_p]
    src = np.ascontiguousarray(cells, dtype=np.longdouble)
    dst = np.zeros(dst_dimension * dst_dimension, dtype=np.longdouble)
    fl = L.qref_distribution_slice_copy_scale(src_dimension, src.ctypes.data, flags, dst_dimension,
                                              dst.ctypes.data)
    return dst, int(fl)


def linear_distribution_slice_compute(params: RefParameters, dimension: int,
                                      min_log_alpha: int, target: int,
                                      richardson: bool = True) -> RefSlice:
    cells = np.zeros(dimension, dtype=np.longdouble)
    tp, te = _ld1(), _ld1()
    fl = C.c_uint32(0)
    lib().qref_linear_distribution_slice_compute(
        params.h, int(richardson), target, dimension, min_log_alpha,
        cells.ctypes.data, tp.ctypes.data, te.ctypes.data, C.byref(fl))
    return RefSlice(cells, tp[0], te[0], fl.value)


def diagonal_distribution_slice_compute(params: RefDiagonalParameters,
                                        dimension: int, min_log_alpha_r: int,
                                        eta: int,
                                        richardson: bool = True) -> RefSlice:
    cells = np.zeros(dimension, dtype=np.longdouble)
    tp, te = _ld1(), _ld1()
    fl = C.c_uint32(0)
    lib().qref_diagonal_distribution_slice_compute(
        params.h, int(richardson), dimension, min_log_alpha_r, eta,
        cells.ctypes.data, tp.ctypes.data, te.ctypes.data, C.byref(fl))
    return RefSlice(cells, tp[0], te[0], fl.value)


_CAP = 256


def probability_approx(params: RefParameters, sigma: int, theta_d: str,
                       theta_r: str) -> tuple[str, str, bool]:
    n = C.create_string_buffer(_CAP)
    e = C.create_string_buffer(_CAP)
    b = lib().qref_probability_approx(
        params.h, sigma, theta_d.encode(), theta_r.encode(), n, e, _CAP)
    return n.value.decode(), e.value.decode(), bool(b)


def probability_approx_quick(params: RefParameters, theta_d: str,
                             theta_r: str) -> str:
    n = C.create_string_buffer(_CAP)
    lib().qref_probability_approx_quick(
        params.h, theta_d.encode(), theta_r.encode(), n, _CAP)
    return n.value.decode()


def linear_probability(params: RefParameters, target: int, theta: str) -> str:
    n = C.create_string_buffer(_CAP)
    lib().qref_linear_probability(params.h, target, theta.encode(), n, _CAP)
    return n.value.decode()


def diagonal_probability_f_eta(params: RefDiagonalParameters, alpha_r: int,
                               eta: int, theta_precision: int = 0) -> str:
    n = C.create_string_buffer(_CAP)
    lib().qref_diagonal_probability_f_eta(
        params.h, str(alpha_r).encode(), eta, theta_precision, n, _CAP)
    return n.value.decode()


def heuristic_sigma(l: int) -> int:
    """sigma = round((l + tau + 4 - 1.6515) / 2), tau = 11, in float32 as the
    reference (src/distribution_slice_compute.cpp:149-158)."""
    f = np.float32
    v = (f(l) + f(11) + f(4) - f(1.6515)) / f(2.0)
    # C round(): half away from zero.
    return int(np.floor(float(v) + 0.5))


# --------------------------------------------------------------------------- #
# Sampling from stored distributions              #
# --------------------------------------------------------------------------- #

class RefRandom:
    """Random_State expanded from a 32-byte seed (src/keccak_random.c:52)."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.h = lib().qref_random_new(seed)

    def bytes(self, n: int) -> bytes:
        out = np.zeros(n, dtype=np.uint8)
        lib().qref_random_bytes(self.h, out.ctypes.data, n)
        return out.tobytes()

    def words(self, n: int) -> np.ndarray:
        """n consecutive 8-byte draws as the uint64 values random_generate_pivot_* see."""
        return np.frombuffer(self.bytes(8 * n), dtype="<u8").copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().qref_random_free(self.h)
            self.h = None


class RefDistribution:
    """Distribution (dims = 2) or Linear_Distribution (dims = 1) owned by the reference."""

    def __init__(self, dims: int, params: RefParameters, dimension, c0, c1, cells, totals=None):
        self.dims, self.params = dims, params
        self.n = len(dimension)
        dim = np.ascontiguousarray(dimension, dtype=np.uint32)
        a = np.ascontiguousarray(c0, dtype=np.int32)
        b = np.ascontiguousarray(c1 if c1 is not None else np.zeros(self.n), dtype=np.int32)
        cl = np.ascontiguousarray(cells, dtype=np.longdouble)
        assert cl.size == int(sum(int(d) ** dims for d in dim))
        t = None if totals is None else np.ascontiguousarray(totals, dtype=np.longdouble)
        self.h = lib().qref_dist_new(dims, params.h, self.n, dim.ctypes.data, a.ctypes.data,
                                     b.ctypes.data, cl.ctypes.data, None if t is None else t.ctypes.data)

    def sort(self):
        lib().qref_dist_sort(self.h)

    def ptr(self) -> int:
        """Distribution * / Linear_Distribution * (the reference's own struct)."""
        return lib().qref_dist_ptr(self.h)

    def set_total(self, total):
        lib().qref_dist_set_total(self.h, np.longdouble(total))

    def describe(self):
        dim = np.zeros(self.n, dtype=np.uint32)
        a = np.zeros(self.n, dtype=np.int32)
        b = np.zeros(self.n, dtype=np.int32)
        t = np.zeros(self.n, dtype=np.longdouble)
        tot = np.zeros(1, dtype=np.longdouble)
        lib().qref_dist_describe(self.h, dim.ctypes.data, a.ctypes.data, b.ctypes.data, t.ctypes.data,
                                 tot.ctypes.data)
        return dim, a, b, t, tot[0]

    def collapse(self, axis: int):
        """linear_distribution_init_collapse_d (axis 0) / _r (axis 1): (coords, vectors, totals)."""
        L = lib()
        L.qref_dist_collapse.restype = C.c_uint32
        L.qref_dist_collapse.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_uint32)]
        md = C.c_uint32(0)
        n = L.qref_dist_collapse(self.h, axis, 0, None, None, None, C.byref(md))
        coords = np.zeros(n, dtype=np.int32)
        vec = np.zeros((n, md.value), dtype=np.longdouble)
        tot = np.zeros(n, dtype=np.longdouble)
        L.qref_dist_collapse(self.h, axis, n, coords.ctypes.data, vec.ctypes.data, tot.ctypes.data,
                             C.byref(md))
        return coords, vec, tot

    def sample_region(self, rng: RefRandom, k: int):
        out = np.zeros((k, 4))
        ok = np.zeros(k, dtype=np.uint8)
        lib().qref_dist_sample_region(self.h, rng.h, k, out.ctypes.data, ok.ctypes.data)
        return out, ok.astype(bool)

    def sample_alpha(self, rng: RefRandom, k: int):
        """alpha / 2^m of k samples (long double), and the success flags."""
        a0 = np.zeros(k, dtype=np.longdouble)
        a1 = np.zeros(k, dtype=np.longdouble)
        ok = np.zeros(k, dtype=np.uint8)
        lib().qref_dist_sample_alpha(self.h, rng.h, k, a0.ctypes.data, a1.ctypes.data, ok.ctypes.data)
        return a0, a1, ok.astype(bool)

    def tau_estimate(self, rng: RefRandom, n: int, count: int):
        t0 = np.zeros(count, dtype=np.longdouble)
        t1 = np.zeros(count, dtype=np.longdouble)
        ok = np.zeros(count, dtype=np.uint8)
        lib().qref_tau_estimate(self.h, rng.h, n, count, t0.ctypes.data, t1.ctypes.data, ok.ctypes.data)
        return t0, t1, ok.astype(bool)

    def __del__(self):
        if getattr(self, "h", None):
            lib().qref_dist_free(self.h)
            self.h = None


def sample_k_from_diagonal_j_eta_pivot(params: RefDiagonalParameters, pivot, j: int, eta: int,
                                       delta_bound: int = 0xffffffff, precision: int = 0):
    """sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412): (ok, k, alpha_phi scaled by
    2^-(m + sigma - l) as a long double, alpha_phi with 60 digits as a string)."""
    kb = C.create_string_buffer(4096)
    ab = C.create_string_buffer(256)
    a = np.zeros(1, dtype=np.longdouble)
    rc = lib().qref_sample_k_from_diagonal(
        params.h, C.c_longdouble(pivot), str(j).encode(), eta, delta_bound, precision, kb, 4096,
        a.ctypes.data_as(C.c_void_p), ab, 256)
    if rc < 0:
        raise RuntimeError("buffer too small")
    return bool(rc), int(kb.value), a[0], ab.value.decode()


def diagonal_probability_h(params: RefDiagonalParameters, phi: str, precision: int = 0):
    """diagonal_probability_approx_h (src/diagonal_probability.cpp:99) as a long double."""
    return np.longdouble(lib().qref_diagonal_probability_h(params.h, phi.encode(), precision))


# --------------------------------------------------------------------------- #
# The exact samplers (src/sample.cpp:78-410)                                  #
# --------------------------------------------------------------------------- #

def sample_alpha_from_region(min_log_alpha: float, max_log_alpha: float, kappa: int,
                             rng: RefRandom) -> int:
    """sample_alpha_from_region (src/sample.cpp:78-158) on the given Random_State."""
    buf = C.create_string_buffer(16384)
    rc = lib().qref_sample_alpha_from_region(min_log_alpha, max_log_alpha, kappa, rng.h, buf, 16384)
    if rc:
        raise RuntimeError("buffer too small")
    return int(buf.value, 16)


def sample_j_k(mode: int, params, alpha_d, alpha_r, rng: RefRandom):
    """mode 0: sample_j_from_alpha_r, 1: sample_j_k_from_alpha_d_r, 2: sample_j_k_from_alpha_d
    (RefParameters); 3: sample_j_from_diagonal_alpha_r (RefDiagonalParameters). Returns (j, k)."""
    jb = C.create_string_buffer(16384)
    kb = C.create_string_buffer(16384)
    rc = lib().qref_sample_j_k(mode, params.h,
                               None if alpha_d is None else format(alpha_d, "x").encode(),
                               None if alpha_r is None else format(alpha_r, "x").encode(),
                               rng.h, jb, kb, 16384)
    if rc:
        raise RuntimeError("buffer too small")
    return int(jb.value, 16), int(kb.value, 16)
