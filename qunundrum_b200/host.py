"""Host-side mirror of the reference's slice-integration interface (ctypes over the C ABI).

Names follow the reference:

  Parameters / Diagonal_Parameters            src/parameters.h:33-114, src/diagonal_parameters.h:32-106
  Distribution_Slice                          src/distribution_slice.h:86-139
  Linear_Distribution_Slice                   src/linear_distribution_slice.h:48-91
  Diagonal_Distribution_Slice                 src/diagonal_distribution_slice.h:32-83
  distribution_slice_compute[_richardson]     src/distribution_slice.h:364-392
  linear_distribution_slice_compute[_richardson]
  diagonal_distribution_slice_compute[_richardson]

Semantics kept: the caller creates the slice with its dimension, the callee
fills the cells, total_probability, total_error, coordinates (and eta) and
REPLACES the method bits of flags; parameters are read-only; errors are fatal
(the reference calls critical() -> exit(-1), src/errors.c; here CriticalError).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

DISTRIBUTION_SLICE_COMPUTE_METHOD_HEURISTIC_SIGMA = 0
DISTRIBUTION_SLICE_COMPUTE_METHOD_OPTIMAL_LOCAL_SIGMA = 1
DISTRIBUTION_SLICE_COMPUTE_METHOD_QUICK = 2
LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D = 0
LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_R = 1
KIND_DIAGONAL = 2

SLICE_FLAGS_ERROR_BOUND_WARNING = 0x00000001
SLICE_FLAGS_METHOD_SIMPSON = 0x00020000
SLICE_FLAGS_METHOD_RICHARDSON = 0x00080000
SLICE_FLAGS_MASK_METHOD = 0x000F0000

SUMMARY_STRIDE = 8


class CriticalError(RuntimeError):
    """The reference's critical(): unrecoverable error in a slice computation."""


def lib_path() -> str:
    # QB200_LIB: development override (A/B-testing kernel builds); the default is the in-tree build
    return os.environ.get("QB200_LIB") or os.path.join(_HERE, "libqunundrum_b200.so")


class _Params(C.Structure):
    _fields_ = [("m", C.c_uint32), ("l", C.c_uint32), ("sigma", C.c_uint32),
                ("d_be", C.c_char_p), ("d_len", C.c_size_t),
                ("r_be", C.c_char_p), ("r_len", C.c_size_t)]


_lib = None


def lib():
    """Load libqunundrum_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise CriticalError(
            f"{path} is missing: build it with `python -m qunundrum_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u32, i32p = C.c_void_p, C.c_uint32, C.c_void_p
    PP = C.POINTER(_Params)
    L.qb200_version.restype = C.c_int
    L.qb200_last_error.restype = C.c_char_p
    L.qb200_device_count.restype = C.c_int
    L.qb200_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.qb200_destroy.argtypes = [vp]
    L.qb200_launch_count.argtypes = [vp]
    L.qb200_launch_count.restype = C.c_uint64
    L.qb200_host_alloc.argtypes = [C.c_size_t]
    L.qb200_host_alloc.restype = vp
    L.qb200_host_free.argtypes = [vp]
    L.qb200_slice2d_compute.argtypes = [vp, PP, C.c_int, C.c_int, u32, u32, i32p, i32p,
                                        vp, vp, vp, vp]
    L.qb200_slice1d_compute.argtypes = [vp, PP, C.c_int, C.c_int, u32, u32, i32p, i32p,
                                        vp, vp, vp]
    L.qb200_plan2d_create.argtypes = [vp, PP, C.c_int, C.c_int, u32, u32, i32p, i32p,
                                      C.POINTER(vp)]
    L.qb200_plan1d_create.argtypes = [vp, PP, C.c_int, C.c_int, u32, u32, i32p, i32p,
                                      C.POINTER(vp)]
    L.qb200_plan_destroy.argtypes = [vp]
    L.qb200_plan_cells.argtypes = [vp]
    L.qb200_plan_cells.restype = C.c_uint64
    L.qb200_plan_launches.argtypes = [vp]
    L.qb200_plan_launches.restype = C.c_uint32
    L.qb200_plan_set_algorithm.argtypes = [vp, C.c_int]
    L.qb200_plan_algorithm.argtypes = [vp]
    L.qb200_plan_run.argtypes = [vp, vp, vp, vp]
    L.qb200_plan_finish.argtypes = [vp, vp, vp, vp, vp]
    L.qb200_heuristic_sigma.argtypes = [u32]
    L.qb200_heuristic_sigma.restype = u32
    L.qb200_host_constants.argtypes = [PP, vp]
    L.qb200_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.qb200_text_bound.argtypes = [C.c_size_t]
    L.qb200_text_bound.restype = C.c_size_t
    L.qb200_text_format_ld.argtypes = [vp, vp, C.c_size_t, vp, C.POINTER(vp),
                                       C.POINTER(C.c_size_t)]
    L.qb200_text_format_f64.argtypes = [vp, vp, C.c_size_t, vp, C.POINTER(vp),
                                        C.POINTER(C.c_size_t)]
    L.qb200_text_format_device.argtypes = [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp, vp]
    L.qb200_text_parse_ld.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t, vp,
                                      C.POINTER(C.c_size_t)]
    L.qb200_text_parse_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.qb200_text_pow10.argtypes = [C.c_int, vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    L.qb200_text_set_force_exact.argtypes = [vp, C.c_int]
    L.qb200_text_exact_count.argtypes = [vp]
    L.qb200_text_exact_count.restype = C.c_uint64
    u64p = C.c_void_p
    L.qb200_sampler_create.argtypes = [vp, C.c_int, u32, u32, vp, vp, vp, vp, vp, C.c_longdouble,
                                       C.POINTER(vp)]
    L.qb200_sampler_destroy.argtypes = [vp]
    L.qb200_sampler_words_per_sample.argtypes = [vp]
    L.qb200_sampler_words_per_sample.restype = u32
    L.qb200_sampler_cells.argtypes = [vp]
    L.qb200_sampler_cells.restype = C.c_uint64
    L.qb200_sampler_sample.argtypes = [vp, u32, u64p, vp, vp, vp, vp, vp]
    L.qb200_sampler_tau_estimate.argtypes = [vp, u32, u32, u64p, C.c_size_t, C.POINTER(C.c_size_t),
                                             C.POINTER(u32), vp, vp, vp]
    L.qb200_sampler_tau_device.argtypes = [vp, u32, u32, vp, vp, vp, vp]
    L.qb200_sampler_set_force_exact.argtypes = [vp, C.c_int]
    L.qb200_sampler_exact_count.argtypes = [vp]
    L.qb200_sampler_exact_count.restype = C.c_uint64
    L.qb200_sampler_first_failing_word.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.qb200_diagk_create.argtypes = [vp, PP, C.POINTER(vp)]
    L.qb200_diagk_destroy.argtypes = [vp]
    L.qb200_diagk_j_limbs.argtypes = [vp]
    L.qb200_diagk_j_limbs.restype = u32
    L.qb200_diagk_k_limbs.argtypes = [vp]
    L.qb200_diagk_k_limbs.restype = u32
    L.qb200_diagk_sample.argtypes = [vp, u32, vp, vp, vp, u32, vp, vp, vp, vp, vp]
    L.qb200_diagk_sample_device.argtypes = [vp, u32, vp, vp, vp, u32, vp, vp, vp]
    L.qb200_diagk_tau_estimate.argtypes = [vp, u32, u32, vp, vp, vp, u32, u32, vp, vp]
    L.qb200_diagk_h.argtypes = [vp, u32, vp, vp, vp]
    L.qb200_diagk_set_force_exact.argtypes = [vp, C.c_int]
    L.qb200_exact_create.argtypes = [vp, PP, C.c_int, u32, u32, C.POINTER(vp)]
    L.qb200_exact_destroy.argtypes = [vp]
    L.qb200_exact_dims.argtypes = [vp, vp]
    L.qb200_exact_dims.restype = None
    L.qb200_exact_kernel_ms.argtypes = [vp, vp]
    L.qb200_exact_region_bytes.argtypes = [vp, C.c_int32, u32, u32, C.POINTER(u32)]
    L.qb200_exact_alpha.argtypes = [vp, u32, vp, u32, vp, C.c_uint64, vp, vp, vp]
    L.qb200_exact_j_k.argtypes = [vp, C.c_int, u32, vp, vp, vp, vp, vp, vp, vp]
    L.qb200_diagk_sample_drawn.argtypes = [vp, vp, u32, vp, vp, vp, C.c_uint64, vp, vp, u32, vp, vp, vp,
                                           vp, vp, vp]
    L.qb200_slice2d_compute_scaled.argtypes = [vp, PP, C.c_int, C.c_int, u32, u32, u32, i32p, i32p,
                                               vp, vp, vp, vp]
    L.qb200_resident_create.argtypes = [vp, u32, vp, vp, vp, C.POINTER(vp)]
    L.qb200_resident_destroy.argtypes = [vp]
    L.qb200_resident_cells.argtypes = [vp]
    L.qb200_resident_cells.restype = C.c_uint64
    L.qb200_resident_collapse2d.argtypes = [vp, C.c_int, vp, u32, vp, vp, u32, vp]
    L.qb200_resident_format.argtypes = [vp, u32, u32, C.POINTER(vp), vp, vp]
    L.qb200_resident_format_prefetch.argtypes = [vp, u32, u32]
    _lib = L
    return L


def _err() -> str:
    return lib().qb200_last_error().decode(errors="replace")


def _check(rc: int, what: str):
    if rc != 0:
        raise CriticalError(f"{what}: {_err()} (code {rc})")


def _be(x: int) -> bytes:
    if x < 0:
        raise CriticalError("d and r must be non-negative")
    return x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")


# --------------------------------------------------------------------------- #
# Parameters                                                                  #
# --------------------------------------------------------------------------- #

@dataclass
class Parameters:
    """Parameters (src/parameters.h:33-114); regions as parameters_setup_regions
    (src/parameters.cpp:30-51). Give s (l = ceil(m / s), parameters_explicit_m_s)
    or l (parameters_explicit_m_l, s = 0)."""
    m: int
    s: int
    d: int
    r: int
    t: int = 30
    l: int = 0

    def __post_init__(self):
        if self.l == 0:
            if self.s <= 0:
                raise CriticalError("Parameters: s or l must be given")
            self.l = int(math.ceil(self.m / self.s))
        else:
            self.s = 0
        self.min_alpha_d = 0 if self.t > self.m else self.m - self.t
        self.max_alpha_d = (self.m + self.l - 2 if self.t >= self.l
                            else self.m + self.t - 1)
        self.min_alpha_r = self.min_alpha_d
        self.max_alpha_r = self.max_alpha_d

    def _c(self, sigma: int = 0):
        db, rb = _be(self.d), _be(self.r)
        p = _Params(self.m, self.l, sigma, db, len(db), rb, len(rb))
        p._keep = (db, rb)
        return p


@dataclass
class Diagonal_Parameters:
    """Diagonal_Parameters (src/diagonal_parameters.h:32-106)."""
    m: int
    sigma: int
    s: int
    d: int
    r: int
    eta_bound: int = 0
    t: int = 30
    l: int = 0

    def __post_init__(self):
        if self.l == 0:
            if self.s <= 0:
                raise CriticalError("Diagonal_Parameters: s or l must be given")
            self.l = int(math.ceil(self.m / self.s))
        else:
            self.s = 0
        self.min_alpha_r = 0 if self.t > self.m else self.m - self.t
        self.max_alpha_r = (self.m + self.sigma - 2 if self.t >= self.sigma
                            else self.m + self.t - 1)

    def _c(self):
        db, rb = _be(self.d), _be(self.r)
        p = _Params(self.m, self.l, self.sigma, db, len(db), rb, len(rb))
        p._keep = (db, rb)
        return p


# --------------------------------------------------------------------------- #
# Slices                                                                      #
# --------------------------------------------------------------------------- #

@dataclass
class Distribution_Slice:
    """Distribution_Slice (src/distribution_slice.h:86-139);
    norm_matrix[i_d + dimension * j_r]."""
    dimension: int
    min_log_alpha_d: int = 0
    min_log_alpha_r: int = 0
    total_probability: np.longdouble = np.longdouble(0)
    total_error: np.longdouble = np.longdouble(0)
    flags: int = 0
    norm_matrix: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.norm_matrix is None:
            self.norm_matrix = np.zeros(self.dimension * self.dimension, dtype=np.longdouble)


@dataclass
class Linear_Distribution_Slice:
    """Linear_Distribution_Slice (src/linear_distribution_slice.h:48-91)."""
    dimension: int
    min_log_alpha: int = 0
    total_probability: np.longdouble = np.longdouble(0)
    total_error: np.longdouble = np.longdouble(0)
    flags: int = 0
    norm_vector: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.norm_vector is None:
            self.norm_vector = np.zeros(self.dimension, dtype=np.longdouble)


@dataclass
class Diagonal_Distribution_Slice:
    """Diagonal_Distribution_Slice (src/diagonal_distribution_slice.h:32-83)."""
    dimension: int
    min_log_alpha_r: int = 0
    eta: int = 0
    total_probability: np.longdouble = np.longdouble(0)
    total_error: np.longdouble = np.longdouble(0)
    flags: int = 0
    norm_vector: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.norm_vector is None:
            self.norm_vector = np.zeros(self.dimension, dtype=np.longdouble)


# --------------------------------------------------------------------------- #
# Context and batch API                                                       #
# --------------------------------------------------------------------------- #

def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Plan:
    """A batch resident on the device (qb200_plan)."""

    def __init__(self, ctx: "Context", handle, n: int, dimension: int, two_d: bool):
        self.ctx, self.h, self.n, self.dimension, self.two_d = ctx, handle, n, dimension, two_d

    @property
    def cells(self) -> int:
        return int(lib().qb200_plan_cells(self.h))

    @property
    def launches(self) -> int:
        return int(lib().qb200_plan_launches(self.h))

    @property
    def algorithm(self) -> int:
        return int(lib().qb200_plan_algorithm(self.h))

    def set_algorithm(self, algo: int):
        _check(lib().qb200_plan_set_algorithm(self.h, algo), "qb200_plan_set_algorithm")

    def run(self, d_cells_ptr: int, d_summary_ptr: int, stream: int = 0):
        """Enqueue one pass over the batch; pointers are device addresses."""
        _check(lib().qb200_plan_run(self.h, C.c_void_p(stream), C.c_void_p(d_cells_ptr),
                                    C.c_void_p(d_summary_ptr)), "qb200_plan_run")

    def finish(self, h_summary: np.ndarray):
        h_summary = np.ascontiguousarray(h_summary, dtype=np.float64)
        tp = np.zeros(self.n, dtype=np.longdouble)
        te = np.zeros(self.n, dtype=np.longdouble)
        fl = np.zeros(self.n, dtype=np.uint32)
        _check(lib().qb200_plan_finish(self.h, h_summary.ctypes.data, tp.ctypes.data,
                                       te.ctypes.data, fl.ctypes.data), "qb200_plan_finish")
        return tp, te, fl

    def close(self):
        if self.h:
            lib().qb200_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU (qb200_context). In an MPI farm: one per worker rank,
    device = (rank - 1) mod device_count."""

    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        _check(lib().qb200_create(device, C.byref(self.h)), "qb200_create")
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().qb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(lib().qb200_launch_count(self.h))

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0)
        _check(lib().qb200_measure_fp64_peak(self.h, C.byref(v)), "qb200_measure_fp64_peak")
        return v.value

    # ---- synchronous batches (host buffers) ---------------------------------
    def slice2d_batch(self, params: Parameters, method: int, richardson: bool,
                      dimension: int, min_log_alpha_d, min_log_alpha_r, out=None):
        a_d, a_r = _i32(min_log_alpha_d), _i32(min_log_alpha_r)
        n = len(a_d)
        if len(a_r) != n:
            raise CriticalError("coordinate arrays differ in length")
        cells = out if out is not None else np.empty((n, dimension * dimension))
        tp = np.zeros(n, dtype=np.longdouble)
        te = np.zeros(n, dtype=np.longdouble)
        fl = np.zeros(n, dtype=np.uint32)
        p = params._c()
        _check(lib().qb200_slice2d_compute(
            self.h, C.byref(p), method, int(bool(richardson)), dimension, n,
            a_d.ctypes.data, a_r.ctypes.data, cells.ctypes.data, tp.ctypes.data,
            te.ctypes.data, fl.ctypes.data), "qb200_slice2d_compute")
        return cells, tp, te, fl

    def slice2d_batch_scaled(self, params: Parameters, method: int, richardson: bool, dimension: int,
                             store_dimension: int, min_log_alpha_d, min_log_alpha_r):
        """Slices computed at `dimension` and scaled to `store_dimension` on the device as
        distribution_slice_copy_scale does (src/distribution_slice.cpp:230-264): long double cells."""
        a_d, a_r = _i32(min_log_alpha_d), _i32(min_log_alpha_r)
        n = len(a_d)
        if len(a_r) != n:
            raise CriticalError("coordinate arrays differ in length")
        cells = np.zeros((n, store_dimension * store_dimension), dtype=np.longdouble)
        tp = np.zeros(n, dtype=np.longdouble)
        te = np.zeros(n, dtype=np.longdouble)
        fl = np.zeros(n, dtype=np.uint32)
        p = params._c()
        _check(lib().qb200_slice2d_compute_scaled(
            self.h, C.byref(p), method, int(bool(richardson)), dimension, store_dimension, n,
            a_d.ctypes.data, a_r.ctypes.data, cells.ctypes.data, tp.ctypes.data, te.ctypes.data,
            fl.ctypes.data), "qb200_slice2d_compute_scaled")
        return cells, tp, te, fl

    def slice1d_batch(self, params, kind: int, richardson: bool, dimension: int,
                      min_log_alpha, eta=None, out=None):
        a = _i32(min_log_alpha)
        n = len(a)
        e = _i32(eta) if eta is not None else None
        cells = out if out is not None else np.empty((n, dimension))
        tp = np.zeros(n, dtype=np.longdouble)
        fl = np.zeros(n, dtype=np.uint32)
        p = params._c()
        _check(lib().qb200_slice1d_compute(
            self.h, C.byref(p), kind, int(bool(richardson)), dimension, n, a.ctypes.data,
            e.ctypes.data if e is not None else None, cells.ctypes.data, tp.ctypes.data,
            fl.ctypes.data), "qb200_slice1d_compute")
        return cells, tp, fl

    # ---- text export ----------------------------------------------------------
    def text_format(self, values, tail=None) -> bytes:
        """One '%.24Lg\\n' line per value (then `tail`), as the reference's slice
        exporters print them. values: long double (x87) or float64 array."""
        v = np.ascontiguousarray(values)
        if v.dtype == np.float64:
            fn, dt = lib().qb200_text_format_f64, np.float64
        else:
            v = np.ascontiguousarray(v, dtype=np.longdouble)
            fn, dt = lib().qb200_text_format_ld, np.longdouble
        t = None if tail is None else np.array([tail], dtype=dt)
        text, n = C.c_void_p(), C.c_size_t()
        _check(fn(self.h, v.ctypes.data, v.size, t.ctypes.data if t is not None else None,
                  C.byref(text), C.byref(n)), "qb200_text_format")
        return C.string_at(text, n.value) if n.value else b""

    def text_format_view(self, values, tail=None):
        """As text_format, without copying the text out of the context's pinned buffer:
        (address, length), valid until the next text call on this context."""
        v = np.ascontiguousarray(values)
        if v.dtype == np.float64:
            fn, dt = lib().qb200_text_format_f64, np.float64
        else:
            v = np.ascontiguousarray(v, dtype=np.longdouble)
            fn, dt = lib().qb200_text_format_ld, np.longdouble
        t = None if tail is None else np.array([tail], dtype=dt)
        text, n = C.c_void_p(), C.c_size_t()
        _check(fn(self.h, v.ctypes.data, v.size, t.ctypes.data if t is not None else None,
                  C.byref(text), C.byref(n)), "qb200_text_format")
        return text.value, n.value

    def text_format_device(self, kind: int, d_values_ptr: int, n: int, d_text_ptr: int,
                           cap: int, d_len_ptr: int, stream: int = 0):
        """Enqueue one formatting launch on device-resident values."""
        _check(lib().qb200_text_format_device(self.h, kind, d_values_ptr, n, d_text_ptr, cap,
                                              d_len_ptr, stream), "qb200_text_format_device")

    def text_parse(self, text: bytes, n: int):
        """The first n white-space separated '%Lg' numbers of text as long doubles,
        and the offset where the reference's FILE position would be afterwards."""
        v = np.zeros(max(n, 1), dtype=np.longdouble)
        used = C.c_size_t()
        _check(lib().qb200_text_parse_ld(self.h, text, len(text), n, v.ctypes.data,
                                         C.byref(used)), "qb200_text_parse_ld")
        return v[:n], used.value

    def text_parse_device(self, d_text_ptr: int, length: int, n: int, d_values_ptr: int,
                          d_info_ptr: int, stream: int = 0):
        _check(lib().qb200_text_parse_device(self.h, d_text_ptr, length, n, d_values_ptr,
                                             d_info_ptr, stream), "qb200_text_parse_device")

    def text_set_force_exact(self, on: bool):
        _check(lib().qb200_text_set_force_exact(self.h, int(on)), "qb200_text_set_force_exact")

    @property
    def text_exact_count(self) -> int:
        return int(lib().qb200_text_exact_count(self.h))

    # ---- device-resident plans ------------------------------------------------
    def plan2d(self, params: Parameters, method: int, richardson: bool, dimension: int,
               min_log_alpha_d, min_log_alpha_r) -> Plan:
        a_d, a_r = _i32(min_log_alpha_d), _i32(min_log_alpha_r)
        h = C.c_void_p()
        p = params._c()
        _check(lib().qb200_plan2d_create(
            self.h, C.byref(p), method, int(bool(richardson)), dimension, len(a_d),
            a_d.ctypes.data, a_r.ctypes.data, C.byref(h)), "qb200_plan2d_create")
        return Plan(self, h, len(a_d), dimension, True)

    def plan1d(self, params, kind: int, richardson: bool, dimension: int, min_log_alpha,
               eta=None) -> Plan:
        a = _i32(min_log_alpha)
        e = _i32(eta) if eta is not None else None
        h = C.c_void_p()
        p = params._c()
        _check(lib().qb200_plan1d_create(
            self.h, C.byref(p), kind, int(bool(richardson)), dimension, len(a),
            a.ctypes.data, e.ctypes.data if e is not None else None, C.byref(h)),
            "qb200_plan1d_create")
        return Plan(self, h, len(a), dimension, False)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def heuristic_sigma(l: int) -> int:
    return int(lib().qb200_heuristic_sigma(l))


def host_constants(m: int, l: int, sigma: int, d: int, r: int) -> dict:
    db, rb = _be(d), _be(r)
    p = _Params(m, l, sigma, db, len(db), rb, len(rb))
    out = np.zeros(20)
    _check(lib().qb200_host_constants(C.byref(p), out.ctypes.data), "qb200_host_constants")
    names = ["kappa", "kappa_q", "c_over_L", "n_over_L", "n1_over_L", "beta_m", "rbeta_m",
             "r_m", "d_m", "rho"]
    return {k: (out[2 * i], out[2 * i + 1]) for i, k in enumerate(names)}


# --------------------------------------------------------------------------- #
# The reference's six entry points                                            #
# --------------------------------------------------------------------------- #

def _apply_flags(old: int, new_bits: int) -> int:
    # slice->flags &= ~MASK_METHOD; |= SIMPSON [| RICHARDSON] [| WARNING]
    # (src/distribution_slice_compute.cpp:413-418, ..._richardson.cpp:69)
    return (old & ~SLICE_FLAGS_MASK_METHOD) | int(new_bits)


def _compute_2d(slice_, parameters, method, a_d, a_r, richardson, ctx):
    ctx = ctx or default_context()
    cells, tp, te, fl = ctx.slice2d_batch(parameters, int(method), richardson,
                                          slice_.dimension, [a_d], [a_r])
    slice_.norm_matrix[:] = cells[0].astype(np.longdouble)
    slice_.total_probability = tp[0]
    slice_.total_error = te[0]
    slice_.min_log_alpha_d = int(a_d)
    slice_.min_log_alpha_r = int(a_r)
    slice_.flags = _apply_flags(slice_.flags, fl[0])


def distribution_slice_compute(slice, parameters, method, min_log_alpha_d,
                               min_log_alpha_r, ctx=None):
    """src/distribution_slice_compute.cpp:38."""
    _compute_2d(slice, parameters, method, min_log_alpha_d, min_log_alpha_r, False, ctx)


def distribution_slice_compute_richardson(slice, parameters, method, min_log_alpha_d,
                                          min_log_alpha_r, ctx=None):
    """src/distribution_slice_compute_richardson.cpp:17."""
    _compute_2d(slice, parameters, method, min_log_alpha_d, min_log_alpha_r, True, ctx)


def _compute_linear(slice_, parameters, target, a, richardson, ctx):
    if target not in (LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D,
                      LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_R):
        raise CriticalError("linear_distribution_slice_compute(): Unknown target.")
    ctx = ctx or default_context()
    cells, tp, fl = ctx.slice1d_batch(parameters, int(target), richardson,
                                      slice_.dimension, [a])
    slice_.norm_vector[:] = cells[0].astype(np.longdouble)
    slice_.total_probability = tp[0]
    slice_.total_error = np.longdouble(0)
    slice_.min_log_alpha = int(a)
    slice_.flags = _apply_flags(slice_.flags, fl[0])


def linear_distribution_slice_compute(slice, parameters, target, min_log_alpha, ctx=None):
    """src/linear_distribution_slice_compute.cpp:30."""
    _compute_linear(slice, parameters, target, min_log_alpha, False, ctx)


def linear_distribution_slice_compute_richardson(slice, parameters, target, min_log_alpha,
                                                 ctx=None):
    """src/linear_distribution_slice_compute_richardson.cpp:17."""
    _compute_linear(slice, parameters, target, min_log_alpha, True, ctx)


def _compute_diagonal(slice_, parameters, a_r, eta, richardson, ctx):
    ctx = ctx or default_context()
    cells, tp, fl = ctx.slice1d_batch(parameters, KIND_DIAGONAL, richardson,
                                      slice_.dimension, [a_r], [eta])
    slice_.norm_vector[:] = cells[0].astype(np.longdouble)
    slice_.total_probability = tp[0]
    slice_.total_error = np.longdouble(0)
    slice_.min_log_alpha_r = int(a_r)
    slice_.eta = int(eta)
    slice_.flags = _apply_flags(slice_.flags, fl[0])


def diagonal_distribution_slice_compute(slice, parameters, min_log_alpha_r, eta, ctx=None):
    """src/diagonal_distribution_slice_compute.cpp:30."""
    _compute_diagonal(slice, parameters, min_log_alpha_r, eta, False, ctx)


def diagonal_distribution_slice_compute_richardson(slice, parameters, min_log_alpha_r, eta,
                                                   ctx=None):
    """src/diagonal_distribution_slice_compute_richardson.cpp:17."""
    _compute_diagonal(slice, parameters, min_log_alpha_r, eta, True, ctx)


# --------------------------------------------------------------------------- #
# The reference's slice exporters                                             #
# --------------------------------------------------------------------------- #

TEXT_X87, TEXT_F64 = 0, 1


def text_pow10(k: int):
    """(T, e2, exact): 10^k ~ T * 2^(e2 - 191), the table the text kernels use (host logic)."""
    w = (C.c_uint32 * 6)()
    e2, ex = C.c_int32(), C.c_uint32()
    _check(lib().qb200_text_pow10(k, w, C.byref(e2), C.byref(ex)), "qb200_text_pow10")
    return sum(int(w[i]) << (32 * i) for i in range(6)), e2.value, bool(ex.value)


def distribution_slice_export(slice, file, ctx=None):
    """src/distribution_slice_import_export.cpp:89-103; file: a binary file object."""
    ctx = ctx or default_context()
    file.write(b"%u\n%d\n%d\n%.8x\n" % (slice.dimension, slice.min_log_alpha_d,
                                       slice.min_log_alpha_r, slice.flags))
    file.write(ctx.text_format(slice.norm_matrix, slice.total_error))


def linear_distribution_slice_export(slice, file, ctx=None):
    """src/linear_distribution_slice_import_export.cpp:82-97."""
    ctx = ctx or default_context()
    file.write(b"%u\n%d\n%.8x\n" % (slice.dimension, slice.min_log_alpha, slice.flags))
    file.write(ctx.text_format(slice.norm_vector, slice.total_error))


def diagonal_distribution_slice_export(slice, file, ctx=None):
    """src/diagonal_distribution_slice_import_export.cpp:87-103."""
    ctx = ctx or default_context()
    file.write(b"%u\n%d\n%d\n%.8x\n" % (slice.dimension, slice.min_log_alpha_r, slice.eta,
                                       slice.flags))
    file.write(ctx.text_format(slice.norm_vector, slice.total_error))


# --------------------------------------------------------------------------- #
# The reference's slice importers                                             #
# --------------------------------------------------------------------------- #

def _import_numbers(file, n: int, ctx):
    """n numbers from the current position of a seekable binary file; leaves the
    position where fscanf("%Lg\\n") x n would (after the white space that follows)."""
    pos = file.tell()
    want = 40 * n + 64
    malformed_attempts = 0
    while True:
        chunk = file.read(want)
        try:
            values, used = ctx.text_parse(chunk, n)
            # a block that is not the rest of the file may end inside the n-th number: accept
            # only if something follows it in the block, or the block reached the end of file
            if used < len(chunk) or len(chunk) < want:
                break
        except CriticalError as e:
            # too few numbers, or a malformed stump of a number cut by the block end ("1.5e-"):
            # read more (a stump is cured by the next, twice as large block); a file that really
            # is malformed fails after two more attempts or when the block reaches its end
            if "(code -20)" not in str(e):
                malformed_attempts += 1
            if len(chunk) < want or malformed_attempts > 2:
                raise
        file.seek(pos)
        want *= 2
    file.seek(pos + used)
    return values


def _import_common(file, n_head: int):
    head = []
    for i in range(n_head):
        line = file.readline()
        if not line.strip():
            raise CriticalError("slice import: failed to import a header field")
        head.append(int(line, 16) if i == n_head - 1 else int(line))
    return head


def distribution_slice_import(file, ctx=None) -> Distribution_Slice:
    """distribution_slice_init_import (src/distribution_slice_import_export.cpp:72-87):
    dimension, min_log_alpha_d, min_log_alpha_r, flags, dimension^2 cells, total_error;
    total_probability is the running long double sum of the cells (:38-46)."""
    ctx = ctx or default_context()
    dimension, a_d, a_r, flags = _import_common(file, 4)
    v = _import_numbers(file, dimension * dimension + 1, ctx)
    s = Distribution_Slice(dimension, a_d, a_r, flags=flags, norm_matrix=v[:-1].copy(),
                           total_error=v[-1])
    s.total_probability = _sequential_sum(s.norm_matrix)
    return s


def linear_distribution_slice_import(file, ctx=None) -> Linear_Distribution_Slice:
    """linear_distribution_slice_init_import (src/linear_distribution_slice_import_export.cpp:67-80)."""
    ctx = ctx or default_context()
    dimension, a, flags = _import_common(file, 3)
    v = _import_numbers(file, dimension + 1, ctx)
    s = Linear_Distribution_Slice(dimension, a, flags=flags, norm_vector=v[:-1].copy(),
                                  total_error=v[-1])
    s.total_probability = _sequential_sum(s.norm_vector)
    return s


def diagonal_distribution_slice_import(file, ctx=None) -> Diagonal_Distribution_Slice:
    """diagonal_distribution_slice_init_import (src/diagonal_distribution_slice_import_export.cpp:72-85)."""
    ctx = ctx or default_context()
    dimension, a_r, eta, flags = _import_common(file, 4)
    v = _import_numbers(file, dimension + 1, ctx)
    s = Diagonal_Distribution_Slice(dimension, a_r, eta, flags=flags, norm_vector=v[:-1].copy(),
                                    total_error=v[-1])
    s.total_probability = _sequential_sum(s.norm_vector)
    return s


def _sequential_sum(v) -> np.longdouble:
    # the reference accumulates in index order in long double; np.cumsum does exactly that
    return np.cumsum(np.asarray(v, dtype=np.longdouble))[-1] if len(v) else np.longdouble(0)


# --------------------------------------------------------------------------- #
# Sampling from stored distributions: tau estimation      #
# --------------------------------------------------------------------------- #

@dataclass
class Distribution:
    """Distribution (src/distribution.h:55-85): the slices in walk order, the parameters' m and
    the running totals distribution_insert_slice keeps (src/distribution.cpp:159-176)."""
    m: int
    slices: list = field(default_factory=list)
    total_probability: np.longdouble = np.longdouble(0)
    total_error: np.longdouble = np.longdouble(0)

    def insert_slice(self, slice_):
        self.slices.append(slice_)
        self.total_probability = np.longdouble(self.total_probability + slice_.total_probability)
        self.total_error = np.longdouble(self.total_error + slice_.total_error)

    def sort_slices(self):
        """distribution_sort_slices (src/distribution.cpp:258-270): descending total probability.
        (qsort leaves the order of equal keys unspecified; this one is stable.)"""
        self.slices.sort(key=lambda s: -s.total_probability)


class Resident:
    """The cells of a stored two-dimensional distribution in device memory (qb200_resident):
    uploaded once, collapsed to both marginals and exported from there."""

    def __init__(self, slices, ctx: "Context" = None):
        self.ctx = ctx or default_context()
        self.slices = list(slices)
        self._keep = [np.ascontiguousarray(s.norm_matrix, dtype=np.longdouble) for s in self.slices]
        n = len(self.slices)
        n_cells = np.array([a.size for a in self._keep], dtype=np.uint64)
        ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in self._keep])
        tails = np.array([s.total_error for s in self.slices], dtype=np.longdouble)
        self.h = C.c_void_p()
        _check(lib().qb200_resident_create(self.ctx.h, n, n_cells.ctypes.data, ptrs, tails.ctypes.data,
                                           C.byref(self.h)), "qb200_resident_create")

    def close(self):
        if getattr(self, "h", None):
            lib().qb200_resident_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def collapse(self, axis: int):
        """(coordinates, vectors [n_dst, max_dimension], totals): the destination slices in the order
        linear_distribution_init_collapse_d (axis 0) / _r (axis 1) creates them."""
        dims = np.array([s.dimension for s in self.slices], dtype=np.uint32)
        if len(dims) == 0:
            return np.zeros(0, dtype=np.int32), np.zeros((0, 0), dtype=np.longdouble), np.zeros(0, dtype=np.longdouble)
        md = int(dims.max())
        order, members, totals = [], {}, {}
        for i, s in enumerate(self.slices):
            c = int(s.min_log_alpha_d if axis == 0 else s.min_log_alpha_r)
            if c not in members:
                members[c] = []
                totals[c] = np.longdouble(0)
                order.append(c)
            members[c].append(i)
            totals[c] = totals[c] + np.longdouble(s.total_probability)
        begin = np.zeros(len(order) + 1, dtype=np.uint32)
        lst = []
        for j, c in enumerate(order):
            lst += members[c]
            begin[j + 1] = len(lst)
        lst = np.array(lst, dtype=np.uint32)
        out = np.zeros((len(order), md), dtype=np.longdouble)
        _check(lib().qb200_resident_collapse2d(self.h, axis, dims.ctypes.data, len(order), begin.ctypes.data,
                                               lst.ctypes.data, md, out.ctypes.data), "qb200_resident_collapse2d")
        return np.array(order, dtype=np.int32), out, np.array([totals[c] for c in order], dtype=np.longdouble)

    def format(self, first: int, count: int, prefetch_next: int = 0):
        """The "%.24Lg\\n" text of the slices [first, first + count): a list of bytes objects.
        prefetch_next: start formatting that many slices after them before returning."""
        text = C.c_void_p()
        offsets = np.zeros(max(1, count), dtype=np.uint64)
        lengths = np.zeros(max(1, count), dtype=np.uint64)
        _check(lib().qb200_resident_format(self.h, first, count, C.byref(text), offsets.ctypes.data,
                                           lengths.ctypes.data), "qb200_resident_format")
        out = [C.string_at(text.value + int(offsets[i]), int(lengths[i])) for i in range(count)]
        if prefetch_next:
            _check(lib().qb200_resident_format_prefetch(self.h, first + count, prefetch_next),
                   "qb200_resident_format_prefetch")
        return out


def linear_distribution_init_collapse_d(distribution, ctx=None):
    """linear_distribution_init_collapse_d (src/linear_distribution.cpp:152-237) on the device."""
    return _collapse(distribution, 0, ctx)


def linear_distribution_init_collapse_r(distribution, ctx=None):
    """linear_distribution_init_collapse_r (src/linear_distribution.cpp:239-324) on the device."""
    return _collapse(distribution, 1, ctx)


def _collapse(distribution, axis, ctx):
    res = Resident(distribution.slices, ctx)
    try:
        coords, vec, tot = res.collapse(axis)
    finally:
        res.close()
    out = Linear_Distribution(distribution.m)
    for k in range(len(coords)):
        sl = Linear_Distribution_Slice(vec.shape[1], int(coords[k]), norm_vector=vec[k].copy())
        sl.total_probability = tot[k]
        out.insert_slice(sl)
    out.total_probability = distribution.total_probability
    return out


@dataclass
class Linear_Distribution(Distribution):
    """Linear_Distribution (src/linear_distribution.h)."""


class Sampler:
    """A distribution resident on the GPU (qb200_sampler)."""

    def __init__(self, distribution, ctx: "Context" = None):
        self.ctx = ctx or default_context()
        sl = distribution.slices
        linear = isinstance(distribution, Linear_Distribution) or (
            len(sl) > 0 and isinstance(sl[0], Linear_Distribution_Slice))
        self.dims = 1 if linear else 2
        n = len(sl)
        dim = np.array([x.dimension for x in sl], dtype=np.uint32)
        if linear:
            c0 = _i32([x.min_log_alpha for x in sl])
            c1 = _i32(np.zeros(n))
            cells = [np.ascontiguousarray(x.norm_vector, dtype=np.longdouble) for x in sl]
        else:
            c0 = _i32([x.min_log_alpha_d for x in sl])
            c1 = _i32([x.min_log_alpha_r for x in sl])
            cells = [np.ascontiguousarray(x.norm_matrix, dtype=np.longdouble) for x in sl]
        ptrs = (C.c_void_p * max(1, n))(*[c.ctypes.data for c in cells])
        totals = np.array([x.total_probability for x in sl], dtype=np.longdouble)
        h = C.c_void_p()
        _check(lib().qb200_sampler_create(self.ctx.h, self.dims, distribution.m, n, dim.ctypes.data,
                                          c0.ctypes.data, c1.ctypes.data, ptrs, totals.ctypes.data,
                                          C.c_longdouble(distribution.total_probability), C.byref(h)),
               "qb200_sampler_create")
        self.h = h
        self.words_per_sample = int(lib().qb200_sampler_words_per_sample(h))

    def close(self):
        if getattr(self, "h", None):
            lib().qb200_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_force_exact(self, on):
        """Test switch: 1 / True = bit-exact replay of every walk; 2 = no quick pass in doubles."""
        lib().qb200_sampler_set_force_exact(self.h, int(on))

    @property
    def exact_count(self) -> int:
        return int(lib().qb200_sampler_exact_count(self.h))

    def first_failing_word(self):
        """Smallest slice-pivot word whose walk runs out of bounds, or None."""
        w = C.c_uint64(0)
        any_fail = lib().qb200_sampler_first_failing_word(self.h, C.byref(w))
        return int(w.value) if any_fail else None

    def sample(self, words):
        """len(words) // words_per_sample independent samples: (slice, cell, x0, x1, status)."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        k = w.size // self.words_per_sample
        sl = np.zeros(k, dtype=np.int32)
        ce = np.zeros(k, dtype=np.int32)
        x0 = np.zeros(k)
        x1 = np.zeros(k)
        st = np.zeros(k, dtype=np.int32)
        _check(lib().qb200_sampler_sample(self.h, k, w.ctypes.data, sl.ctypes.data, ce.ctypes.data,
                                          x0.ctypes.data, x1.ctypes.data, st.ctypes.data),
               "qb200_sampler_sample")
        return sl, ce, x0, x1, st

    def tau_estimate(self, n: int, count: int, words):
        """Up to `count` consecutive tau estimates of n samples on the word stream; returns
        (tau0, tau1, ok, words_used) for the estimates completed."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        t0 = np.zeros(count, dtype=np.longdouble)
        t1 = np.zeros(count, dtype=np.longdouble)
        ok = np.zeros(count, dtype=np.uint8)
        used = C.c_size_t(0)
        done = C.c_uint32(0)
        _check(lib().qb200_sampler_tau_estimate(self.h, n, count, w.ctypes.data, w.size, C.byref(used),
                                                C.byref(done), t0.ctypes.data, t1.ctypes.data,
                                                ok.ctypes.data), "qb200_sampler_tau_estimate")
        d = int(done.value)
        return t0[:d], t1[:d], ok[:d].astype(bool), int(used.value)

    def tau_device(self, n: int, count: int, d_words_ptr: int, d_sums_ptr: int, d_status_ptr: int,
                   stream: int = 0):
        _check(lib().qb200_sampler_tau_device(self.h, n, count, d_words_ptr, d_sums_ptr, d_status_ptr,
                                              stream), "qb200_sampler_tau_device")


class WordStream:
    """The random stream as the reference consumes it: consecutive 8-byte draws of a
    Random_State (src/random.c:88-156), handed over as little-endian 64-bit words."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.pos = 0


_samplers = {}


def _sampler_for(distribution, ctx):
    key = id(distribution)
    s = _samplers.get(key)
    if s is None or s[1] is not distribution or s[2] != len(distribution.slices):
        s = (Sampler(distribution, ctx), distribution, len(distribution.slices))
        _samplers[key] = s
    return s[0]


def tau_estimate(distribution, random_state: WordStream, n: int, ctx=None):
    """tau_estimate (src/tau_estimate.cpp:23-87): returns (result, tau_d, tau_r)."""
    s = _sampler_for(distribution, ctx)
    t0, t1, ok, used = s.tau_estimate(n, 1, random_state.words[random_state.pos:])
    if len(t0) == 0:
        raise CriticalError("tau_estimate(): the random stream is exhausted")
    random_state.pos += used
    return bool(ok[0]), t0[0], t1[0]


def tau_estimate_linear(distribution, random_state: WordStream, n: int, ctx=None):
    """tau_estimate_linear (src/tau_estimate.cpp:89-133): returns (result, tau)."""
    s = _sampler_for(distribution, ctx)
    t0, _, ok, used = s.tau_estimate(n, 1, random_state.words[random_state.pos:])
    if len(t0) == 0:
        raise CriticalError("tau_estimate_linear(): the random stream is exhausted")
    random_state.pos += used
    return bool(ok[0]), t0[0]


# --------------------------------------------------------------------------- #
# Diagonal distribution: k given (j, eta)                                     #
# --------------------------------------------------------------------------- #

def int_to_limbs(x: int, n: int) -> np.ndarray:
    """Little-endian 32-bit words (mpz_export(buf, &count, -1, 4, 0, 0, z)), zero padded to n."""
    return np.frombuffer(int(x).to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def limbs_to_int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


class DiagonalKSampler:
    """d, r and the reciprocal of r on the GPU (qb200_diagk)."""

    def __init__(self, parameters: Diagonal_Parameters, ctx: "Context" = None):
        self.ctx = ctx or default_context()
        self.parameters = parameters
        h = C.c_void_p()
        p = parameters._c()
        _check(lib().qb200_diagk_create(self.ctx.h, C.byref(p), C.byref(h)), "qb200_diagk_create")
        self.h = h
        self.j_limbs = int(lib().qb200_diagk_j_limbs(h))
        self.k_limbs = int(lib().qb200_diagk_k_limbs(h))

    def close(self):
        if getattr(self, "h", None):
            lib().qb200_diagk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_force_exact(self, on: bool):
        """Test switch: every walk through the exact x87 path."""
        lib().qb200_diagk_set_force_exact(self.h, int(bool(on)))

    def pack_j(self, js) -> np.ndarray:
        J = np.zeros((len(js), self.j_limbs), dtype=np.uint32)
        for i, j in enumerate(js):
            J[i] = int_to_limbs(j, self.j_limbs)
        return J

    def sample(self, js, etas, pivots, delta_bound: int = 0xffffffff, want_k: bool = True):
        """n calls of sample_k_from_diagonal_j_eta_pivot: (k as Python ints or None, x = alpha_phi /
        2^(m + sigma - l) as rows (hi, lo), delta, status). js: Python ints or packed rows."""
        J = js if isinstance(js, np.ndarray) else self.pack_j(js)
        J = np.ascontiguousarray(J, dtype=np.uint32)
        n = J.shape[0]
        eta = np.ascontiguousarray(etas, dtype=np.int32)
        piv = np.ascontiguousarray(pivots, dtype=np.longdouble)
        K = np.zeros((n, self.k_limbs), dtype=np.uint32) if want_k else None
        xh, xl = np.zeros(n), np.zeros(n)
        delta = np.zeros(n, dtype=np.int64)
        status = np.zeros(n, dtype=np.int32)
        _check(lib().qb200_diagk_sample(self.h, n, J.ctypes.data, eta.ctypes.data, piv.ctypes.data,
                                        delta_bound, K.ctypes.data if want_k else None, xh.ctypes.data,
                                        xl.ctypes.data, delta.ctypes.data, status.ctypes.data),
               "qb200_diagk_sample")
        ks = [limbs_to_int(K[i]) for i in range(n)] if want_k else None
        return ks, np.stack([xh, xl], axis=1), delta, status

    def sample_device(self, n: int, d_j_ptr: int, d_eta_ptr: int, d_pivot_ptr: int, delta_bound: int,
                      d_k_ptr: int, d_out_ptr: int, stream: int = 0):
        _check(lib().qb200_diagk_sample_device(self.h, n, d_j_ptr, d_eta_ptr, d_pivot_ptr, delta_bound,
                                               d_k_ptr or None, d_out_ptr, stream or None),
               "qb200_diagk_sample_device")

    def tau_estimate(self, n: int, count: int, js, etas, pivots, delta_bound: int, eta_bound: int):
        J = js if isinstance(js, np.ndarray) else self.pack_j(js)
        J = np.ascontiguousarray(J, dtype=np.uint32)
        assert J.shape[0] == n * count
        eta = np.ascontiguousarray(etas, dtype=np.int32)
        piv = np.ascontiguousarray(pivots, dtype=np.longdouble)
        tau = np.zeros(count, dtype=np.longdouble)
        ok = np.zeros(count, dtype=np.uint8)
        _check(lib().qb200_diagk_tau_estimate(self.h, n, count, J.ctypes.data, eta.ctypes.data,
                                              piv.ctypes.data, delta_bound, eta_bound, tau.ctypes.data,
                                              ok.ctypes.data), "qb200_diagk_tau_estimate")
        return tau, ok.astype(bool)

    def approx_h(self, x):
        """diagonal_probability_approx_h at phi = 2 pi x / 2^l; x: rows (hi, lo)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 2)
        xh, xl = np.ascontiguousarray(x[:, 0]), np.ascontiguousarray(x[:, 1])
        out = np.zeros(len(xh), dtype=np.longdouble)
        _check(lib().qb200_diagk_h(self.h, len(xh), xh.ctypes.data, xl.ctypes.data, out.ctypes.data),
               "qb200_diagk_h")
        return out


def sample_k_from_diagonal_j_eta_pivot(parameters: Diagonal_Parameters, pivot, j: int, eta: int,
                                       delta_bound: int, ctx=None):
    """sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412-646): returns (result, k,
    alpha_phi / 2^(m + sigma - l) as (hi, lo))."""
    key = (id(parameters), parameters.m, parameters.sigma, parameters.l, parameters.d, parameters.r)
    s = _diagk.get(key)
    if s is None:
        s = _diagk[key] = DiagonalKSampler(parameters, ctx)
    ks, x, _, st = s.sample([j], [eta], [pivot], delta_bound)
    if st[0] == 4:
        raise CriticalError("sample_k_from_diagonal_j_eta_pivot(): gave up after 2^22 steps")
    if st[0] == 2:  # QB200_DIAGK_OK_NEGATIVE_PHI
        return True, ks[0], (x[0, 0] - 2.0 ** parameters.l, x[0, 1])
    return st[0] == 0, ks[0], (x[0, 0], x[0, 1])


_diagk = {}


# --------------------------------------------------------------------------- #
# Exact samplers: alpha from a region, (j, k) from alpha                      #
# --------------------------------------------------------------------------- #

EXACT_TWO_DIMENSIONAL, EXACT_DIAGONAL = 0, 1
EXACT_REGION_DTYPE = np.dtype([("min_log_alpha", "<i4"), ("region", "<u4"), ("dimension", "<u4"),
                               ("length", "<u4"), ("offset", "<u8")])


def _abs_rows(vals, w: int) -> np.ndarray:
    out = np.zeros((len(vals), w), dtype=np.uint32)
    for i, v in enumerate(vals):
        out[i] = int_to_limbs(abs(int(v)), w)
    return out


def _signs(vals) -> np.ndarray:
    return np.array([1 if v < 0 else 0 for v in vals], dtype=np.int32)


def pack_regions(regions) -> np.ndarray:
    """(min_log_alpha, region, dimension, offset, length) per sample -> qb200_exact_region records."""
    g = np.zeros(len(regions), dtype=EXACT_REGION_DTYPE)
    for i, (a, reg, dim, off, ln) in enumerate(regions):
        g[i] = (a, reg, dim, ln, off)
    return g


class ExactSampler:
    """sample_alpha_from_region and the (j, k) samplers of src/sample.cpp:78-410 for batches, on the
    GPU (qb200_exact). parameters: Parameters (kind EXACT_TWO_DIMENSIONAL) or Diagonal_Parameters
    (EXACT_DIAGONAL)."""

    def __init__(self, parameters, dimension_max: int, emax: int = 0, ctx: "Context" = None):
        self.ctx = ctx or default_context()
        self.parameters = parameters
        self.kind = EXACT_DIAGONAL if isinstance(parameters, Diagonal_Parameters) else EXACT_TWO_DIMENSIONAL
        h = C.c_void_p()
        p = parameters._c()
        _check(lib().qb200_exact_create(self.ctx.h, C.byref(p), self.kind, dimension_max, emax, C.byref(h)),
               "qb200_exact_create")
        self.h = h
        out = (C.c_uint32 * 6)()
        lib().qb200_exact_dims(h, out)
        self.wa, self.wn, self.wk, self.kappa_d, self.kappa_r, self.emax = [int(x) for x in out]

    def close(self):
        if getattr(self, "h", None):
            lib().qb200_exact_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def kernel_ms(self):
        """(k_exact_alpha, k_exact_jk): CUDA-event durations of the last launches, milliseconds."""
        out = (C.c_float * 2)()
        _check(lib().qb200_exact_kernel_ms(self.h, out), "qb200_exact_kernel_ms")
        return float(out[0]), float(out[1])

    def region_bytes(self, min_log_alpha: int, region: int, dimension: int):
        """(bytes random_generate_mpz reads, status) -- status 0, 2 (ambiguous) or 3 (unsupported)."""
        n = C.c_uint32(0)
        rc = lib().qb200_exact_region_bytes(self.h, min_log_alpha, region, dimension, C.byref(n))
        return int(n.value), {0: 0, -51: 2, -50: 3}.get(rc, rc)

    def alpha(self, regions, kappa: int, stream: bytes):
        """n calls of sample_alpha_from_region: signed Python ints and the status codes."""
        g = regions if isinstance(regions, np.ndarray) else pack_regions(regions)
        buf = np.frombuffer(stream, dtype=np.uint8)
        n = len(g)
        rows = np.zeros((n, self.wa), dtype=np.uint32)
        neg = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        _check(lib().qb200_exact_alpha(self.h, n, g.ctypes.data, kappa, buf.ctypes.data, len(buf),
                                       rows.ctypes.data, neg.ctypes.data, st.ctypes.data), "qb200_exact_alpha")
        vals = [(-limbs_to_int(r) if s else limbs_to_int(r)) for r, s in zip(rows, neg)]
        return vals, st

    def _jk(self, mode, alpha_d, alpha_r, t, k_in=None):
        n = len(alpha_d if alpha_d is not None else alpha_r)
        kap = self.kappa_d if mode == 2 else self.kappa_r
        tl = max(1, (kap + 31) // 32)
        ad = _abs_rows(alpha_d, self.wa) if alpha_d is not None else None
        ar = _abs_rows(alpha_r, self.wa) if alpha_r is not None else None
        nd = _signs(alpha_d) if alpha_d is not None else None
        nr = _signs(alpha_r) if alpha_r is not None else None
        tt = _abs_rows(t, tl) if (t is not None and kap) else None
        j = np.zeros((n, self.wn), dtype=np.uint32)
        k = _abs_rows(k_in, self.wk) if k_in is not None else np.zeros((n, max(1, self.wk)), dtype=np.uint32)
        p = lambda a: a.ctypes.data if a is not None else None
        _check(lib().qb200_exact_j_k(self.h, mode, n, p(ad), p(nd), p(ar), p(nr), p(tt), p(j), p(k)),
               "qb200_exact_j_k")
        return [limbs_to_int(r) for r in j], [limbs_to_int(r) for r in k]

    def j_from_alpha_r(self, alpha_r, t=None):
        """sample_j_from_alpha_r / sample_j_from_diagonal_alpha_r with t_r as drawn by the caller."""
        return self._jk(0, None, alpha_r, t)[0]

    def j_k_from_alpha_d_r(self, alpha_d, alpha_r, t=None):
        """sample_j_k_from_alpha_d_r: (j list, k list)."""
        return self._jk(1, alpha_d, alpha_r, t)

    def j_from_alpha_d_k(self, alpha_d, k, t=None):
        """sample_j_k_from_alpha_d with k and t_d as drawn by the caller."""
        return self._jk(2, alpha_d, None, t, k_in=k)[0]


def diagonal_sample_drawn(diagk: DiagonalKSampler, exact: ExactSampler, regions, t_r, stream: bytes, etas,
                          pivots, delta_bound: int = 0xffffffff, want_k: bool = True):
    """qb200_diagk_sample_drawn: a diagonal sample from its random bytes to k on the device. Returns
    (k ints or None, x rows (hi, lo), delta, status, exact_status)."""
    g = regions if isinstance(regions, np.ndarray) else pack_regions(regions)
    n = len(g)
    buf = np.frombuffer(stream, dtype=np.uint8)
    tl = max(1, (exact.kappa_r + 31) // 32)
    tt = _abs_rows(t_r, tl) if (t_r is not None and exact.kappa_r) else None
    eta = np.ascontiguousarray(etas, dtype=np.int32)
    piv = np.ascontiguousarray(pivots, dtype=np.longdouble)
    K = np.zeros((n, diagk.k_limbs), dtype=np.uint32) if want_k else None
    xh, xl = np.zeros(n), np.zeros(n)
    delta = np.zeros(n, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    est = np.zeros(n, dtype=np.int32)
    _check(lib().qb200_diagk_sample_drawn(diagk.h, exact.h, n, g.ctypes.data, tt.ctypes.data if tt is not None else None,
                                          buf.ctypes.data, len(buf), eta.ctypes.data, piv.ctypes.data, delta_bound,
                                          K.ctypes.data if want_k else None, xh.ctypes.data, xl.ctypes.data,
                                          delta.ctypes.data, status.ctypes.data, est.ctypes.data),
           "qb200_diagk_sample_drawn")
    ks = [limbs_to_int(K[i]) for i in range(n)] if want_k else None
    return ks, np.stack([xh, xl], axis=1), delta, status, est


# ---- the reference's own entry points (src/sample.h), one sample per call --------------------------

class ByteStream:
    """The random stream as random_generate (src/random.c:88-113) delivers it: bytes, in order."""

    def __init__(self, data: bytes):
        self.data = bytes(data)
        self.pos = 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise CriticalError("random_generate(): the random stream is exhausted")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def mpz(self, modulus_bits: int, modulus: int) -> int:
        """random_generate_mpz (src/random.c:158-181) for a modulus of the given bit length."""
        return int.from_bytes(self.take((modulus_bits + 64 + 8) // 8), "big") % modulus


_exact = {}


def _exact_for(parameters, sampler=None, ctx=None):
    """The ExactSampler of a parameter set (kept; any power-of-two dimension up to 16384), or the object
    handed in (tests pass the CPU twin, which has the same methods)."""
    if sampler is not None:
        return sampler
    key = (type(parameters).__name__, parameters.m, parameters.l, getattr(parameters, "sigma", 0), parameters.d,
           parameters.r)
    s = _exact.get(key)
    if s is None:
        s = _exact[key] = ExactSampler(parameters, 16384, 0, ctx)
    return s


def _region_of(min_log_alpha: float, max_log_alpha: float, who: str):
    """(signed slice coordinate, region index, dimension) of a region given by its bounds as
    distribution_slice_region_coordinates (src/distribution_slice.cpp:130-165) forms them."""
    sgn = lambda x: -1 if x < 0 else 1  # sgn_d, src/math.cpp:32-34
    if sgn(min_log_alpha) != sgn(max_log_alpha):
        raise CriticalError(f"{who}(): Incompatible signs for min_log_alpha and max_log_alpha.")
    lo, hi = abs(min_log_alpha), abs(max_log_alpha)
    if lo >= hi:
        raise CriticalError(f"{who}(): Incompatible absolute values for min_log_alpha and max_log_alpha.")
    e = math.floor(lo)
    dim = 1.0 / (hi - lo)
    region = (lo - e) * dim
    if dim != int(dim) or int(dim) & (int(dim) - 1) or region != int(region):
        raise CriticalError(f"{who}(): [{min_log_alpha}, {max_log_alpha}] is not a region of a slice "
                            "(e + i / D with a power of two D).")
    return sgn(min_log_alpha) * e, int(region), int(dim)


def sample_alpha_from_region(min_log_alpha: float, max_log_alpha: float, kappa: int, random_state: ByteStream,
                             parameters=None, sampler=None, ctx=None) -> int:
    """sample_alpha_from_region (src/sample.cpp:78-158): alpha, reading from random_state what the
    reference reads."""
    ex = _exact_for(parameters, sampler, ctx)
    se, region, dim = _region_of(min_log_alpha, max_log_alpha, "sample_alpha_from_region")
    nbytes, status = ex.region_bytes(se, region, dim)
    if status:
        raise CriticalError("sample_alpha_from_region(): the region is outside the sampler's range.")
    vals, st = ex.alpha([(se, region, dim, 0, nbytes)], kappa, random_state.take(nbytes))
    if st[0]:
        raise CriticalError(f"sample_alpha_from_region(): status {int(st[0])}")
    return vals[0]


def _draw_t(kappa: int, random_state: ByteStream) -> int:
    # random_generate_mpz(t, 2^kappa, random_state), src/sample.cpp:176-180: 2^kappa has kappa + 1 bits
    return random_state.mpz(kappa + 1, 1 << kappa) if kappa > 0 else 0


def sample_j_from_alpha_r(alpha_r: int, parameters, random_state: ByteStream, sampler=None, ctx=None) -> int:
    """sample_j_from_alpha_r (src/sample.cpp:160-208)."""
    ex = _exact_for(parameters, sampler, ctx)
    return ex.j_from_alpha_r([alpha_r], [_draw_t(ex.kappa_r, random_state)])[0]


def sample_j_from_diagonal_alpha_r(alpha_r: int, parameters, random_state: ByteStream, sampler=None, ctx=None) -> int:
    """sample_j_from_diagonal_alpha_r (src/sample.cpp:354-410); parameters: Diagonal_Parameters."""
    return sample_j_from_alpha_r(alpha_r, parameters, random_state, sampler, ctx)


def sample_j_k_from_alpha_d(alpha_d: int, parameters, random_state: ByteStream, sampler=None, ctx=None):
    """sample_j_k_from_alpha_d (src/sample.cpp:210-273): (j, k); t_d and k are drawn as the reference draws
    them (:230-239)."""
    ex = _exact_for(parameters, sampler, ctx)
    t = _draw_t(ex.kappa_d, random_state)
    k = random_state.mpz(parameters.l + 1, 1 << parameters.l)
    return ex.j_from_alpha_d_k([alpha_d], [k], [t])[0], k


def sample_j_k_from_alpha_d_r(alpha_d: int, alpha_r: int, parameters, random_state: ByteStream, sampler=None,
                              ctx=None):
    """sample_j_k_from_alpha_d_r (src/sample.cpp:275-352): (j, k); t_r a multiple of 2^kappa_t_r (:294-309)."""
    ex = _exact_for(parameters, sampler, ctx)
    kt = max(0, ex.kappa_r - ex.kappa_d - parameters.l)
    t = 0
    if ex.kappa_r > 0:
        t = random_state.mpz(ex.kappa_r - kt + 1, 1 << (ex.kappa_r - kt)) << kt
    js, ks = ex.j_k_from_alpha_d_r([alpha_d], [alpha_r], [t])
    return js[0], ks[0]
