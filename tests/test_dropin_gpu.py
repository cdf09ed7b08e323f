"""The reference-side forwarding TU (qunundrum_b200/dropin/dropin.cpp), compiled against
the reference's own headers, called exactly as the generator clients call the
reference: C++ symbols, Distribution_Slice / Parameters structs built by the
reference's own parameters_* functions (from oracle/_ref)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ref_or_none
from tests.util import CELL_RTOL, cell_errors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "qunundrum_b200", "dropin", "libqunundrum_dropin.so")

SYM_2D = "_Z37distribution_slice_compute_richardsonP18Distribution_SlicePK10Parameters33Distribution_Slice_Compute_Methodii"
SYM_2D_SINGLE = "_Z26distribution_slice_computeP18Distribution_SlicePK10Parameters33Distribution_Slice_Compute_Methodii"
SYM_LIN = "_Z44linear_distribution_slice_compute_richardsonP25Linear_Distribution_SlicePK10Parameters40Linear_Distribution_Slice_Compute_Targeti"
SYM_DIAG = "_Z46diagonal_distribution_slice_compute_richardsonP27Diagonal_Distribution_SlicePK19Diagonal_Parametersii"


class Distribution_Slice(C.Structure):   # src/distribution_slice.h:86-139
    _fields_ = [("dimension", C.c_uint32), ("min_log_alpha_d", C.c_int32),
                ("min_log_alpha_r", C.c_int32), ("total_probability", C.c_longdouble),
                ("total_error", C.c_longdouble), ("flags", C.c_uint32),
                ("norm_matrix", C.c_void_p)]


class Linear_Distribution_Slice(C.Structure):   # src/linear_distribution_slice.h:48-91
    _fields_ = [("dimension", C.c_uint32), ("min_log_alpha", C.c_int32),
                ("total_probability", C.c_longdouble), ("total_error", C.c_longdouble),
                ("flags", C.c_uint32), ("norm_vector", C.c_void_p)]


class Diagonal_Distribution_Slice(C.Structure):   # src/diagonal_distribution_slice.h:32-83
    _fields_ = [("dimension", C.c_uint32), ("min_log_alpha_r", C.c_int32), ("eta", C.c_int32),
                ("total_probability", C.c_longdouble), ("total_error", C.c_longdouble),
                ("flags", C.c_uint32), ("norm_vector", C.c_void_p)]


def test_dropin_exports_the_reference_symbols():
    if not os.path.exists(DROPIN):
        pytest.skip("drop-in not built (needs the reference headers at build time)")
    out = subprocess.run(["nm", "-D", "--defined-only", DROPIN], capture_output=True, text=True).stdout
    for s in (SYM_2D, SYM_2D_SINGLE, SYM_LIN, SYM_DIAG):
        assert s in out


@pytest.mark.gpu
def test_dropin_called_like_the_reference():
    ref = ref_or_none()
    if ref is None or not os.path.exists(DROPIN):
        pytest.skip("needs oracle/_ref and the built drop-in")
    os.environ["QB200_DEVICE"] = "0"
    L = C.CDLL(DROPIN, mode=os.RTLD_LOCAL)
    m, s, D = 128, 2, 32
    d, r = ref.deterministic_d_r(m)
    RP = ref.RefParameters(m, s, d, r)

    cells = np.zeros(D * D, dtype=np.longdouble)
    sl = Distribution_Slice(D, 0, 0, 0, 0, 0x00000100, cells.ctypes.data)
    f = getattr(L, SYM_2D)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_int32]
    f.restype = None
    f(C.byref(sl), RP.h, 0, 130, 129)
    R = ref.distribution_slice_compute(RP, D, 130, 129)
    assert cell_errors(cells, R.cells) <= CELL_RTOL
    assert sl.flags == (0x00000100 | R.flags) and (sl.min_log_alpha_d, sl.min_log_alpha_r) == (130, 129)
    assert abs(float(np.longdouble(sl.total_probability) - R.total_probability)) <= 1e-12
    assert abs(float((np.longdouble(sl.total_error) - R.total_error) / R.total_error)) <= 1e-9

    vec = np.zeros(64, dtype=np.longdouble)
    ls = Linear_Distribution_Slice(64, 0, 0, 0, 0, vec.ctypes.data)
    g = getattr(L, SYM_LIN)
    g.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int32]
    g.restype = None
    g(C.byref(ls), RP.h, 0, -127)
    R = ref.linear_distribution_slice_compute(RP, 64, -127, 0)
    assert cell_errors(vec, R.cells) <= CELL_RTOL and ls.flags == R.flags and ls.min_log_alpha == -127

    RDP = ref.RefDiagonalParameters(m, 5, 1, d, r, eta_bound=25)
    ds = Diagonal_Distribution_Slice(64, 0, 0, 0, 0, 0, vec.ctypes.data)
    h = getattr(L, SYM_DIAG)
    h.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    h.restype = None
    h(C.byref(ds), RDP.h, 126, -2)
    R = ref.diagonal_distribution_slice_compute(RDP, 64, 126, -2)
    assert cell_errors(vec, R.cells) <= CELL_RTOL and ds.eta == -2 and ds.min_log_alpha_r == 126


SYM_TAU = "_Z12tau_estimatePK12DistributionP12Random_StatejReS4_"
SYM_TAU_LINEAR = "_Z19tau_estimate_linearPK19Linear_DistributionP12Random_StatejRe"


def test_dropin_exports_the_tau_symbols():
    if not os.path.exists(DROPIN):
        pytest.skip("drop-in not built (needs the reference headers at build time)")
    out = subprocess.run(["nm", "-D", "--defined-only", DROPIN], capture_output=True, text=True).stdout
    assert SYM_TAU in out and SYM_TAU_LINEAR in out


def _tau_dropin_called_like_the_reference(libpath, name):
    """dropin_tau.cpp on the reference's own Distribution / Random_State structs (built by the
    reference's functions in oracle/_ref): for the same generator seed, the sequence of
    tau_estimate() calls an estimate_runs client makes returns what the reference's
    tau_estimate() returns -- across batch boundaries (QB200_TAU_BATCH = 64 here, 128 calls per
    n) and across a change of n, with the words of early-failing estimates carried over."""
    ref = ref_or_none()
    if ref is None or not os.path.exists(libpath):
        pytest.skip("needs oracle/_ref and the built drop-in")
    from tests.test_sampler import GOLD, Gold, assert_tau
    g = Gold(np.load(GOLD), name)
    os.environ["QB200_DEVICE"] = "0"
    os.environ["QB200_TAU_BATCH"] = "64"
    try:
        L = C.CDLL(libpath, mode=os.RTLD_LOCAL)
        d, r = ref.deterministic_d_r(g.m)
        P = ref.RefParameters(g.m, 2, d, r)
        dist = ref.RefDistribution(g.dims, P, g.dimension, g.c0, g.c1, g.cells, g.totals)
        seed = bytes(range(50, 82))
        if g.dims == 2:
            f = getattr(L, SYM_TAU)
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        else:
            f = getattr(L, SYM_TAU_LINEAR)
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        f.restype = C.c_bool
        rs_gpu, rs_ref = ref.RefRandom(seed), ref.RefRandom(seed)

        def call(n):
            t0 = np.zeros(1, dtype=np.longdouble)
            t1 = np.zeros(1, dtype=np.longdouble)
            if g.dims == 2:
                ok = f(dist.ptr(), rs_gpu.h, n, t0.ctypes.data, t1.ctypes.data)
            else:
                ok = f(dist.ptr(), rs_gpu.h, n, t0.ctypes.data)
            return bool(ok), t0[0], t1[0]
        # two full batches of n = 3, two of n = 5, one of n = 2: words queued by the drop-in
        # (estimates that failed early) must be consumed exactly as the reference consumes them,
        # or the later estimates drift apart
        for n, calls in ((3, 128), (5, 128), (2, 64)):
            want0, want1, want_ok = dist.tau_estimate(rs_ref, n, calls)
            for i in range(calls):
                ok, t0, t1 = call(n)
                assert ok == bool(want_ok[i]), (n, i)
                if ok:
                    assert_tau(t0, want0[i], g.m)
                    if g.dims == 2:
                        assert_tau(t1, want1[i], g.m)
                else:
                    assert float(t0) == np.finfo(np.float64).max
        # n = 0: the reference's loop does not run -> FALSE, DBL_MAX, nothing drawn
        ok, t0, _ = call(0)
        assert not ok and float(t0) == np.finfo(np.float64).max
        # (the Random_State itself may be AHEAD of the reference's by the words still queued in
        # the drop-in -- the logical stream, queue first, is what stays in step, as the rounds
        # above show)
    finally:
        os.environ.pop("QB200_TAU_BATCH", None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2d", "lin"])
def test_tau_dropin_called_like_the_reference(name):
    _tau_dropin_called_like_the_reference(DROPIN, name)


@pytest.mark.parametrize("name", ["2d", "lin"])
def test_tau_dropin_host_logic_on_the_cpu_shim(name):
    """The same scenario in the GPU-less suite: dropin_tau.cpp and csrc/sampler_host.hpp (stream
    layout, failure threshold, batching, the queue of pre-drawn words, the fast Keccak stream)
    over the TEST-ONLY CPU stand-in of the sampler entry points (tests/hostsim/abi_shim.cpp)."""
    from tests.hostsim import shim_flavour
    lib = shim_flavour.build_tau()
    if lib is None:
        pytest.skip("tests/hostsim/_build/shim/libdropin_tau_shim.so missing (needs /root/reference at build time)")
    _tau_dropin_called_like_the_reference(lib, name)


def test_tau_dropin_draws_the_reference_stream():
    """draw_words() of dropin_tau.cpp (unrolled Keccak-f[1600], whole-lane extraction) against the
    reference's random_generate() (src/random.c:88, src/keccak_random.c:96) on the reference's own
    Random_State: the same bytes, and the same state afterwards (the next draws of the
    reference's generator agree), for lengths around the 21-lane block boundary, for a state
    left in mid-lane by a 3-byte draw (falls back to the reference), and for 2^18 words.
    Host code only: runs without a GPU."""
    ref = ref_or_none()
    if ref is None or not os.path.exists(DROPIN):
        pytest.skip("needs oracle/_ref and the built drop-in")
    L = C.CDLL(DROPIN, mode=os.RTLD_LOCAL)
    f = L.qb200_dropin_tau_draw
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    f.restype = None
    seed = bytes((11 * i + 1) & 0xff for i in range(32))
    for pre_bytes in (0, 8, 3):
        for n in (1, 20, 21, 22, 41, 42, 43, 1000, 1 << 18):
            a, b = ref.RefRandom(seed), ref.RefRandom(seed)
            if pre_bytes:
                a.bytes(pre_bytes)
                b.bytes(pre_bytes)
            x = np.zeros(n, dtype=np.uint64)
            y = np.zeros(n, dtype=np.uint64)
            f(a.h, x.ctypes.data, n, 0)
            f(b.h, y.ctypes.data, n, 1)
            assert (x == y).all(), (pre_bytes, n)
            assert a.bytes(300) == b.bytes(300), (pre_bytes, n)     # the states are in step
    # several consecutive fast draws == one reference draw
    a, b = ref.RefRandom(seed), ref.RefRandom(seed)
    parts = []
    for n in (5, 16, 1, 100, 21, 64):
        x = np.zeros(n, dtype=np.uint64)
        f(a.h, x.ctypes.data, n, 0)
        parts.append(x)
    assert (np.concatenate(parts) == b.words(207)).all()
