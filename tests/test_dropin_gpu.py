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
