"""Host logic, the C ABI surface and error behaviour -- no GPU required."""
import ctypes as C
import math
import os
import re
from fractions import Fraction as Fr

import numpy as np
import pytest

import qunundrum_b200 as qb
from oracle import restate as rs
from tests import hostsim as hs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built_and_exports_the_declared_abi():
    from qunundrum_b200 import build as qbuild
    qbuild.build()
    L = qb.lib()
    hdr = open(os.path.join(ROOT, "include", "qunundrum_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/qunundrum_b200.h but not exported"
    assert L.qb200_version() == 2


def test_no_cpu_fallback_without_a_device():
    if qb.lib().qb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(qb.CriticalError, match="no CPU path"):
        qb.Context(0)
    sl = qb.Distribution_Slice(16)
    d, r = rs.deterministic_d_r(128)
    with pytest.raises(qb.CriticalError):
        qb.distribution_slice_compute_richardson(sl, qb.Parameters(128, 2, d, r), 0, 128, 128)


def test_product_never_touches_the_oracle():
    """Nothing under qunundrum_b200/, include/ or integration/ (the product, its C ABI and the
    reference-side integration build) may import, load, link, include or build from oracle/ or
    tests/ -- not even a header path in a compiler command line."""
    for top in ("qunundrum_b200", "include", "integration"):
        for root, dirs, files in os.walk(os.path.join(ROOT, top)):
            dirs[:] = [d for d in dirs if d not in ("_build", "_obj", "__pycache__")]
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c", ".sh")):
                    txt = open(os.path.join(root, f), errors="replace").read()
                    assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", txt, re.M), f
                    assert "libqref" not in txt and "libhostsim" not in txt, f
                    assert not re.search(r'#include\s+"[^"]*(oracle|hostsim)', txt), f
                    # path components in build commands: os.path.join(..., "oracle", ...) etc.
                    assert not re.search(r'["\'](oracle|tests|hostsim)["\']', txt), f


@pytest.mark.parametrize("l", list(range(3, 70)) + [128, 683, 768, 1023, 2048, 3072, 8192])
def test_heuristic_sigma(l):
    # sigma = round((l + 11 + 4 - 1.6515) / 2) in float32, src/distribution_slice_compute.cpp:149-158
    want = rs.heuristic_sigma(l)
    assert qb.heuristic_sigma(l) == want == hs.heuristic_sigma(l)


def _rnd(x, prec=192):
    """Round a positive Fraction to `prec` bits, nearest-even (MPFR_RNDN)."""
    e = x.numerator.bit_length() - x.denominator.bit_length()
    sh = prec - e
    while True:
        y = x * Fr(2) ** sh
        fl = y.numerator // y.denominator
        if fl.bit_length() > prec:
            sh -= 1
        elif fl.bit_length() < prec:
            sh += 1
        else:
            break
    rem = y - fl
    if rem > Fr(1, 2) or (rem == Fr(1, 2) and fl % 2 == 1):
        fl += 1
    return Fr(fl) / Fr(2) ** sh


def _dd_err(v, exact):
    got = Fr(float(v[0])) + Fr(float(v[1]))
    return float(abs(got - exact) / abs(exact)) if exact != 0 else float(abs(got))


@pytest.mark.parametrize("m,s", [(128, 2), (128, 1), (256, 3), (1023, 8), (2048, 1), (3072, 4)])
def test_host_constants_reproduce_the_mpfr_roundings(m, s):
    d, r = rs.deterministic_d_r(m)
    l = math.ceil(m / s)
    sigma = rs.heuristic_sigma(l)
    c = qb.host_constants(m, l, sigma, d, r)
    K = -math.floor(_rnd(_rnd(Fr(2) ** sigma * d) / r))   # src/probability.cpp:165-170
    Q = _rnd(Fr(2) ** (m + l) / r)                        # src/probability.cpp:216-220
    N, Cc = math.floor(Q), math.ceil(Q)
    beta = pow(2, l + m, r)                               # src/linear_probability.cpp:190-192
    exact = {
        "kappa": Fr(K, 2 ** sigma), "kappa_q": -_rnd(_rnd(Fr(d)) / r),
        "c_over_L": Fr(Cc, 2 ** l), "n_over_L": Fr(N, 2 ** l), "n1_over_L": _rnd(Fr(N + 1)) / 2 ** l,
        "beta_m": Fr(beta, 2 ** m), "rbeta_m": Fr(r - beta, 2 ** m), "r_m": Fr(r, 2 ** m),
        "d_m": Fr(d, 2 ** m), "rho": Fr(2 ** m, r),
    }
    for k, ex in exact.items():
        assert _dd_err(c[k], ex) < 2e-32, k
    if sigma <= 100:   # K_sigma is then an exact integer: the double-double must be EXACT
        assert Fr(float(c["kappa"][0])) + Fr(float(c["kappa"][1])) == exact["kappa"]
    assert c == hs.host_consts(m, l, sigma, d, r)


def test_exp2_table():
    import mpmath as mp
    for n in (12, 256, 4096):
        t = hs.exp2_table(n)
        assert t[0, 0] == 1.0 and t[0, 1] == 0.0
        if n & (n - 1) == 0:
            assert t[n, 0] == 2.0 and t[n, 1] == 0.0
        with mp.workprec(250):
            step = mp.mpf(1.0 / n)   # the reference's step is this double (exact for powers of two)
            for i in range(0, n + 1, max(1, n // 64)):
                ex = mp.power(2, step * i)
                assert abs((mp.mpf(float(t[i, 0])) + mp.mpf(float(t[i, 1]))) / ex - 1) < 1e-31


def test_parameter_regions_match_the_reference_rules():
    # parameters_setup_regions, src/parameters.cpp:30-51
    d, r = rs.deterministic_d_r(128)
    for kw in (dict(s=2), dict(s=1), dict(s=8), dict(s=0, l=20), dict(s=0, l=31)):
        a = qb.Parameters(128, d=d, r=r, **kw)
        b = rs.Parameters(128, d=d, r=r, **kw)
        for f in ("l", "min_alpha_d", "max_alpha_d", "min_alpha_r", "max_alpha_r"):
            assert getattr(a, f) == getattr(b, f)
    a = qb.Diagonal_Parameters(128, 5, 1, d, r, eta_bound=25)
    b = rs.DiagonalParameters(128, 5, 1, d, r, eta_bound=25)
    assert (a.l, a.min_alpha_r, a.max_alpha_r) == (b.l, b.min_alpha_r, b.max_alpha_r)
    assert qb.Diagonal_Parameters(128, 40, 1, d, r).max_alpha_r == 128 + 30 - 1


def test_flag_semantics():
    from qunundrum_b200.host import _apply_flags
    old = 0x00000100 | 0x00040000          # MIRRORED + some stale method bit
    new = _apply_flags(old, qb.SLICE_FLAGS_METHOD_SIMPSON | qb.SLICE_FLAGS_METHOD_RICHARDSON)
    assert new == 0x00000100 | 0x00020000 | 0x00080000


def test_enumeration_and_partition():
    from qunundrum_b200 import shard
    coords = shard.enumerate_2d(2048)
    assert len(coords) == 3362 and len(set(coords)) == 3362          # 2 * 41 * 41
    assert coords[0] in ((2048, 2048), (-2048, 2048))
    dist = [(abs(a) - 2048) ** 2 + (b - 2048) ** 2 for a, b in coords]
    assert dist == sorted(dist)
    for world in (1, 2, 3, 8):
        parts = [shard.partition(len(coords), world, k) for k in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(coords)))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
