"""Pins the oracle: the mpmath restatement (oracle/restate.py) and, when built,
the compiled reference (oracle/_ref) against

  * a committed sample of the reference's known-answer vectors
    (res/test-vectors -> tests/golden/kat, src/test/test_probability.cpp:31-189,
    test_linear_probability.cpp:29-224, test_diagonal_probability.cpp:31-172);
  * the Mathematica NIntegrate slice totals quoted in the reference's
    src/test/test_linear_distribution.cpp:93-131,353-388 and
    src/test/test_diagonal_distribution.cpp:87-389 (tolerance 1e-6, as there);
  * each other, cell by cell, on full slices.
"""
import glob
import json
import os
import re

import mpmath as mp
import numpy as np
import pytest

from oracle import restate as rs
from tests.conftest import GOLDEN, golden_slices, ref_or_none
from tests.util import cell_errors

KAT = os.path.join(GOLDEN, "kat")
REF = ref_or_none()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference)")


def _records(path, n):
    lines = open(path).read().split()
    return [lines[i:i + n] for i in range(0, len(lines) - n + 1, n)]


def _rel(a, b):
    with mp.workprec(300):
        return float(abs(mp.mpf(a) / mp.mpf(b) - 1))


def _mf(s):
    with mp.workprec(rs.PRECISION):
        return mp.mpf(s)


def _check_kat_errors(errs):
    """The vectors carry 38 digits of theta AND of the value; where the integrand
    oscillates fast in theta the truncated theta limits the agreement, so: every
    record within the reference's own tolerance (1e-6, test_cmp_ld), and the
    bulk at the precision of the vectors."""
    errs = sorted(errs)
    assert errs[-1] < 1e-6, errs[-1]
    assert errs[len(errs) // 2] < 1e-30, errs[len(errs) // 2]


# ---- point-wise integrands -------------------------------------------------------

FILES_2D = sorted(glob.glob(os.path.join(KAT, "probabilities-det-m-*.txt")))


@pytest.mark.parametrize("path", FILES_2D, ids=os.path.basename)
def test_probability_approx_kat(path):
    m, s = map(int, re.search(r"m-(\d+)-s-(\d+)", path).groups())
    d, r = rs.deterministic_d_r(m)
    P = rs.Parameters(m, s, d, r)
    sigma = rs.heuristic_sigma(P.l)
    recs = _records(path, 4)
    assert len(recs) >= 32
    RP = REF.RefParameters(m, s, d, r) if REF else None
    errs = []
    for td, tr, en, ee in recs[::2 if m > 1024 else 1]:
        n, e, _ = rs.probability_approx(sigma, _mf(td), _mf(tr), P)
        errs += [_rel(n, en), _rel(e, ee)]
        q = rs.probability_approx_quick(_mf(td), _mf(tr), P)
        assert _rel(q, en) < 1e-4  # the reference's own check of the quick form
        if RP is not None:
            # same 192-bit theta in, same value out (mpmath vs MPFR differ by ulps)
            n2, e2, _ = REF.probability_approx(RP, sigma, td, tr)
            assert _rel(n2, n) < 1e-45 and _rel(e2, e) < 1e-45
            assert _rel(REF.probability_approx_quick(RP, td, tr), q) < 1e-45
    _check_kat_errors(errs)


FILES_LIN = sorted(glob.glob(os.path.join(KAT, "linear-probabilities-det-*.txt")))


@pytest.mark.parametrize("path", FILES_LIN, ids=os.path.basename)
def test_linear_probability_kat(path):
    t, m, s = re.search(r"det-([dr])-m-(\d+)-s-(\d+)", path).groups()
    m, s = int(m), int(s)
    if m > 2048:
        pytest.skip("24576-bit mpmath sines: minutes; covered by oracle/_ref below")
    d, r = rs.deterministic_d_r(m)
    P = rs.Parameters(m, s, d, r)
    f = rs.linear_probability_d if t == "d" else rs.linear_probability_r
    recs = _records(path, 2)
    step = 8 if (t == "d" and m >= 2048) else 1
    _check_kat_errors([_rel(f(_mf(th), P), en) for th, en in recs[::step]])


@needs_ref
@pytest.mark.parametrize("path", FILES_LIN, ids=os.path.basename)
def test_linear_probability_kat_ref(path):
    t, m, s = re.search(r"det-([dr])-m-(\d+)-s-(\d+)", path).groups()
    m, s = int(m), int(s)
    d, r = REF.deterministic_d_r(m)
    RP = REF.RefParameters(m, s, d, r)
    recs = _records(path, 2)
    step = 8 if (t == "d" and m >= 2048) else 1
    _check_kat_errors([_rel(REF.linear_probability(RP, 0 if t == "d" else 1, th), en)
                       for th, en in recs[::step]])
    if m <= 512:  # restatement == reference on identical input
        P = rs.Parameters(m, s, d, r)
        f = rs.linear_probability_d if t == "d" else rs.linear_probability_r
        for th, _ in recs[::6]:
            assert _rel(REF.linear_probability(RP, 0 if t == "d" else 1, th), f(_mf(th), P)) < 1e-45


FILES_DIAG = sorted(glob.glob(os.path.join(KAT, "diagonal-probabilities-f-eta-*.txt")))


@pytest.mark.parametrize("path", FILES_DIAG, ids=os.path.basename)
def test_diagonal_f_eta_kat(path):
    m, sigma, s = map(int, re.search(r"m-(\d+)-sigma-(\d+)-s-(\d+)", path).groups())
    l = int(np.ceil(m / s))
    d, r = rs.deterministic_d_r(m)
    P = rs.DiagonalParameters(m, sigma, 0, d, r, eta_bound=25, l=l)
    lines = open(path).read().split()
    recs = [lines[i:i + 52] for i in range(0, len(lines) - 51, 52)]
    RP = REF.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, l=l) if REF else None
    for rec in recs:
        alpha = int(rec[0])
        with mp.workprec(rs.PRECISION):  # the reference's KAT forms theta_r at 192 bits
            theta = (2 * mp.pi / mp.ldexp(mp.mpf(1), m + sigma)) * alpha
        etas = range(-25, 26) if m <= 512 else (-25, -1, 0, 1, 25)
        errs = []
        for eta in etas:
            exp = rec[1 + eta + 25]
            v = rs.diagonal_probability_approx_f_eta(theta, eta, P)
            errs.append(_rel(v, exp))
            if RP is not None:
                assert _rel(REF.diagonal_probability_f_eta(RP, alpha, eta, 192), v) < 1e-45
        assert max(errs) < 1e-6 and sorted(errs)[len(errs) // 2] < 1e-25, max(errs)


# ---- restatement vs compiled reference on full slices ------------------------------

def test_restatement_matches_golden_slices():
    """oracle/restate.py reproduces the reference's slices (tests/golden) bit for bit
    up to the last long-double digit."""
    done = 0
    for g in golden_slices():
        k = g.meta
        if k["kind"] == "2d" and k["D"] <= 16 and not (k["method"] == 1 and k["m"] > 1024):
            P = rs.Parameters(k["m"], k["s"], g.d, g.r)
            f = rs.distribution_slice_compute_richardson if k["richardson"] else rs.distribution_slice_compute
            sl = f(P, k["D"], k["a_d"], k["a_r"], k["method"])
        elif k["kind"] == "lin" and k["D"] <= 64 and k["m"] <= 1024:
            P = rs.Parameters(k["m"], k["s"], g.d, g.r)
            f = (rs.linear_distribution_slice_compute_richardson if k["richardson"]
                 else rs.linear_distribution_slice_compute)
            sl = f(P, k["D"], k["a"], k["target"])
        else:
            continue
        assert cell_errors(sl.cells, g.cells) < 1e-17, g
        assert abs(float(sl.total_probability - g.total_probability)) < 1e-18
        assert sl.flags == g.flags
        done += 1
    assert done >= 4


@needs_ref
def test_restatement_matches_reference_small_slices():
    for (m, s) in ((128, 2), (2048, 1)):
        d, r = REF.deterministic_d_r(m)
        assert (d, r) == rs.deterministic_d_r(m)
        P, RP = rs.Parameters(m, s, d, r), REF.RefParameters(m, s, d, r)
        a = rs.distribution_slice_compute_richardson(P, 4, m + 1, m)
        b = REF.distribution_slice_compute(RP, 4, m + 1, m)
        assert cell_errors(a.cells, b.cells) < 1e-18 and a.flags == b.flags
        assert abs(float((a.total_error - b.total_error) / b.total_error)) < 1e-12
        for sigma, eta in ((5, 0), (3, -2)):
            DP = rs.DiagonalParameters(m, sigma, s, d, r, eta_bound=5)
            RDP = REF.RefDiagonalParameters(m, sigma, s, d, r, eta_bound=5)
            a = rs.diagonal_distribution_slice_compute_richardson(DP, 8, m - 1, eta)
            b = REF.diagonal_distribution_slice_compute(RDP, 8, m - 1, eta)
            assert cell_errors(a.cells, b.cells) < 1e-18


# ---- Mathematica slice totals (the reference's own slice-level golden values) ---------

TOTALS = json.load(open(os.path.join(GOLDEN, "mathematica_totals.json")))
LIN_OFFSETS = list(range(-5, 11))
DIAG_OFFSETS = list(range(-5, 4))
DIAG_ETAS = [0, 1, -1, 2, -2, 25, -25]


def _close(a, b, tol):
    a, b = float(a), float(b)
    return a > 0 and b > 0 and abs(a - b) / min(a, b) <= tol  # test_cmp_ld, src/test/test_common.cpp:75-93


@needs_ref
def test_reference_reproduces_mathematica_linear_totals():
    m = 128
    d, r = REF.deterministic_d_r(m)
    RP = REF.RefParameters(m, 1, d, r)
    for target, arr in ((0, TOTALS["linear"][0]), (1, TOTALS["linear"][1])):
        assert len(arr["values"]) == 16
        for off, exp in list(zip(LIN_OFFSETS, arr["values"]))[::3]:
            for sign in (1, -1):
                sl = REF.linear_distribution_slice_compute(RP, 2048, sign * (m + off), target)
                tol = 1e-4 if (target == 1 and off >= 10) else 1e-6
                assert _close(sl.total_probability, exp, tol), (target, off, sign)


@needs_ref
def test_reference_reproduces_mathematica_diagonal_totals():
    m, sigma = 128, 5
    d, r = REF.deterministic_d_r(m)
    RDP = REF.RefDiagonalParameters(m, sigma, 1, d, r, eta_bound=25)
    pos, neg = TOTALS["diagonal"][0]["values"], TOTALS["diagonal"][1]["values"]
    assert len(pos) == 63 and len(neg) == 63
    for i in range(0, 63, 5):
        off, eta = DIAG_OFFSETS[i % 9], DIAG_ETAS[i // 9]
        a = REF.diagonal_distribution_slice_compute(RDP, 2048, m + off, eta)
        b = REF.diagonal_distribution_slice_compute(RDP, 2048, -(m + off), eta)
        assert _close(a.total_probability, pos[i], 1e-6), (off, eta)
        assert _close(b.total_probability, neg[i], 1e-6), (off, eta)


def test_kernel_math_reproduces_mathematica_totals():
    """All 32 linear and 126 diagonal Mathematica totals through the kernels'
    mathematics (host twin), at the reference test's own dimension 2048."""
    from tests import hostsim as hs
    m = 128
    d, r = rs.deterministic_d_r(m)
    for target in (0, 1):
        vals = TOTALS["linear"][target]["values"]
        coords = [m + o for o in LIN_OFFSETS] + [-(m + o) for o in LIN_OFFSETS]
        _, tp, _ = hs.slice1d(m, m, 0, d, r, target, 1, 2048, coords)
        for i, off in enumerate(LIN_OFFSETS):
            tol = 1e-4 if (target == 1 and off >= 10) else 1e-6
            assert _close(tp[i], vals[i], tol) and _close(tp[16 + i], vals[i], tol), (target, off)
    pos, neg = TOTALS["diagonal"][0]["values"], TOTALS["diagonal"][1]["values"]
    coords, etas = [], []
    for i in range(63):
        coords.append(m + DIAG_OFFSETS[i % 9])
        etas.append(DIAG_ETAS[i // 9])
    _, tp_p, _ = hs.slice1d(m, m, 5, d, r, 2, 1, 2048, coords, etas)
    _, tp_n, _ = hs.slice1d(m, m, 5, d, r, 2, 1, 2048, [-c for c in coords], etas)
    for i in range(63):
        assert _close(tp_p[i], pos[i], 1e-6) and _close(tp_n[i], neg[i], 1e-6), i
