"""The diagonal k sampler (second half):
sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412-646) over
diagonal_probability_approx_h (src/diagonal_probability.cpp:99-162), and the sum of
tau_estimate_diagonal (src/tau_estimate.cpp:135-210).

Oracle, in this order of authority:
  * the reference's known-answer vectors res/test-vectors/sample-k-from-diagonal-j-eta-pivot-*
    and diagonal-probabilities-h-det-* (src/test/test_sample.cpp:675-838,
    src/test/test_diagonal_probability.cpp:174-300): a sample of 9 + 9 files is committed under
    tests/golden/kat; where /root/reference is present ALL 522 sample-k files are run;
  * the compiled reference (oracle/_ref) on seeded random inputs and the edge cases of the
    domain, committed as tests/golden/diagk.npz (tests/golden/make_diagk_golden.py);
  * oracle/restate.py, the mpmath restatement, pinned on the same vectors.

The bar: k identical to the reference's (every bit), the same success flag, alpha_phi to
1e-25 relative (the product carries it as a double-double; the vectors themselves carry it at
2 l bits, so small l limits what they can show).

CPU tests run the very same __host__ __device__ code through tests/hostsim; GPU tests call the
C ABI.
"""
import glob
import math
import os
import re
import sys

import mpmath as mp
import numpy as np
import pytest

from oracle import restate as rs
from tests import hostsim as hs
from tests.conftest import GOLDEN, ref_or_none

sys.set_int_max_str_digits(0)

LD = np.longdouble
KAT = os.path.join(GOLDEN, "kat")
TV = "/root/reference/res/test-vectors"
REF = ref_or_none()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference)")

K_FILES = sorted(glob.glob(os.path.join(KAT, "sample-k-from-diagonal-j-eta-pivot-m-*.txt")))
H_FILES = sorted(glob.glob(os.path.join(KAT, "diagonal-probabilities-h-det-m-*.txt")))


def _msl(path):
    m, sigma, s = map(int, re.search(r"m-(\d+)-sigma-(\d+)-s-(\d+)", path).groups())
    return m, sigma, int(math.ceil(m / s))  # src/test/test_sample.cpp:713


def k_records(path):
    """(j, eta, pivot, k, alpha_phi) x 25, src/test/test_sample.cpp:764-790."""
    L = open(path).read().split()
    return [(int(L[i]), int(L[i + 1]), LD(L[i + 2]), int(L[i + 3]), L[i + 4])
            for i in range(0, len(L) - 4, 5)]


def _alpha_errors(path, recs, x):
    """Relative error of x = alpha_phi / 2^(m+sigma-l) against the vectors."""
    m, sigma, l = _msl(path)
    errs = []
    with mp.workprec(400):
        for q, xi in zip(recs, x):
            want = mp.mpf(q[4]) / mp.mpf(2) ** (m + sigma - l)
            got = mp.mpf(float(xi[0])) + mp.mpf(float(xi[1]))
            errs.append(float(abs(got - want) / abs(want)) if want != 0 else float(abs(got)))
    return errs


def _alpha_tolerance(l):
    # the vectors carry alpha_phi at 2 l bits (src/test/test_sample.cpp:717)
    return max(1e-25, 2.0 ** (3 - 2 * l))


def check_k_file(path, sampler_factory, exact_too=True):
    """Both ways of walking: decided in doubles with an error band (the default), and the exact
    x87 walk for every sample."""
    m, sigma, l = _msl(path)
    d, r = rs.deterministic_d_r(m)
    recs = k_records(path)
    assert len(recs) == 25
    S = sampler_factory(m, sigma, l, d, r)
    for force_exact in ((False, True) if exact_too else (False,)):
        S.set_force_exact(force_exact)
        ks, x, delta, st = S.sample([q[0] for q in recs], [q[1] for q in recs], [q[2] for q in recs])
        assert list(st) == [0] * 25
        assert ks == [q[3] for q in recs], os.path.basename(path)
        assert max(_alpha_errors(path, recs, x)) < _alpha_tolerance(l)


# ---- pieces -------------------------------------------------------------------------

def test_sinpi_acc_is_double_double_accurate():
    rng = np.random.default_rng(5)
    ts = list(rng.uniform(-0.5, 0.5, 300)) + [0.25, -0.25, 0.5, -0.5, 1e-200, 2.0 ** -60, 0.2500000001]
    worst = 0.0
    with mp.workprec(300):
        for t in ts:
            lo = float(rng.uniform(-1, 1)) * abs(t) * 2.0 ** -54 if 1e-200 < abs(t) < 0.49 else 0.0
            got = hs.sinpi_acc(float(t), lo)
            want = mp.sin(mp.pi * (mp.mpf(float(t)) + mp.mpf(lo)))
            worst = max(worst, float(abs((mp.mpf(got[0]) + mp.mpf(got[1])) / want - 1)))
    assert worst < 1e-31, worst


def test_x87_from_dd_rounds_like_the_fpu():
    rng = np.random.default_rng(6)
    cases = []
    for _ in range(4000):
        hi = float(rng.uniform(0.5, 1.0)) * 2.0 ** int(rng.integers(-200, 10))
        lo = float(rng.uniform(-0.5, 0.5)) * np.spacing(hi) * 2.0 ** -int(rng.integers(0, 30))
        cases.append((hi, lo))
    one = 1.0
    for hi in (one, 2.0 ** -40, np.nextafter(2.0, 0.0), np.nextafter(one, 2.0)):
        u = np.spacing(hi)
        for lo in (0.0, u / 2, -u / 4, -u / 2 ** 11, u / 2 ** 12, -u / 2 ** 12, u / 2 ** 12 * (1 + 2.0 ** -30),
                   u / 2 ** 12 * (1 - 2.0 ** -30), -u / 2 ** 80, u / 2 ** 80, 3 * u / 2 ** 12, -u / 2):
            if hi + lo == hi:
                cases.append((hi, lo))
    for hi, lo in cases:
        want = LD(hi) + LD(lo)  # one rounding to the 64-bit significand
        got = hs.x87_from_dd(hi, lo)
        assert got == want, (hi, lo, got, want)
    assert hs.x87_from_dd(0.0, 0.0) == 0 and hs.x87_from_dd(-1.0, 0.0) == 0


# ---- the oracle on the reference's vectors -----------------------------------------------

@pytest.mark.parametrize("path", [p for p in K_FILES if _msl(p)[0] <= 512], ids=os.path.basename)
def test_restatement_reproduces_the_k_vectors(path):
    m, sigma, l = _msl(path)
    d, r = rs.deterministic_d_r(m)
    P = rs.DiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
    for j, eta, pivot, k, alpha in k_records(path)[:10]:
        ok, got_k, got_alpha = rs.sample_k_from_diagonal_j_eta_pivot(P, pivot, j, eta, 0xffffffff)
        assert ok and got_k == k
        with mp.workprec(400):
            assert abs(got_alpha - mp.mpf(alpha)) <= abs(mp.mpf(alpha)) * mp.mpf(2) ** (3 - 2 * l)


@needs_ref
@pytest.mark.parametrize("path", K_FILES[:5], ids=os.path.basename)
def test_compiled_reference_reproduces_the_k_vectors(path):
    m, sigma, l = _msl(path)
    d, r = rs.deterministic_d_r(m)
    P = REF.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
    for j, eta, pivot, k, _ in k_records(path)[:8]:
        ok, got_k, _, _ = REF.sample_k_from_diagonal_j_eta_pivot(P, pivot, j, eta)
        assert ok and got_k == k


# ---- the device code on the CPU (tests/hostsim) -------------------------------------------

@pytest.mark.parametrize("path", K_FILES, ids=os.path.basename)
def test_twin_k_vectors(path):
    check_k_file(path, hs.DiagK)


@pytest.mark.skipif(not os.path.isdir(TV), reason="needs /root/reference")
def test_twin_all_522_k_vector_files():
    files = sorted(glob.glob(os.path.join(TV, "sample-k-from-diagonal-j-eta-pivot-m-*.txt")))
    assert len(files) == 522
    for n, path in enumerate(files):
        check_k_file(path, hs.DiagK, exact_too=(n % 7 == 0))


def h_records(path):
    L = open(path).read().split()
    return [(L[i], L[i + 1]) for i in range(0, len(L) - 1, 2)]


def h_inputs(path):
    """x = phi 2^l / (2 pi) as (hi, lo) rows, and the expected h as long doubles."""
    m, sigma, l = _msl(path)
    xs, want = [], []
    with mp.workprec(3 * l + 400):
        for phi, h in h_records(path):
            x = mp.mpf(phi) * mp.mpf(2) ** l / (2 * mp.pi)
            hi = float(x)
            xs.append((hi, float(x - mp.mpf(hi))))
            want.append(rs._get_ld(mp.mpf(h)))
    return l, np.array(xs), np.array(want, dtype=LD)


def check_h(l, got, want):
    # the reference's own tolerance is 1e-6 (test_cmp_ld, src/test/test_common.cpp:75-93)
    assert np.all(got > 0)
    rel = np.abs(got - want) / np.minimum(got, want)
    assert float(rel.max()) < 1e-15, float(rel.max())


@pytest.mark.parametrize("path", H_FILES, ids=os.path.basename)
def test_twin_h_vectors(path):
    l, xs, want = h_inputs(path)
    assert len(xs) == 25
    check_h(l, hs.diagk_h(l, xs), want)


class DiagGold:
    def __init__(self, z, name):
        g = lambda k: z[f"{name}_{k}"]  # noqa: E731
        self.name = name
        self.m, self.sigma, self.l = [int(v) for v in g("params")]
        self.d = int.from_bytes(g("d").tobytes(), "big")
        self.r = int.from_bytes(g("r").tobytes(), "big")
        self.J, self.eta, self.pivot, self.bound = g("j"), g("eta"), g("pivot"), g("bound")
        self.K, self.ok, self.alpha = g("k"), g("ok"), g("alpha")


def diag_gold():
    z = np.load(os.path.join(GOLDEN, "diagk.npz"))
    return [DiagGold(z, str(n)) for n in z["names"]]


GOLD = diag_gold()


def check_gold(g, S):
    """S.sample(js, etas, pivots, delta_bound) against the reference's outputs, for both ways of
    walking."""
    for force_exact in (False, True):
        S.set_force_exact(force_exact)
        _check_gold(g, S)
    S.set_force_exact(False)


def _check_gold(g, S):
    js = [hs.limbs_to_int(row) for row in g.J]
    for bound in sorted(set(int(b) for b in g.bound)):
        idx = [i for i in range(len(js)) if int(g.bound[i]) == bound]
        ks, x, delta, st = S.sample([js[i] for i in idx], g.eta[idx], g.pivot[idx], bound)
        for n, i in enumerate(idx):
            assert (st[n] in (0, 2)) == bool(g.ok[i]), (g.name, i)
            assert st[n] in (0, 1, 2)
            assert ks[n] == hs.limbs_to_int(g.K[i]), (g.name, i, js[i], int(g.eta[i]))
            got = LD(x[n, 0]) + LD(x[n, 1])
            if st[n] == 2:  # negative unreduced phi: j < |eta| 2^(m+sigma) / r (src/sample.cpp:566-574)
                assert js[i] < 26 * (1 << (g.m + g.sigma)) // g.r + 2
                got -= LD(2) ** g.l
            want = g.alpha[i]
            assert abs(got - want) <= abs(want) * LD(2) ** -62, (g.name, i, got, want)
            assert abs(delta[n]) <= bound


@pytest.mark.parametrize("g", GOLD, ids=lambda g: g.name)
def test_twin_matches_the_reference_on_random_and_edge_inputs(g):
    check_gold(g, hs.DiagK(g.m, g.sigma, g.l, g.d, g.r))


@needs_ref
def test_twin_matches_the_live_reference_on_fresh_seeds():
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    import random
    prng = random.Random(int(rng.integers(1 << 30)))
    for m, sigma, l in ((128, 3, 64), (160, 0, 20), (512, 7, 300), (2048, 5, 1027)):
        r = (1 << (m - 1)) + 1 + prng.randrange((1 << (m - 1)) - 1)
        d = r // 2 + prng.randrange(r // 2)
        P = REF.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
        S = hs.DiagK(m, sigma, l, d, r)
        js = [prng.randrange(1 << (m + sigma)) for _ in range(12)]
        etas = [prng.randrange(-25, 26) for _ in js]
        piv = np.array([LD(prng.random()) * LD(0.98) for _ in js], dtype=LD)
        ks, x, _, st = S.sample(js, etas, piv, 4000)
        for i, j in enumerate(js):
            ok, k, a, _ = REF.sample_k_from_diagonal_j_eta_pivot(P, piv[i], j, etas[i], 4000, precision=256)
            assert ok == (st[i] == 0) and k == ks[i], (m, sigma, l, d, r, j, etas[i], piv[i])
            assert abs((LD(x[i, 0]) + LD(x[i, 1])) - a) <= abs(a) * LD(2) ** -62


@pytest.mark.parametrize("m,sigma,l", [(256, 4, 256), (128, 5, 20)])
def test_twin_long_walks_decided_in_doubles_equal_the_exact_walk(m, sigma, l):
    """Pivots next to 1 walk thousands of steps (for l = 20 around the whole ring of 2^20 values of
    k): the pass in doubles with its error band must stop where the exact x87 walk stops."""
    import random
    prng = random.Random(m + l)
    r = (1 << (m - 1)) + 1 + prng.randrange((1 << (m - 1)) - 1)
    d = r // 2 + prng.randrange(r // 2)
    n = 40
    js = [prng.randrange(1 << (m + sigma)) for _ in range(n)]
    etas = [prng.randrange(-25, 26) for _ in range(n)]
    piv = np.array([LD(1) - LD(prng.random()) * LD(10.0 ** -prng.uniform(2.5, 4.5)) for _ in range(n)], dtype=LD)
    S = hs.DiagK(m, sigma, l, d, r)
    out = []
    for exact in (False, True):
        S.set_force_exact(exact)
        out.append(S.sample(js, etas, piv, 30000))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.abs(out[0][2]).max() > 2000 and (out[0][3] == 0).sum() >= n // 2


def _exact_sample(m, sigma, l, d, r, j, eta, pivot, delta_bound):
    """The scale-free form in exact integers and 200-bit floats (the derivation in
    qunundrum_b200/csrc/diagk.cuh; itself checked against the reference's vectors in
    test_exact_form_reproduces_the_k_vectors): (ok, k, x)."""
    n = m + sigma
    Z = r * j
    q = (Z >> n) + (1 if (Z & ((1 << n) - 1)) >= (1 << (n - 1)) else 0)
    s = (q + eta) % r
    w = (d * s) % r
    Qv, w2 = divmod(w << l, r)
    c = 1 if 2 * w2 >= r else 0
    k0 = (-(Qv + c)) % (1 << l)
    with mp.workprec(200):
        t = mp.mpf(w2 - c * r) / r
        S = mp.sin(mp.pi * t) ** 2
        p = rs._get_ld(mp.mpf(0)) + pivot
        for idx in range(2 * delta_bound + 1):
            delta = (idx + 1) // 2 if idx % 2 else -(idx // 2)
            x = t + delta
            if l < 62:
                dm = delta % (1 << l)
                if dm >= (1 << (l - 1)):
                    dm -= 1 << l
                x = t + dm
                if x >= (1 << (l - 1)):
                    x -= 1 << l
                if x < -(1 << (l - 1)):
                    x += 1 << l
            h = mp.mpf(1) if x == 0 else S / (mp.mpf(2) ** l * mp.sin(mp.pi * x / mp.mpf(2) ** l)) ** 2
            p = p - rs._get_ld(h)
            if p <= 0:
                return True, (k0 + delta) % (1 << l), x
    return False, 0, mp.mpf(0)


def test_exact_form_reproduces_the_k_vectors():
    path = [p for p in K_FILES if "m-512-" in p][0]
    m, sigma, l = _msl(path)
    d, r = rs.deterministic_d_r(m)
    for j, eta, pivot, k, _ in k_records(path)[:12]:
        ok, got, _ = _exact_sample(m, sigma, l, d, r, j, eta, pivot, 5000)
        assert ok and got == k


def test_twin_integer_part_on_odd_shapes():
    """Sizes the vectors do not have: r of 1 .. 9 limbs with any number of top bits, r much shorter
    than 2^m and than 2^l (the shift by l then takes several Barrett rounds), m + sigma and l not
    multiples of 32, l from 1 to m + sigma, d tiny and d = r - 1, j at both ends of its range."""
    import random
    prng = random.Random(99)
    checked = 0
    for trial in range(160):
        rbits = prng.choice([33, 40, 63, 64, 65, 95, 96, 97, 128, 130, 191, 200, 257, 288])
        m = rbits + prng.choice([0, 0, 1, 7, 40, 150])
        sigma = prng.choice([0, 1, 5, 11, 31, 32])
        l = prng.choice([1, 2, 13, 31, 32, 33, 61, 62, 63, 64, 96, 109, 110, 111, m, m + sigma])
        l = max(1, min(l, m + sigma))
        r = (1 << (rbits - 1)) + prng.randrange(1 << (rbits - 1))
        d = prng.choice([1, 2, r - 1, r // 2, 1 + prng.randrange(r - 1)])
        S = hs.DiagK(m, sigma, l, d, r)
        n = m + sigma
        js = [0, 1, (1 << n) - 1, (1 << n) - 2, 1 << (n - 1)] + [prng.randrange(1 << n) for _ in range(5)]
        etas = [prng.randrange(-25, 26) for _ in js]
        piv = np.array([LD(prng.random()) * LD(0.6) for _ in js], dtype=LD)
        ks, x, delta, st = S.sample(js, etas, piv, 40)
        for i, j in enumerate(js):
            ok, k, xe = _exact_sample(m, sigma, l, d, r, j, etas[i], piv[i], 40)
            assert ok == (st[i] in (0, 2)), (m, sigma, l, d, r, j, etas[i])
            assert ks[i] == k, (m, sigma, l, d, r, j, etas[i], piv[i])
            if ok:
                with mp.workprec(200):
                    got = mp.mpf(float(x[i, 0])) + mp.mpf(float(x[i, 1]))
                    assert abs(got - xe) <= abs(xe) * mp.mpf(2) ** -95 + mp.mpf(2) ** -1000, (m, sigma, l, d, r, j)
            checked += 1
    assert checked == 1600


def test_twin_skipped_columns_of_r_j():
    """The product r j starts three guard columns below the first column that is needed and is
    formed in full when the third guard comes out as 0xffffffff (diagk_fraction). Both paths against
    the exact integers: random inputs with and without the switch that disables the short cut, and
    inputs CONSTRUCTED so that the guard is all ones (one sample in 2^32 otherwise)."""
    import os
    import random
    prng = random.Random(4242)
    m, sigma, l = 256, 0, 256
    n = m + sigma
    first = (n >> 5) - 4                       # first column of the shortened product
    cases = []
    for trial in range(12):
        r = (1 << (m - 1)) + prng.randrange(1 << (m - 1))
        j0 = prng.randrange(1 << 32) | 1
        if trial < 8:
            # j = j0 (one limb): the partial product is j0 * (r >> 32 first); make its third limb all ones
            target = (0xFFFFFFFF << 64) | prng.randrange(1 << 64)
            X = (target * pow(j0, -1, 1 << 96)) % (1 << 96)
            mask = ((1 << 96) - 1) << (32 * first)
            r = (r & ~mask) | (X << (32 * first))
            assert ((j0 * (r >> (32 * first))) >> 64) & 0xFFFFFFFF == 0xFFFFFFFF
            j = j0
        else:
            j = prng.randrange(1 << n)
        d = 1 + prng.randrange(r - 1)
        cases.append((d, r, j, prng.randrange(-25, 26), LD(prng.random()) * LD(0.6)))
    for full in ("0", "1"):
        os.environ["QB200_DIAGK_FULL_PRODUCT"] = full
        try:
            for d, r, j, eta, piv in cases:
                S = hs.DiagK(m, sigma, l, d, r)
                ks, x, delta, st = S.sample([j], [eta], np.array([piv], dtype=LD), 40)
                ok, k, xe = _exact_sample(m, sigma, l, d, r, j, eta, piv, 40)
                assert ok == (st[0] in (0, 2)) and ks[0] == k, (full, d, r, j, eta)
        finally:
            del os.environ["QB200_DIAGK_FULL_PRODUCT"]


def test_twin_fixed_point_fraction_equals_the_exact_one():
    """diagk_fraction_fixed_point (one truncated product with 2^l d / r in fixed point) against
    diagk_fraction_exact (s rho, Barrett division, low bits of s D'; QB200_DIAGK_EXACT_FRACTION=1
    sends every sample there) and against the exact integers: k identical, x to 2^-95 of itself --
    over odd shapes, and on inputs where the fixed-point path must decline (s = 0: t = 0)."""
    import os
    import random
    prng = random.Random(777)
    cases = []
    for trial in range(60):
        rbits = prng.choice([128, 130, 191, 200, 257, 288, 512])
        m = rbits + prng.choice([0, 1, 7, 40])
        sigma = prng.choice([0, 1, 5, 31, 32])
        l = max(1, min(prng.choice([1, 13, 32, 33, 64, 109, 110, 111, m, m + sigma]), m + sigma))
        r = (1 << (rbits - 1)) + prng.randrange(1 << (rbits - 1))
        d = prng.choice([1, r - 1, r // 2, 1 + prng.randrange(r - 1)])
        n = m + sigma
        js = [0, 1, (1 << n) - 1, 1 << (n - 1)] + [prng.randrange(1 << n) for _ in range(4)]
        etas = [0, 0, 1, -1] + [prng.randrange(-25, 26) for _ in range(4)]
        piv = np.array([LD(prng.random()) * LD(0.6) for _ in js], dtype=LD)
        cases.append((m, sigma, l, d, r, js, etas, piv))
    got = {}
    for exact in ("0", "1"):
        os.environ["QB200_DIAGK_EXACT_FRACTION"] = exact
        try:
            for ci, (m, sigma, l, d, r, js, etas, piv) in enumerate(cases):
                S = hs.DiagK(m, sigma, l, d, r)
                ks, x, delta, st = S.sample(js, etas, piv, 40)
                got[(exact, ci)] = (ks, x.copy(), delta.copy(), st.copy())
                for i, j in enumerate(js):
                    ok, k, xe = _exact_sample(m, sigma, l, d, r, j, etas[i], piv[i], 40)
                    assert ok == (st[i] in (0, 2)) and ks[i] == k, (exact, m, sigma, l, d, r, j, etas[i])
                    if ok:
                        with mp.workprec(200):
                            v = mp.mpf(float(x[i, 0])) + mp.mpf(float(x[i, 1]))
                            assert abs(v - xe) <= abs(xe) * mp.mpf(2) ** -95 + mp.mpf(2) ** -1000, (exact, m, l, j)
        finally:
            del os.environ["QB200_DIAGK_EXACT_FRACTION"]
    for ci in range(len(cases)):
        a, b = got[("0", ci)], got[("1", ci)]
        assert a[0] == b[0] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_twin_refuses_bad_parameters():
    with pytest.raises(ValueError):
        hs.DiagK(128, 0, 64, 5, 3)          # d >= r
    with pytest.raises(ValueError):
        hs.DiagK(128, 0, 200, 3, 5)         # l > m + sigma
    with pytest.raises(ValueError):
        hs.DiagK(64, 0, 32, 3, (1 << 70) + 1)  # r >= 2^m


# ---- the C ABI on the GPU ------------------------------------------------------------------

def _gpu_factory(gpu_ctx):
    import qunundrum_b200 as qb

    def make(m, sigma, l, d, r):
        return qb.DiagonalKSampler(qb.Diagonal_Parameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l), gpu_ctx)
    return make


@pytest.mark.gpu
@pytest.mark.parametrize("path", K_FILES, ids=os.path.basename)
def test_gpu_k_vectors(path, gpu_ctx):
    check_k_file(path, _gpu_factory(gpu_ctx))


@pytest.mark.gpu
@pytest.mark.parametrize("path", H_FILES, ids=os.path.basename)
def test_gpu_h_vectors(path, gpu_ctx):
    l, xs, want = h_inputs(path)
    m, sigma, _ = _msl(path)
    d, r = rs.deterministic_d_r(m)
    S = _gpu_factory(gpu_ctx)(m, sigma, l, d, r)
    got = S.approx_h(xs)
    check_h(l, got, want)
    assert np.array_equal(got, hs.diagk_h(l, xs))


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLD, ids=lambda g: g.name)
def test_gpu_matches_the_reference_on_random_and_edge_inputs(g, gpu_ctx):
    check_gold(g, _gpu_factory(gpu_ctx)(g.m, g.sigma, g.l, g.d, g.r))


def _random_batch(m, sigma, n, seed):
    import random
    prng = random.Random(seed)
    r = (1 << (m - 1)) + 1 + prng.randrange((1 << (m - 1)) - 1)
    d = r // 2 + prng.randrange(r // 2)
    wj = (m + sigma + 31) // 32
    rng = np.random.default_rng(seed)
    J = rng.integers(0, 1 << 32, size=(n, wj), dtype=np.uint64).astype(np.uint32)
    if (m + sigma) % 32:
        J[:, -1] &= np.uint32((1 << ((m + sigma) % 32)) - 1)
    eta = rng.integers(-25, 26, size=n).astype(np.int32)
    piv = rng.random(n).astype(LD) * LD(0.995)
    return d, r, J, eta, piv


@pytest.mark.gpu
@pytest.mark.parametrize("m,sigma,l,n", [(2048, 5, 2048, 3000), (128, 2, 40, 400000), (1023, 0, 1003, 2000)])
def test_gpu_equals_the_twin_on_a_batch(m, sigma, l, n, gpu_ctx):
    """Same code on both sides: k, delta and status identical, x to the last bits (the host
    compiler may contract a multiply-add the device build keeps apart). 400000 samples span
    more than one launch."""
    d, r, J, eta, piv = _random_batch(m, sigma, n, 77 + m)
    S = _gpu_factory(gpu_ctx)(m, sigma, l, d, r)
    ks, x, delta, st = S.sample(J, eta, piv, 3000)
    sub = np.arange(n) if n <= 3000 else np.random.default_rng(1).choice(n, 3000, replace=False)
    T = hs.DiagK(m, sigma, l, d, r)
    ks2, x2, delta2, st2 = T.sample([hs.limbs_to_int(J[i]) for i in sub], eta[sub], piv[sub], 3000)
    assert [ks[i] for i in sub] == ks2
    assert np.array_equal(delta[sub], delta2) and np.array_equal(st[sub], st2)
    xa = x[sub, 0].astype(LD) + x[sub, 1].astype(LD)
    xb = x2[:, 0].astype(LD) + x2[:, 1].astype(LD)
    assert np.all(np.abs(xa - xb) <= np.abs(xb) * LD(2) ** -60)
    assert (st == 0).mean() > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("m,sigma,l", [(2048, 5, 2048), (128, 5, 20), (256, 0, 100)])
def test_gpu_long_walks_equal_the_twin(m, sigma, l, gpu_ctx):
    """Pivots next to 1: the walks that the kernel takes thirty-two steps at a time across the warp
    (several lanes of a warp at once, and lanes that run out of bounds) against the sequential twin,
    and against its own exact walk."""
    n = 600
    d, r, J, eta, piv = _random_batch(m, sigma, n, 1234 + l)
    rng = np.random.default_rng(l)
    far = rng.random(n) < 0.3
    piv = np.where(far, LD(1) - (rng.random(n) * 10.0 ** -rng.uniform(2.0, 5.0, n)).astype(LD), piv)
    S = _gpu_factory(gpu_ctx)(m, sigma, l, d, r)
    ks, x, delta, st = S.sample(J, eta, piv, 20000)
    T = hs.DiagK(m, sigma, l, d, r)
    ks2, x2, delta2, st2 = T.sample([hs.limbs_to_int(J[i]) for i in range(n)], eta, piv, 20000)
    assert ks == ks2 and np.array_equal(delta, delta2) and np.array_equal(st, st2)
    S.set_force_exact(True)
    sub = np.where(far)[0][:40]
    ks3, x3, delta3, st3 = S.sample(J[sub], eta[sub], piv[sub], 20000)
    assert ks3 == [ks[i] for i in sub] and np.array_equal(delta3, delta[sub]) and np.array_equal(st3, st[sub])
    assert np.abs(delta).max() > 1000 and (st == 1).any() and (st == 0).sum() > n // 2


@pytest.mark.gpu
def test_gpu_tau_estimate(gpu_ctx):
    """tau = log2(mean alpha_phi^2) / 2 - (m + sigma - l) per estimate; an estimate with an
    out-of-bounds sample or |eta| > eta_bound gives DBL_MAX and FALSE (src/tau_estimate.cpp:163-201)."""
    m, sigma, l, n, count = 2048, 3, 2048, 16, 64
    d, r, J, eta, piv = _random_batch(m, sigma, n * count, 5)
    piv[7 * n + 3] = LD(1)          # runs out of bounds with delta_bound = 50
    eta[9 * n + 1] = 26             # above eta_bound
    S = _gpu_factory(gpu_ctx)(m, sigma, l, d, r)
    tau, ok = S.tau_estimate(n, count, J, eta, piv, 50, 25)
    _, x, _, st = S.sample(J, eta, piv, 50, want_k=False)
    for t in range(count):
        sl = slice(t * n, (t + 1) * n)
        good = np.all(st[sl] == 0) and np.all(np.abs(eta[sl]) <= 25)
        assert ok[t] == good
        if not good:
            assert tau[t] == LD(np.finfo(np.float64).max)
            continue
        with mp.workprec(192):
            acc = mp.mpf(0)
            for hi, lo in x[sl]:
                a = mp.mpf(float(hi)) + mp.mpf(float(lo))
                acc += a * a
            want = mp.log(acc / n, 2) / 2
        assert abs(float(tau[t]) - float(want)) < 1e-15 and abs(tau[t] - rs._get_ld(want)) <= LD(2) ** -60
    assert not ok[7] and not ok[9] and ok.sum() >= count - 8


@pytest.mark.gpu
def test_gpu_pivot_out_of_bounds_is_fatal(gpu_ctx):
    import qunundrum_b200 as qb
    d, r, J, eta, piv = _random_batch(128, 0, 4, 9)
    S = _gpu_factory(gpu_ctx)(128, 0, 128, d, r)
    piv[2] = LD(1.5)
    with pytest.raises(qb.CriticalError, match="pivot is out of bounds"):
        S.sample(J, eta, piv, 10)
    with pytest.raises(qb.CriticalError):
        qb.DiagonalKSampler(qb.Diagonal_Parameters(128, 0, 0, 5, 3, l=64), gpu_ctx)


# ---- tau_estimate_diagonal: the drop-in next to the reference's own, in one process -----------

import json  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IB = os.path.join(ROOT, "integration", "_build")


def _generate_diagonal(flavour, cwd, args, np_=5):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    env = dict(os.environ, QB200_DEVICE="0", QB200_TEXT_DEVICE="0")
    p = subprocess.run([os.path.join(IB, "minimpirun"), "-np", str(np_),
                        os.path.join(IB, flavour, "generate_diagonal_distribution"), *args],
                       cwd=cwd, env=env, capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    files = os.listdir(os.path.join(cwd, "distributions"))
    assert len(files) == 1
    return os.path.join(cwd, "distributions", files[0])


def _check(exe, dist, n, estimates, delta_bound, eta_bound, seed, batch=1):
    """batch = 1: the Random_States are compared after every call; batch > 1: the drop-in computes
    `batch` estimates per GPU call and the states are compared at every batch boundary."""
    p = subprocess.run([exe, dist, str(n), str(estimates), str(delta_bound), str(eta_bound), str(seed),
                        str(batch)],
                       capture_output=True, text=True, timeout=1800,
                       env=dict(os.environ, QB200_DEVICE="0", QB200_TAU_BATCH=str(batch)))
    assert p.stdout.strip(), p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert p.returncode == 0 and out["ok"], out
    assert out["mismatched_flags"] == 0 and out["mismatched_taus"] == 0 and out["mismatched_states"] == 0
    return out


# (n, estimates, delta_bound, eta_bound): plain; a delta bound that pivots outrun and an eta bound
# below the distribution's (failing samples in mid-estimate: the stream is put back and replayed);
# delta_bound 0
TAU_RUNS = [(4, 300, 1000, 2), (3, 200, 1, 1), (16, 150, 0, 2)]   # estimates: multiples of the batches used


@pytest.mark.skipif(not os.path.exists(os.path.join(IB, "obj", "tau_diagonal_check.o")),
                    reason="integration/_build missing (needs /root/reference at build time)")
def test_tau_diagonal_dropin_host_logic_on_the_cpu_shim():
    """qunundrum_b200/dropin/dropin_tau_diagonal.cpp over the CPU stand-in of qb200_diagk_*
    (tests/hostsim/abi_shim.cpp): the same success flags, tau and -- after every call -- the same
    Random_State as the reference's tau_estimate_diagonal on identically seeded generators."""
    from tests.hostsim import shim_flavour as sf
    exe = sf.build_tau_diagonal()
    assert exe
    with tempfile.TemporaryDirectory() as t:
        dist = _generate_diagonal("ref", t, ["-dim", "128", "-eta-bound", "2", "-det", "128", "3", "1"], np_=9)
        failed = 0
        for seed, (n, est, db, eb) in enumerate(TAU_RUNS):
            for batch in (1, 50):
                out = _check(exe, dist, n, est, db, eb, seed + 1, batch)
                assert out["worst_tau_difference"] <= 2.0 ** -58
                failed += out["failed_estimates"]
        assert failed > 40          # the replay path was taken
    with tempfile.TemporaryDirectory() as t:
        # m = 64: slices at |log alpha_r| from 34, where the bounds of a region carry the reference's own
        # rounding of 2^x to 3 (e + 1) bits (exact.cuh, exact_bound)
        dist = _generate_diagonal("ref", t, ["-dim", "128", "-eta-bound", "2", "-det", "64", "6", "2"], np_=9)
        for seed, (n, est, db, eb) in enumerate([(2, 300, 1000, 2), (3, 200, 1, 1)]):
            for batch in (1, 50):
                out = _check(exe, dist, n, est, db, eb, seed + 1, batch)
                assert out["worst_tau_difference"] <= 2.0 ** -58


@pytest.mark.gpu
@pytest.mark.parametrize("args,runs", [
    (["-dim", "128", "-eta-bound", "2", "-det", "128", "3", "1"], TAU_RUNS),
    (["-dim", "256", "-eta-bound", "2", "-det", "2048", "5", "1"], [(8, 40, 1000, 2), (3, 60, 1, 1)]),
    (["-dim", "256", "-eta-bound", "1", "-det", "1024", "2", "4"], [(5, 60, 1000, 1)]),
], ids=["m128", "m2048", "m1024_s4"])
def test_gpu_tau_diagonal_dropin_equals_the_reference_in_process(args, runs):
    exe = os.path.join(IB, "gpu", "tau_diagonal_check")
    if not os.path.exists(exe):
        pytest.skip("integration/_build missing (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as t:
        dist = _generate_diagonal("gpu", t, args)
        for seed, (n, est, db, eb) in enumerate(runs):
            _check(exe, dist, n, 20, db, eb, seed + 7, 1)
            out = _check(exe, dist, n, est, db, eb, seed + 1, 20)
            assert out["worst_tau_difference"] <= 2.0 ** -58
            print(f"\n{os.path.basename(dist)} n={n}: reference {out['reference_s']:.3f} s, "
                  f"drop-in {out['dropin_s']:.3f} s for {est} estimates")


# ---- the reference's own known-answer test function, against the drop-in ---------------------------

@pytest.mark.skipif(not os.path.isdir(TV) or not os.path.exists(os.path.join(IB, "obj", "math.o")),
                    reason="needs /root/reference and integration/_build")
def test_the_references_own_kat_test_passes_with_the_dropin_sampler():
    """test_sample_k_from_diagonal_j_eta_pivot_kat() of src/test/test_sample.cpp:679-836 (all 522 files,
    mpz_cmp on k, test_cmp_ld on alpha_phi) and test_diagonal_probability_h_approx_kat() of
    src/test/test_diagonal_probability.cpp:174-298 (522 files), compiled in place and linked against
    sample_k_from_diagonal_j_eta_pivot and diagonal_probability_approx_h of
    qunundrum_b200/dropin/dropin_tau_diagonal.cpp (-DQB200_DROPIN_SAMPLE_K) over the CPU stand-in of
    the library -- and, with the same driver, against
    the reference's own sample.cpp (the reference's test_cmp_ld refuses the negative alpha_phi of half
    of the records, so the unmodified test fails with the reference's own sampler as well; the driver,
    integration/tools/sample_k_kat_check.cpp, says what it supplies instead)."""
    from tests.hostsim import shim_flavour as sf
    for flavour in ("dropin", "reference"):
        exe = sf.build_sample_k_kat(flavour=flavour)
        assert exe
        p = subprocess.run([exe], cwd="/root/reference", capture_output=True, text=True, timeout=1500)
        assert p.returncode == 0 and p.stdout.strip().endswith("ok"), (flavour, p.stdout[-500:], p.stderr[-500:])
        assert p.stdout.count("Processing:") == 1044    # 522 files of k vectors + 522 of h vectors
