"""The exact samplers ("missing" #3):
sample_alpha_from_region, sample_j_from_alpha_r, sample_j_k_from_alpha_d, sample_j_k_from_alpha_d_r and
sample_j_from_diagonal_alpha_r (src/sample.cpp:78-410), and the diagonal distribution's sample from its
random bytes to k on the device (diagonal_distribution_sample_pair_j_k after the region is chosen,
src/diagonal_distribution.cpp:474-552).

Oracle: the compiled reference (oracle/_ref: the reference's own sample.cpp over its own random.c and
keccak) called on the same Random_State stream -- every integer must be the reference's bit for bit and
the stream position afterwards the same -- and Python integers / mpmath for the pieces (the table of
2^(i/D), the inverses, the division).

CPU tests run the very same __host__ __device__ code through tests/hostsim; GPU tests call the C ABI.
"""
import sys

import mpmath as mp
import numpy as np
import pytest

from oracle import restate as rs
from tests import hostsim as hs
from tests.conftest import ref_or_none

sys.set_int_max_str_digits(0)

REF = ref_or_none()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference)")


def big(g, bits):
    return int.from_bytes(g.bytes((bits + 7) // 8), "little") >> ((-bits) % 8)


def d_r_with_kappa(g, m, kd, kr):
    d = ((big(g, m - 1) | (1 << (m - 2))) >> kd << kd) | (1 << kd)
    r = ((big(g, m) | (1 << (m - 1))) >> kr << kr) | (1 << kr)
    return d, r


def mpz_bytes(bits):
    return (bits + 64 + 8) // 8  # src/random.c:163-164


def twin_factory(kind, m, l, sigma, d, r, dimension_max, emax=0):
    return hs.Exact(kind, m, l, sigma, d, r, dimension_max, emax)


def gpu_factory(ctx):
    from qunundrum_b200 import host

    def make(kind, m, l, sigma, d, r, dimension_max, emax=0):
        if kind == 1:
            p = host.Diagonal_Parameters(m=m, sigma=sigma, s=0, d=d, r=r, l=l)
        else:
            p = host.Parameters(m=m, s=0, d=d, r=r, l=l)
        return host.ExactSampler(p, dimension_max, emax, ctx)
    return make


# ---- pieces (no reference needed) -----------------------------------------------------------

def test_table_is_the_floor_of_two_to_the_fraction():
    d, r = rs.deterministic_d_r(128)
    ex = hs.Exact(0, 128, 128, 0, d, r, 256)
    T = ex.table()
    with mp.workprec(ex.P + 200):
        worst = max(abs(int(mp.floor(mp.mpf(2) ** (mp.mpf(i) / 256) * mp.mpf(2) ** ex.P)) - T[i])
                    for i in range(256))
    assert worst <= 2, worst
    assert T[0] == 1 << ex.P
    # a large one: m = 2048, 2048 regions (the diagonal distribution's configuration), sampled
    d, r = rs.deterministic_d_r(2048)
    ex = hs.Exact(1, 2048, 2048, 5, d, r, 2048)
    assert ex.emax == 2053 and ex.wn == 65
    T = ex.table()
    with mp.workprec(ex.P + 200):
        for i in (1, 2, 777, 1024, 2047):
            assert abs(int(mp.floor(mp.mpf(2) ** (mp.mpf(i) / 2048) * mp.mpf(2) ** ex.P)) - T[i]) <= 2


def test_inverses_and_kappa():
    g = np.random.default_rng(3)
    for m, l, kd, kr in ((128, 128, 0, 0), (96, 48, 5, 0), (200, 100, 0, 9), (128, 16, 64, 33)):
        d, r = d_r_with_kappa(g, m, kd, kr)
        ex = hs.Exact(0, m, l, 0, d, r, 64)
        assert (ex.kappa_d, ex.kappa_r) == (kd, kr)
        n = m + l
        assert (ex.inverse(0) * (r >> kr)) % (1 << n) == 1
        assert (ex.inverse(1) * (d >> kd)) % (1 << n) == 1
    with pytest.raises(ValueError):
        hs.Exact(0, 128, 128, 0, 1 << 65, 3, 64)  # kappa_d = 65
    with pytest.raises(ValueError):
        hs.Exact(0, 128, 128, 0, 5, 7, 100)  # dimension not a power of two


def test_division_against_python_integers():
    """exact_mod (Knuth D, 32-bit digits) on adversarial operands: quotient digits of 0 and 2^32 - 1,
    top limbs that force the qhat corrections and the add-back step."""
    import ctypes as C
    g = np.random.default_rng(4)
    L = hs.lib()
    cases = []
    specials = [0, 1, 0x7fffffff, 0x80000000, 0x80000001, 0xfffffffe, 0xffffffff]
    for _ in range(4000):
        wm = int(g.integers(2, 9))
        extra = int(g.integers(0, 4))
        M = [int(g.choice(specials)) if g.integers(3) == 0 else int(g.integers(0, 1 << 32)) for _ in range(wm)]
        if M[-1] == 0:
            M[-1] = int(g.choice(specials[1:]))
        nv = wm + extra
        V = [int(g.choice(specials)) if g.integers(3) == 0 else int(g.integers(0, 1 << 32)) for _ in range(nv)]
        cases.append((V, M))
    # the classical add-back trigger (TAOCP 4.3.1 exercise 22 scaled to 32-bit digits) and friends
    cases.append(([0, 0xfffffffe, 0x7fffffff, 0x80000000][::-1][::-1], [0xffffffff, 0, 0x80000000]))
    cases.append(([0, 0, 0x80000000, 0x7fffffff], [1, 0, 0x80000000]))
    cases.append(([0xffffffff] * 7, [0xffffffff, 0xffffffff, 0x80000000]))
    cases.append(([3, 0, 0x80000000, 0, 0], [1, 0, 0x80000000]))
    for V, M in cases:
        v = sum(x << (32 * i) for i, x in enumerate(V))
        mm = sum(x << (32 * i) for i, x in enumerate(M))
        a = np.array(V + [0], dtype=np.uint32)
        b = np.array(M + [0], dtype=np.uint32)
        L.hostsim_exact_mod(a.ctypes.data_as(C.c_void_p), C.c_uint32(len(V)), b.ctypes.data_as(C.c_void_p),
                            C.c_uint32(len(M)))
        got = sum(int(x) << (32 * i) for i, x in enumerate(a[:len(M)]))
        assert got == v % mm, (V, M)
        assert [int(x) for x in b[:len(M)]] == M  # the modulus is restored


# ---- alpha from a region -----------------------------------------------------------------------

def expected_alpha_python(e, region, D, sign, kappa, chunk: bytes):
    """sample_alpha_from_region with exact real arithmetic (the reference's 3 m-bit rounding of 2^x
    cannot change round(2^x) unless 2^x is within 2^(-2 e) of a half-integer)."""
    with mp.workprec(e + 400):
        lo = int(mp.nint(mp.mpf(2) ** (mp.mpf(e) + mp.mpf(region) / D)))
        hi = int(mp.nint(mp.mpf(2) ** (mp.mpf(e) + mp.mpf(region + 1) / D)))
    a = lo + int.from_bytes(chunk, "big") % (hi - lo)
    a -= a % (1 << kappa)
    return sign * a, mpz_bytes((hi - lo).bit_length())


def check_alpha_against_python(factory):
    g = np.random.default_rng(11)
    m = 160
    d, r = rs.deterministic_d_r(m)
    D = 64
    ex = factory(0, m, m, 0, d, r, D)
    stream = g.bytes(60000)
    regs, want, off = [], [], 0
    picks = [(64, 0), (64, D - 1), (ex.emax - 1, D - 1), (ex.emax - 1, 0), (100, 17)]
    for it in range(200):
        e, reg = picks[it] if it < len(picks) else (int(g.integers(64, ex.emax)), int(g.integers(0, D)))
        sign = -1 if g.integers(2) else 1
        # sub-dimensions share the table: D' divides D
        Dp = D >> int(g.integers(0, 3))
        reg %= Dp
        nb, st = ex.region_bytes(sign * e, reg, Dp)
        assert st == 0
        kappa = int(g.integers(0, 40)) if it % 3 == 0 else 0
        # bytes at the extremes now and then
        if it % 17 == 0:
            chunk = b"\xff" * nb
        elif it % 19 == 0:
            chunk = b"\x00" * nb
        else:
            chunk = stream[off:off + nb]
        stream = stream[:off] + chunk + stream[off + nb:]
        a, nb_want = expected_alpha_python(e, reg, Dp, sign, kappa, chunk)
        assert nb == nb_want
        regs.append((sign * e, reg, Dp, off, nb, kappa))
        want.append(a)
        off += nb
    for kappa in sorted(set(q[5] for q in regs)):
        idx = [i for i, q in enumerate(regs) if q[5] == kappa]
        got, st = ex.alpha([regs[i][:5] for i in idx], kappa, stream)
        assert list(st) == [0] * len(idx)
        assert got == [want[i] for i in idx], kappa
    # what is not a sample
    e = 100
    nb, _ = ex.region_bytes(e, 3, D)
    got, st = ex.alpha([(e, 3, D, 0, nb + 1), (e, 3, D, len(stream) - 1, nb), (7, 3, D, 0, nb),
                        (ex.emax, 3, D, 0, nb), (e, D, D, 0, nb), (e, 3, 48, 0, nb), (0, 0, D, 0, nb),
                        (e, 3, 2 * D, 0, nb)], 0, stream)
    assert list(st) == [1, 1, 3, 3, 3, 3, 3, 3]
    assert ex.region_bytes(7, 3, D) == (0, 3)


def check_alpha_against_reference(factory, m, D, count, kind=0, sigma=0, e_min=64):
    g = np.random.default_rng(m + D)
    d, r = rs.deterministic_d_r(m)
    ex = factory(kind, m, m, sigma, d, r, D)
    seed = bytes(g.integers(0, 256, 32, dtype=np.uint8))
    rng, twin = REF.RefRandom(seed), REF.RefRandom(seed)
    stream = twin.bytes(count * ((ex.emax + 80) // 8 + 8) + 64)
    regs, want, off = [], [], 0
    kappa = 0
    for it in range(count):
        e = int(g.integers(max(e_min, m - 60), ex.emax))
        sign = -1 if g.integers(2) else 1
        reg = D - 1 if it % 7 == 0 else int(g.integers(0, D))
        nb, st = ex.region_bytes(sign * e, reg, D)
        if st == 3 and e < 16:   # max = min: the reference divides by zero there
            continue
        assert st == 0
        want.append(REF.sample_alpha_from_region(sign * (e + reg / D), sign * (e + (reg + 1) / D), kappa, rng))
        regs.append((sign * e, reg, D, off, nb))
        off += nb
    got, st = ex.alpha(regs, kappa, stream)
    assert list(st) == [0] * len(regs)
    assert got == want
    assert rng.bytes(8) == stream[off:off + 8]  # the same stream position


# ---- (j, k) ---------------------------------------------------------------------------------------

def check_jk_against_reference(factory, m, s, kd, kr, count=40):
    g = np.random.default_rng(1000 * m + 10 * kd + kr)
    d, r = d_r_with_kappa(g, m, kd, kr)
    P = REF.RefParameters(m, s, d, r)
    l = P.l
    ex = factory(0, m, l, 0, d, r, 64)
    assert (ex.kappa_d, ex.kappa_r) == (kd, kr)
    top = min(m + 60, ex.emax)
    A_r = [(-1 if g.integers(2) else 1) * big(g, int(g.integers(1, top))) for _ in range(count)]
    A_d = [(-1 if g.integers(2) else 1) * big(g, int(g.integers(1, top))) for _ in range(count)]
    A_r[0], A_d[1], A_r[2], A_d[3] = 0, 0, -1, -1
    A_r[4] = A_d[4] = (1 << top) - 1
    A_r[5] = A_d[5] = -((1 << top) - 1)
    seed = bytes([kd + 1] * 32)
    for mode in (0, 1, 2):
        rng, twin = REF.RefRandom(seed), REF.RefRandom(seed)
        stream, off = twin.bytes(count * 700), 0
        want, ts, ks = [], [], []
        for i in range(count):
            t = kk = 0
            if mode == 1:  # src/sample.cpp:294-309
                kt = max(0, kr - kd - l)
                if kr > 0:
                    ln = mpz_bytes(kr - kt + 1)
                    t = (int.from_bytes(stream[off:off + ln], "big") % (1 << (kr - kt))) << kt
                    off += ln
            else:
                kap = kd if mode == 2 else kr
                if kap > 0:  # :176-180, :230-234
                    ln = mpz_bytes(kap + 1)
                    t = int.from_bytes(stream[off:off + ln], "big") % (1 << kap)
                    off += ln
            if mode == 2:  # :236-239
                ln = mpz_bytes(l + 1)
                kk = int.from_bytes(stream[off:off + ln], "big") % (1 << l)
                off += ln
            ts.append(t)
            ks.append(kk)
            want.append(REF.sample_j_k(mode, P, A_d[i], A_r[i], rng))
        assert rng.bytes(8) == stream[off:off + 8]
        if mode == 0:
            got = [(j, 0) for j in ex.j_from_alpha_r(A_r, ts)]
        elif mode == 1:
            got = list(zip(*ex.j_k_from_alpha_d_r(A_d, A_r, ts)))
        else:
            got = list(zip(ex.j_from_alpha_d_k(A_d, ks, ts), ks))
        assert got == want, (mode, m, s, kd, kr)


def check_diagonal_j_against_reference(factory, m, sigma, kr, count=40):
    g = np.random.default_rng(m + sigma + kr)
    d, r = d_r_with_kappa(g, m, 0, kr)
    P = REF.RefDiagonalParameters(m, sigma, 1, d, r)
    ex = factory(1, m, m, sigma, d, r, 64)
    assert ex.kappa_r == kr and ex.wk == 0
    top = m + sigma - 1
    A_r = [(-1 if g.integers(2) else 1) * big(g, int(g.integers(1, top))) for _ in range(count)]
    A_r[0], A_r[1] = 0, -1
    seed = bytes([9] * 32)
    rng, twin = REF.RefRandom(seed), REF.RefRandom(seed)
    stream, off = twin.bytes(count * 64), 0
    want, ts = [], []
    for i in range(count):
        t = 0
        if kr > 0:
            ln = mpz_bytes(kr + 1)
            t = int.from_bytes(stream[off:off + ln], "big") % (1 << kr)
            off += ln
        ts.append(t)
        want.append(REF.sample_j_k(3, P, None, A_r[i], rng)[0])
    assert ex.j_from_alpha_r(A_r, ts) == want


JK_CASES = [(128, 1, 0, 0), (128, 2, 3, 5), (96, 1, 0, 7), (160, 3, 33, 0), (128, 1, 2, 40), (128, 4, 1, 64),
            (128, 8, 0, 50)]


def test_alpha_matches_exact_arithmetic_on_the_cpu_twin():
    check_alpha_against_python(twin_factory)


@needs_ref
@pytest.mark.parametrize("m,D,count", [(128, 256, 300), (512, 1024, 60), (2048, 2048, 24)])
def test_alpha_matches_the_reference_on_the_cpu_twin(m, D, count):
    check_alpha_against_reference(twin_factory, m, D, count)


@needs_ref
def test_alpha_diagonal_parameters_on_the_cpu_twin():
    check_alpha_against_reference(twin_factory, 256, 512, 60, kind=1, sigma=5)


@needs_ref
@pytest.mark.parametrize("m,D,sigma", [(64, 128, 6), (40, 16, 4), (34, 1024, 3)])
def test_alpha_small_parameters_follow_both_roundings_of_the_reference(m, D, sigma):
    """|log alpha| < 31: the reference's 3 (e + 1)-bit rounding of 2^(e + i/D) is visible in the bound and
    is carried out step by step (exact_bound); regions whose bounds coincide are declined."""
    check_alpha_against_reference(twin_factory, m, D, 1500, kind=1, sigma=sigma, e_min=8)


@needs_ref
@pytest.mark.parametrize("m,s,kd,kr", JK_CASES)
def test_j_k_match_the_reference_on_the_cpu_twin(m, s, kd, kr):
    check_jk_against_reference(twin_factory, m, s, kd, kr)


@needs_ref
@pytest.mark.parametrize("m,sigma,kr", [(128, 5, 0), (160, 3, 6), (2048, 5, 0)])
def test_diagonal_j_matches_the_reference_on_the_cpu_twin(m, sigma, kr):
    check_diagonal_j_against_reference(twin_factory, m, sigma, kr, count=24 if m > 1000 else 40)


# ---- GPU: the same checks through the C ABI ------------------------------------------------------

@pytest.mark.gpu
def test_alpha_matches_exact_arithmetic_gpu(gpu_ctx):
    check_alpha_against_python(gpu_factory(gpu_ctx))


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("m,D,count", [(128, 256, 700), (512, 1024, 200), (2048, 2048, 40)])
def test_alpha_matches_the_reference_gpu(gpu_ctx, m, D, count):
    check_alpha_against_reference(gpu_factory(gpu_ctx), m, D, count)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("m,s,kd,kr", JK_CASES + [(2048, 1, 0, 0)])
def test_j_k_match_the_reference_gpu(gpu_ctx, m, s, kd, kr):
    check_jk_against_reference(gpu_factory(gpu_ctx), m, s, kd, kr, count=150 if m < 1000 else 40)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("m,sigma,kr", [(128, 5, 0), (160, 3, 6), (2048, 5, 0)])
def test_diagonal_j_matches_the_reference_gpu(gpu_ctx, m, sigma, kr):
    check_diagonal_j_against_reference(gpu_factory(gpu_ctx), m, sigma, kr, count=150 if m < 1000 else 40)


@pytest.mark.gpu
@needs_ref
def test_alpha_small_parameters_gpu(gpu_ctx):
    check_alpha_against_reference(gpu_factory(gpu_ctx), 64, 128, 1500, kind=1, sigma=6, e_min=8)
    check_alpha_against_reference(gpu_factory(gpu_ctx), 34, 1024, 1500, kind=1, sigma=3, e_min=8)


@pytest.mark.gpu
def test_gpu_and_twin_agree_on_a_large_batch(gpu_ctx):
    """Tile boundaries, several chunks' worth of tiles, m = 2048: the GPU against the CPU twin (which the
    tests above pin on the reference) on 5000 samples."""
    g = np.random.default_rng(77)
    m, sigma, D = 2048, 5, 2048
    d, r = rs.deterministic_d_r(m)
    tw = hs.Exact(1, m, m, sigma, d, r, D)
    gp = gpu_factory(gpu_ctx)(1, m, m, sigma, d, r, D)
    n = 5000
    stream = g.bytes(n * 280)
    regs, off = [], 0
    for _ in range(n):
        e = int(g.integers(m - 40, m + sigma - 1))
        sign = -1 if g.integers(2) else 1
        reg = int(g.integers(0, D))
        nb, st = tw.region_bytes(sign * e, reg, D)
        assert st == 0
        regs.append((sign * e, reg, D, off, nb))
        off += nb
    a_t, s_t = tw.alpha(regs, 0, stream)
    a_g, s_g = gp.alpha(regs, 0, stream)
    assert list(s_t) == list(s_g) == [0] * n
    assert a_t == a_g
    ts = [int(x) for x in g.integers(0, 1 << tw.kappa_r, n)]   # the deterministic r at m = 2048 is even
    assert tw.j_from_alpha_r(a_t, ts) == gp.j_from_alpha_r(a_g, ts)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("m,sigma,kr,count", [(128, 5, 0, 300), (160, 3, 6, 200), (2048, 5, 0, 24)])
def test_diagonal_sample_from_bytes_to_k_gpu(gpu_ctx, m, sigma, kr, count):
    """qb200_diagk_sample_drawn: region -> alpha_r -> j -> k without j leaving the device, against the
    reference's sample_alpha_from_region, sample_j_from_diagonal_alpha_r and
    sample_k_from_diagonal_j_eta_pivot called one after the other on the same Random_State."""
    from qunundrum_b200 import host
    g = np.random.default_rng(m + kr)
    d, r = d_r_with_kappa(g, m, 0, kr)
    D = 256
    P = REF.RefDiagonalParameters(m, sigma, 1, d, r)
    hp = host.Diagonal_Parameters(m=m, sigma=sigma, s=1, d=d, r=r)
    ex = host.ExactSampler(hp, D, 0, gpu_ctx)
    dk = host.DiagonalKSampler(hp, gpu_ctx)
    seed = bytes([m % 251] * 32)
    rng, twin = REF.RefRandom(seed), REF.RefRandom(seed)
    stream, off = twin.bytes(count * ((m + sigma + 80) // 8 + 64)), 0
    regs, ts, etas, pivots, want = [], [], [], [], []
    for i in range(count):
        e = int(g.integers(max(64, m - 20), m + sigma - 1))
        sign = -1 if g.integers(2) else 1
        reg = int(g.integers(0, D))
        eta = int(g.integers(-2, 3))
        nb, st = ex.region_bytes(sign * e, reg, D)
        assert st == 0
        alpha = REF.sample_alpha_from_region(sign * (e + reg / D), sign * (e + (reg + 1) / D), kr, rng)
        regs.append((sign * e, reg, D, off, nb))
        off += nb
        t = 0
        if kr > 0:
            ln = mpz_bytes(kr + 1)
            t = int.from_bytes(stream[off:off + ln], "big") % (1 << kr)
            off += ln
        ts.append(t)
        j = REF.sample_j_k(3, P, None, alpha, rng)[0]
        w = np.frombuffer(stream[off:off + 8], dtype="<u8")[0]
        off += 8
        pivot = np.longdouble(w) / np.longdouble(0xffffffffffffffff)  # random_generate_pivot_inclusive
        assert rng.bytes(8) == w.tobytes()
        ok, k, _, _ = REF.sample_k_from_diagonal_j_eta_pivot(P, pivot, j, eta, 1000)
        etas.append(eta)
        pivots.append(pivot)
        want.append((ok, k))
    ks, x, delta, status, est = host.diagonal_sample_drawn(dk, ex, regs, ts, stream, etas, pivots, 1000)
    assert list(est) == [0] * count
    got = [(st in (0, 2), k if st in (0, 2) else 0) for st, k in zip(status, ks)]
    assert got == want


def check_full_size_properties(factory, m=2048, s=1, kd=0, kr=3, D=256, n=6000, seed=5):
    """The properties the reference's own unit tests state (src/test/test_sample.cpp:97-570), at full size
    and without the reference in the loop: alpha lies in its region with the sign of the region and is
    divisible by 2^kappa; {r j} modulo 2^(m+l) (centred) gives alpha_r back; alpha_d - d j - 2^m k is, modulo
    2^(m+l), a non-negative number below 2^m (zero for an admissible pair); j < 2^(m+l), k < 2^l."""
    g = np.random.default_rng(seed)
    d, r = d_r_with_kappa(g, m, kd, kr)
    l = -(-m // s)
    ex = factory(0, m, l, 0, d, r, D)
    nmod = 1 << (m + l)
    stream = g.bytes(2 * n * ((ex.emax + 80) // 8 + 8))
    regs, off = [], 0
    for _ in range(2 * n):
        e = int(g.integers(m - 30, m + 10))
        sign = -1 if g.integers(2) else 1
        reg = int(g.integers(0, D))
        nb, st = ex.region_bytes(sign * e, reg, D)
        assert st == 0
        regs.append((sign * e, reg, D, off, nb))
        off += nb
    a_d, st_d = ex.alpha(regs[:n], kd, stream)
    a_r, st_r = ex.alpha(regs[n:], kr, stream)
    assert not st_d.any() and not st_r.any()
    bounds = {}

    def bound(e, i):
        if (e, i) not in bounds:
            with mp.workprec(e + 200):
                bounds[(e, i)] = int(mp.nint(mp.mpf(2) ** (mp.mpf(e) + mp.mpf(i) / D)))
        return bounds[(e, i)]

    for alphas, rg, kap in ((a_d, regs[:n], kd), (a_r, regs[n:], kr)):
        for a, (se, reg, _, _, _) in zip(alphas, rg):
            assert (a < 0) == (se < 0)
            lo, hi = bound(abs(se), reg), bound(abs(se), reg + 1)
            # clearing the low kappa bits can take alpha below min by less than 2^kappa, as in the reference
            assert lo - (1 << kap) < abs(a) < hi
            assert abs(a) % (1 << kap) == 0
    ts = [int(x) for x in g.integers(0, 1 << kr, n)] if kr else None
    js, ks = ex.j_k_from_alpha_d_r(a_d, a_r, ts)
    half = nmod >> 1
    for ad, ar, j, k in zip(a_d, a_r, js, ks):
        assert 0 <= j < nmod and 0 <= k < (1 << l)
        v = (r * j) % nmod
        assert (v - nmod if v >= half else v) == ar      # src/test/test_sample.cpp:519-527
        assert (ad - d * j - (k << m)) % nmod < (1 << m)  # :507-516 for a pair that need not be admissible


def test_full_size_properties_on_the_cpu_twin():
    check_full_size_properties(twin_factory, n=300)


@pytest.mark.gpu
def test_full_size_properties_gpu(gpu_ctx):
    check_full_size_properties(gpu_factory(gpu_ctx), n=6000)
    check_full_size_properties(gpu_factory(gpu_ctx), m=2048, s=8, kd=2, kr=0, D=2048, n=3000, seed=6)


# ---- the reference's own unit tests, restated on the mirror of its interface -------------------------
# src/test/test_sample.cpp:97-377 with the regions as a slice gives them (sixteenths of a unit of
# log alpha instead of the tenths the reference's test uses: a slice's dimension is a power of two).

def _centred(v, n):
    v %= (1 << n)
    return v - (1 << n) if v >= (1 << (n - 1)) else v   # mod_reduce / the explicit form of :339-343


def reference_unit_tests(make_sampler, span=30, kappas=(0, 3, 9)):
    from qunundrum_b200 import host as qb
    g = np.random.default_rng(2048)
    m, t = 2048, 30
    l = m // 8
    stream = qb.ByteStream(g.bytes(4 << 20))
    # test_sample_alpha_from_region (:97-179)
    d, r = d_r_with_kappa(g, m, 0, 0)
    P = qb.Parameters(m=m, s=0, d=d, r=r, l=l, t=t)
    S = make_sampler(P)
    for kappa in kappas:
        for i in range(m - span, m + span):
            for j in range(0, 16, 5):
                lo, hi = i + j / 16, i + (j + 1) / 16
                for sign in (1, -1):
                    alpha = qb.sample_alpha_from_region(sign * lo, sign * hi, kappa, stream, P, sampler=S)
                    assert (alpha > 0) == (sign > 0) and alpha != 0           # "Incorrect sign."
                    with mp.workprec(200):
                        assert lo - 1e-9 <= float(mp.log(abs(alpha), 2)) <= hi  # "Incorrect magnitude."
                    assert abs(alpha) % (1 << kappa) == 0                      # "Not divisible by 2^kappa."
    # calling with the bounds in the wrong order is a critical error (:176)
    with pytest.raises(qb.CriticalError):
        qb.sample_alpha_from_region(m + 0.5, m, 0, stream, P, sampler=S)
    with pytest.raises(qb.CriticalError):
        qb.sample_alpha_from_region(-m, m + 0.5, 0, stream, P, sampler=S)
    # test_sample_j_from_alpha_r (:181-264): d = r "by convention"; also an even r
    for kr in (0, 4):
        _, r = d_r_with_kappa(g, m, 0, kr)
        P = qb.Parameters(m=m, s=0, d=r, r=r, l=l, t=t)
        S = make_sampler(P)
        for i in range(m - span, m + span):
            for i2 in range(0, 16, 5):
                for sign in (1, -1):
                    alpha = qb.sample_alpha_from_region(sign * (i + i2 / 16), sign * (i + (i2 + 1) / 16), kr, stream, P,
                                                        sampler=S)
                    j = qb.sample_j_from_alpha_r(alpha, P, stream, sampler=S)
                    assert _centred(j * r, m + l) == alpha                     # "Failed to correctly sample j."
    # test_sample_j_k_from_alpha_d (:266-377)
    for kd in (0, 2):
        d, _ = d_r_with_kappa(g, m, kd, 0)
        d |= 1 << (m - 1)
        P = qb.Parameters(m=m, s=0, d=d, r=d, l=l, t=t)
        S = make_sampler(P)
        for i in range(m - span, m + span):
            for i2 in range(0, 16, 5):
                for sign in (1, -1):
                    alpha = qb.sample_alpha_from_region(sign * (i + i2 / 16), sign * (i + (i2 + 1) / 16), kd, stream, P,
                                                        sampler=S)
                    j, k = qb.sample_j_k_from_alpha_d(alpha, P, stream, sampler=S)
                    assert _centred(j * d + (k << m), m + l) == alpha          # "Failed to correctly sample (j, k)."
    return stream.pos


def test_the_references_unit_tests_on_the_cpu_twin():
    def make(P):
        return hs.Exact(0, P.m, P.l, 0, P.d, P.r, 16, P.m + 64)
    assert reference_unit_tests(make, span=6, kappas=(0, 9)) > 0


@pytest.mark.gpu
def test_the_references_unit_tests_gpu(gpu_ctx):
    from qunundrum_b200 import host as qb

    def make(P):
        return qb.ExactSampler(P, 16, 0, gpu_ctx)
    assert reference_unit_tests(make, span=30, kappas=(0, 3, 9)) > 0


def test_j_k_against_python_integers_over_random_sizes():
    """Every (j, k) mode against the formulas in Python integers for random m on [8, 300), l on [1, m], kappa up to
    59 and arguments of any size up to emax: operands shorter than a group of columns, n not a multiple of 32."""
    g = np.random.default_rng(99)
    done = 0
    for _ in range(150):
        m = int(g.integers(8, 300))
        l = int(g.integers(1, m + 1))
        kd = int(g.integers(0, min(60, m - 2))) if g.integers(3) == 0 else 0
        kr = int(g.integers(0, min(60, m - 2))) if g.integers(3) == 0 else 0
        d, r = d_r_with_kappa(g, m, kd, kr)
        ex = hs.Exact(0, m, l, 0, d, r, 4)
        n, N = m + l, 8
        A_r = [(-1 if g.integers(2) else 1) * big(g, int(g.integers(1, ex.emax + 1))) for _ in range(N)]
        A_d = [(-1 if g.integers(2) else 1) * big(g, int(g.integers(1, ex.emax + 1))) for _ in range(N)]
        inv_r, inv_d = pow(r >> kr, -1, 1 << n), pow(d >> kd, -1, 1 << n)
        tr = [big(g, kr) if kr else 0 for _ in range(N)]
        td = [big(g, kd) if kd else 0 for _ in range(N)]
        ks = [big(g, l) for _ in range(N)]
        want_j = [(((inv_r * a) >> kr) + (t << (n - kr))) % (1 << n) for a, t in zip(A_r, tr)]     # :185-204
        want_k = [((a - d * j) >> m) % (1 << l) for a, j in zip(A_d, want_j)]                       # :338-347
        want_j2 = [(inv_d * ((a - (k << m)) >> kd) + (t << (n - kd))) % (1 << n)
                   for a, k, t in zip(A_d, ks, td)]                                                 # :244-268
        assert ex.j_from_alpha_r(A_r, tr) == want_j
        assert ex.j_k_from_alpha_d_r(A_d, A_r, tr) == (want_j, want_k)
        assert ex.j_from_alpha_d_k(A_d, ks, td) == want_j2
        done += 1
    assert done == 150


# ---- the committed golden vectors (tests/golden/exact.json, made by tests/golden/make_exact_golden.py from
#      the reference): no reference needed at run time -------------------------------------------------

def _golden_cases():
    import json
    import os
    from tests.conftest import GOLDEN
    with open(os.path.join(GOLDEN, "exact.json")) as f:
        return json.load(f)


def check_golden(factory, case):
    ih = lambda s: int(s, 16)
    stream = bytes.fromhex(case["stream"])
    ex = factory(case["kind"], case["m"], case["l"], case["sigma"], ih(case["d"]), ih(case["r"]), case["dimension"])
    assert (ex.kappa_d, ex.kappa_r) == (case["kappa_d"], case["kappa_r"])
    D = case["dimension"]
    for which in ("d", "r"):
        recs = [s[which] for s in case["samples"]]
        regs = [(q["min_log_alpha"], q["region"], D, q["offset"], q["length"]) for q in recs]
        for q in recs:
            assert ex.region_bytes(q["min_log_alpha"], q["region"], D) == (q["length"], 0)   # the stream layout
        got, st = ex.alpha(regs, recs[0]["kappa"], stream)
        assert not st.any()
        assert got == [ih(q["alpha"]) for q in recs], (case["name"], which)
    a_d = [ih(s["d"]["alpha"]) for s in case["samples"]]
    a_r = [ih(s["r"]["alpha"]) for s in case["samples"]]
    S = case["samples"]
    if case["kind"] == 1:
        assert ex.j_from_alpha_r(a_r, [ih(s["t_r"]) for s in S]) == [ih(s["j_diagonal"]) for s in S]
        return
    assert ex.j_from_alpha_r(a_r, [ih(s["t_r"]) for s in S]) == [ih(s["j_from_alpha_r"]) for s in S]
    js, ks = ex.j_k_from_alpha_d_r(a_d, a_r, [ih(s["t_r_scaled"]) for s in S])
    assert [[format(j, "x"), format(k, "x")] for j, k in zip(js, ks)] == [s["j_k_from_alpha_d_r"] for s in S]
    assert ex.j_from_alpha_d_k(a_d, [ih(s["k_drawn"]) for s in S], [ih(s["t_d"]) for s in S]) == \
        [ih(s["j_from_alpha_d_k"]) for s in S]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_golden_vectors_on_the_cpu_twin(case):
    check_golden(twin_factory, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_golden_vectors_gpu(gpu_ctx, case):
    check_golden(gpu_factory(gpu_ctx), case)
