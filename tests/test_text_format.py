"""The slice TEXT format ("%.24Lg\\n" per value): SURVEY.md section 8(f) #1.

Bar: byte-exact (integer / byte work). Oracles, strongest first (oracle/text.py): the
reference's own *_slice_export functions, the libc call they make, and an exact-integer
restatement. CPU tests run textfmt.cuh through tests/hostsim; GPU tests go through the C ABI.
"""
import io
import os

import numpy as np
import pytest

from oracle import text as ot
from tests.conftest import GOLDEN, golden_slices, ref_or_none

TEXT = os.path.join(GOLDEN, "text")


def rand_ld(rng, n, emin, emax, denormal=False):
    """Random x87 values with exponent field in [emin, emax] and random sign."""
    mant = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    if denormal:
        mant &= np.uint64(2 ** 63 - 1)
        mant |= np.uint64(1)
        se = np.zeros(n, dtype=np.uint16)
    else:
        mant |= np.uint64(1 << 63)
        se = rng.integers(emin, emax + 1, size=n).astype(np.uint16)
    se |= rng.integers(0, 2, size=n).astype(np.uint16) << 15
    return ot.ld_from_fields(mant, se)


def value_sets(seed, n):
    rng = np.random.default_rng(seed)
    return {
        "probabilities": rand_ld(rng, n, 16383 - 400, 16383),       # [1e-120, 1)
        "full_range": rand_ld(rng, n, 1, 32766),
        "around_one": rand_ld(rng, n, 16383 - 90, 16383 + 90),      # both styles, true ties
        "denormals": rand_ld(rng, max(1, n // 8), 0, 0, denormal=True),
        "doubles": rng.standard_normal(n).astype(np.longdouble) *
                   (np.longdouble(10) ** rng.integers(-30, 5, size=n)),
    }


def adversarial():
    z = np.load(os.path.join(TEXT, "adversarial.npz"))
    return ot.ld_from_fields(z["mant"], z["se"]), open(os.path.join(TEXT, "adversarial.txt"), "rb").read()


# ------------------------------------------------------------------ oracle pinning (CPU)

def test_exact_restatement_matches_libc_and_golden():
    """Pin the oracle: exact-integer restatement == libc == the committed golden text."""
    v, gold = adversarial()
    assert ot.format_ld24(v) == gold, f"this libc ({ot.libc_version()}) prints differently"
    assert ot.format_ld24_exact_array(v) == gold
    for name, vals in value_sets(7, 300).items():
        assert ot.format_ld24_exact_array(vals) == ot.format_ld24(vals), name


def test_reference_exporter_is_the_libc_loop():
    """The reference's exporters (golden text made by them) = header + the libc loop."""
    for name, nhead in (("2d", 4), ("linear", 3), ("diagonal", 4)):
        z = np.load(os.path.join(TEXT, f"slice_{name}.npz"))
        gold = open(os.path.join(TEXT, f"slice_{name}.txt"), "rb").read()
        lines = gold.split(b"\n")
        body = b"\n".join(lines[nhead:])
        assert ot.format_ld24(ot.ld_from_fields(z["mant"], z["se"])) == body
    ref = ref_or_none()
    if ref is not None:   # and the live reference, if built here
        g = next(x for x in golden_slices() if x.meta["name"].startswith("2d/c2/"))
        t = ot.ref_slice_export(0, g.meta["D"], g.meta["a_d"], g.meta["a_r"], g.flags, g.cells,
                                g.total_error)
        assert t == open(os.path.join(TEXT, "slice_2d.txt"), "rb").read()


# ------------------------------------------------------------------ host logic (CPU)

def test_pow10_table_entries_are_exact_truncations():
    from qunundrum_b200 import host
    for k in list(range(-4936, 4954, 7)) + list(range(-100, 130)) + [-4936, 4953]:
        T, e2, exact = host.text_pow10(k)
        assert 2 ** 191 <= T < 2 ** 192
        # 10^k = (T + f) * 2^(e2 - 191), 0 <= f < 1
        if k >= 0:
            num, den = 10 ** k, 1
        else:
            num, den = 1, 10 ** (-k)
        sh = e2 - 191
        if sh >= 0:
            den <<= sh
        else:
            num <<= -sh
        q, r = divmod(num, den)
        assert q == T, k
        assert exact == (r == 0) == (0 <= k <= 82), k


def test_floor_log10_pow2_is_exact_over_the_long_double_range():
    from tests import hostsim as hs
    j, p10 = 0, 10        # 10^j <= 2^n < p10 = 10^(j+1)
    for n in range(0, 16600):
        while (1 << n) >= p10:
            j, p10 = j + 1, p10 * 10
        assert hs.floor_log10_pow2(n) == j, n
        if n:   # 2^-n is never a power of ten: floor(log10 2^-n) = -(floor(log10 2^n) + 1)
            assert hs.floor_log10_pow2(-n) == -(j + 1), -n


# ------------------------------------------------------------------ the formatter on CPU

def test_hostsim_formatter_matches_golden_and_libc():
    from tests import hostsim as hs
    v, gold = adversarial()
    got, n_exact = hs.text_format_ld(v)
    assert got == gold
    assert n_exact > 0          # the adversarial set holds true ties
    assert hs.text_format_ld(v, force_band=True)[0] == gold
    for name, vals in value_sets(11, 60000).items():
        assert hs.text_format_ld(vals)[0] == ot.format_ld24(vals), name
    for name, vals in value_sets(12, 4000).items():
        got, n_exact = hs.text_format_ld(vals, force_band=True)
        assert got == ot.format_ld24(vals), name
        assert n_exact == np.count_nonzero(np.isfinite(vals) & (vals != 0)), name


def test_hostsim_formatter_on_golden_slices():
    from tests import hostsim as hs
    for name, nhead in (("2d", 4), ("linear", 3), ("diagonal", 4)):
        z = np.load(os.path.join(TEXT, f"slice_{name}.npz"))
        gold = open(os.path.join(TEXT, f"slice_{name}.txt"), "rb").read()
        body = b"\n".join(gold.split(b"\n")[nhead:])
        assert hs.text_format_ld(ot.ld_from_fields(z["mant"], z["se"]))[0] == body


# ------------------------------------------------------------------ CUDA path (C ABI)

@pytest.mark.gpu
def test_gpu_formatter_matches_golden(gpu_ctx):
    v, gold = adversarial()
    assert gpu_ctx.text_format(v) == gold
    assert gpu_ctx.text_exact_count > 0
    gpu_ctx.text_set_force_exact(True)
    try:
        assert gpu_ctx.text_format(v) == gold
    finally:
        gpu_ctx.text_set_force_exact(False)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 255, 256, 257, 4097, 65537])
def test_gpu_formatter_ragged_sizes(gpu_ctx, n):
    vals = value_sets(100 + n, max(n, 1))["full_range"][:n]
    assert gpu_ctx.text_format(vals) == ot.format_ld24(vals)
    tail = np.longdouble("2.5e-307")
    assert gpu_ctx.text_format(vals, tail) == ot.format_ld24(np.append(vals, tail))


@pytest.mark.gpu
def test_gpu_formatter_random_sets(gpu_ctx):
    for name, vals in value_sets(21, 400000).items():
        assert gpu_ctx.text_format(vals) == ot.format_ld24(vals), name
    gpu_ctx.text_set_force_exact(True)
    try:
        for name, vals in value_sets(22, 20000).items():
            assert gpu_ctx.text_format(vals) == ot.format_ld24(vals), name
    finally:
        gpu_ctx.text_set_force_exact(False)


@pytest.mark.gpu
def test_gpu_formatter_doubles(gpu_ctx):
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2 ** 64, size=300000, dtype=np.uint64)
    d = bits.view(np.float64)
    d = np.concatenate([d, [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 1.0, 0.5,
                            2.2250738585072014e-308, 1.7976931348623157e308]])
    assert gpu_ctx.text_format(d) == ot.format_ld24(d.astype(np.longdouble))


@pytest.mark.gpu
def test_gpu_slice_exporters_match_the_reference(gpu_ctx):
    import qunundrum_b200 as qb
    from qunundrum_b200 import host
    for name, kind in (("2d", 0), ("linear", 1), ("diagonal", 2)):
        z = np.load(os.path.join(TEXT, f"slice_{name}.npz"))
        gold = open(os.path.join(TEXT, f"slice_{name}.txt"), "rb").read()
        vals = ot.ld_from_fields(z["mant"], z["se"])
        D, c0, c1, flags = (int(x) for x in z["head"])
        f = io.BytesIO()
        if kind == 0:
            s = qb.Distribution_Slice(D, c0, c1, flags=flags, norm_matrix=vals[:-1],
                                      total_error=vals[-1])
            host.distribution_slice_export(s, f, gpu_ctx)
        elif kind == 1:
            s = qb.Linear_Distribution_Slice(D, c0, flags=flags, norm_vector=vals[:-1],
                                             total_error=vals[-1])
            host.linear_distribution_slice_export(s, f, gpu_ctx)
        else:
            s = qb.Diagonal_Distribution_Slice(D, c0, c1, flags=flags, norm_vector=vals[:-1],
                                               total_error=vals[-1])
            host.diagonal_distribution_slice_export(s, f, gpu_ctx)
        assert f.getvalue() == gold, name


@pytest.mark.gpu
def test_gpu_formatter_full_size_properties(gpu_ctx):
    """At the size of a stored 2D distribution chunk (8M values): line count, and
    text(concat) == concat(text) -- the output is position independent and ordered."""
    rng = np.random.default_rng(9)
    a = rand_ld(rng, 1 << 22, 16383 - 300, 16383)
    b = rand_ld(rng, (1 << 22) - 77, 16383 - 300, 16383)
    ta, tb = gpu_ctx.text_format(a), gpu_ctx.text_format(b)
    tab = gpu_ctx.text_format(np.concatenate([a, b]))
    assert tab == ta + tb
    assert tab.count(b"\n") == a.size + b.size
    # every line parses back to the very same long double (24 digits > 21 needed)
    back = ot.parse_ld(ta[:ta.find(b"\n", 3_000_000) + 1], ta[:ta.find(b"\n", 3_000_000) + 1].count(b"\n"))
    assert np.array_equal(back, a[:back.size])
