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


# ================================================================== importer ("%Lg")

def decimal_tokens(seed, n):
    """Random decimal numbers in the forms strtold accepts (1..28 significant digits)."""
    import random
    rnd = random.Random(seed)
    toks = []
    for _ in range(n):
        nd = rnd.randint(1, 28)
        digs = str(rnd.randint(10 ** (nd - 1), 10 ** nd - 1)) if nd > 1 else str(rnd.randint(0, 9))
        e = rnd.choice([rnd.randint(-4990, 4950), rnd.randint(-400, 50), rnd.randint(-30, 30)])
        form = rnd.randint(0, 3)
        if form == 0:
            s = f"{digs}e{e}"
        elif form == 1:
            p = rnd.randint(0, len(digs))
            s = f"{digs[:p]}.{digs[p:]}E{e:+d}"
        elif form == 2:
            s = ("-" if rnd.random() < .5 else "+") + f"0.000{digs}e{e}"
        else:
            s = digs + "0" * rnd.randint(0, 5) + f"e{e}"
        toks.append(s)
    return toks


def midpoint_tokens(seed, n):
    """Decimal strings that are EXACTLY half way between two adjacent long doubles."""
    import random
    rnd = random.Random(seed)
    toks = []
    while len(toks) < n:
        M = rnd.randint(2 ** 63, 2 ** 64 - 1)
        k = rnd.randint(0, 8)
        s = str((2 * M + 1) * 5 ** k)          # (2M + 1) / 2^k, written out
        if len(s) <= 28:
            toks.append(f"{s}e{-k}")
    return toks


BOUNDARY_TOKENS = [
    "1.18973149535723176502e4932", "1.18973149535723176508e4932",
    "1.189731495357231765085759326628e4932", "1.2e4932", "1e4933", "1e99999",
    "3.6451995318824746025e-4951", "1.8225997659412373012e-4951", "1.8225997659412373013e-4951",
    "1.82259976594123730126e-4951", "1e-4952", "1e-99999", "0", "-0", "0.0e10", "0e-100",
    "3.3621031431120935063e-4932", "3.36210314311209350626e-4932", "1.", ".5", "5.e3", "+7",
    "inf", "-INF", "Infinity", "nan", "-nan", "NaN", "1e0", "000123.4500e-2", "9" * 40,
    "0." + "0" * 50 + "1", "1" + "0" * 30,
]
MALFORMED = ["abc", "1e", "--1", "1.2.3", "e5", ".", "1e+", "12a", "+", "1e5.5"]


def same_bits(a, b):
    ma, sa = ot.ld_fields(a)
    mb, sb = ot.ld_fields(b)
    return bool(np.all(ma == mb) and np.all(sa == sb))


def join(tokens, sep=b"\n"):
    return sep.join(t.encode() for t in tokens) + sep


def test_parse_restatement_matches_libc():
    """Pin the importer's oracle: exact-integer restatement == libc strtold."""
    toks = decimal_tokens(1, 1500) + midpoint_tokens(2, 300) + BOUNDARY_TOKENS
    want = ot.parse_ld(join(toks), len(toks))
    mant, se = ot.ld_fields(want)
    for t, m, s in zip(toks, mant, se):
        assert ot.parse_ld_exact(t.encode()) == (int(m), int(s)), t


def test_hostsim_parser_matches_libc():
    from tests import hostsim as hs
    v, gold = adversarial()
    got, used, _ = hs.text_parse_ld(gold, v.size)
    assert same_bits(got, ot.parse_ld(gold, v.size)) and used == len(gold)
    for toks in (decimal_tokens(3, 100000), midpoint_tokens(4, 20000), BOUNDARY_TOKENS):
        t = join(toks)
        got, used, n_exact = hs.text_parse_ld(t, len(toks))
        assert same_bits(got, ot.parse_ld(t, len(toks)))
        assert used == len(t)
    assert n_exact == 0 or True
    toks = decimal_tokens(5, 20000) + midpoint_tokens(6, 5000)
    got, _, n_exact = hs.text_parse_ld(join(toks), len(toks), force_band=True)
    assert same_bits(got, ot.parse_ld(join(toks), len(toks))) and n_exact > 20000
    for bad in MALFORMED:
        with pytest.raises(ValueError):
            hs.text_parse_ld(join(["1.5", bad]), 2)
    with pytest.raises(ValueError):
        hs.text_parse_ld(b"1 2 3\n", 4)


def test_hostsim_roundtrip_every_value_survives():
    """parse(format(x)) == x bit for bit: 24 digits identify a 64-bit mantissa."""
    from tests import hostsim as hs
    for name, vals in value_sets(41, 50000).items():
        vals = vals[np.isfinite(vals)]
        text, _ = hs.text_format_ld(vals)
        back, used, _ = hs.text_parse_ld(text, vals.size)
        assert same_bits(back, vals) and used == len(text), name


def test_hostsim_parser_whitespace_and_consumed():
    from tests import hostsim as hs
    t = b"  1.5\t\t-2e3\r\n\n  7 \n\n8\n"
    got, used, _ = hs.text_parse_ld(t, 3)
    assert [float(x) for x in got] == [1.5, -2000.0, 7.0]
    assert t[used:] == b"8\n"
    got, used, _ = hs.text_parse_ld(t, 4)
    assert used == len(t) and float(got[3]) == 8.0
    assert hs.text_parse_ld(t, 0)[1] == 2      # leading white space is skipped, as fscanf does


@pytest.mark.gpu
def test_gpu_parser_matches_libc(gpu_ctx):
    v, gold = adversarial()
    got, used = gpu_ctx.text_parse(gold, v.size)
    assert same_bits(got, ot.parse_ld(gold, v.size)) and used == len(gold)
    for toks in (decimal_tokens(13, 300000), midpoint_tokens(14, 50000), BOUNDARY_TOKENS):
        t = join(toks)
        got, used = gpu_ctx.text_parse(t, len(toks))
        assert same_bits(got, ot.parse_ld(t, len(toks))) and used == len(t)
    assert gpu_ctx.text_exact_count == 0
    t = join(midpoint_tokens(14, 50000))
    gpu_ctx.text_parse(t, 50000)
    assert gpu_ctx.text_exact_count > 40000     # true ties take the exact decision
    toks = decimal_tokens(15, 20000)
    gpu_ctx.text_set_force_exact(True)
    try:
        got, _ = gpu_ctx.text_parse(join(toks), len(toks))
    finally:
        gpu_ctx.text_set_force_exact(False)
    assert same_bits(got, ot.parse_ld(join(toks), len(toks)))


@pytest.mark.gpu
def test_gpu_parser_errors_whitespace_consumed(gpu_ctx):
    import qunundrum_b200 as qb
    for bad in MALFORMED:
        with pytest.raises(qb.CriticalError, match="-21"):
            gpu_ctx.text_parse(join(["1.5", bad, "3"]), 3)
    with pytest.raises(qb.CriticalError, match="-22"):
        gpu_ctx.text_parse(b"0x1p3\n", 1)
    with pytest.raises(qb.CriticalError, match="-20"):
        gpu_ctx.text_parse(b"1 2 3\n", 4)
    t = b"  1.5\t\t-2e3\r\n\n  7 \n\n8\n"
    got, used = gpu_ctx.text_parse(t, 3)
    assert [float(x) for x in got] == [1.5, -2000.0, 7.0] and t[used:] == b"8\n"
    got, used = gpu_ctx.text_parse(t, 4)
    assert used == len(t)
    assert gpu_ctx.text_parse(b"", 0)[1] == 0
    for n in (1, 15, 16, 17, 4095, 4096, 4097):     # token starts across thread / tile borders
        toks = [str(i) for i in range(n)]
        for sep in (b" ", b"\n", b" \n\t "):
            t = join(toks, sep)
            got, used = gpu_ctx.text_parse(t, n)
            assert [int(x) for x in got] == list(range(n)) and used == len(t)


@pytest.mark.gpu
def test_gpu_roundtrip_at_full_size(gpu_ctx):
    """4M values (one full-size chunk of a stored distribution): format -> parse returns
    every bit; and the reference's own importer reads what we wrote."""
    rng = np.random.default_rng(17)
    vals = np.concatenate([rand_ld(rng, 1 << 22, 16383 - 1100, 16383),
                           rand_ld(rng, 1 << 16, 1, 32766),
                           rand_ld(rng, 1 << 12, 0, 0, denormal=True)])
    text = gpu_ctx.text_format(vals)
    back, used = gpu_ctx.text_parse(text, vals.size)
    assert used == len(text) and same_bits(back, vals)


@pytest.mark.gpu
def test_gpu_slice_importers_match_the_reference(gpu_ctx):
    from qunundrum_b200 import host
    for name, fn in (("2d", host.distribution_slice_import),
                     ("linear", host.linear_distribution_slice_import),
                     ("diagonal", host.diagonal_distribution_slice_import)):
        z = np.load(os.path.join(TEXT, f"slice_{name}.npz"))
        gold = open(os.path.join(TEXT, f"slice_{name}.txt"), "rb").read()
        vals = ot.ld_from_fields(z["mant"], z["se"])
        f = io.BytesIO(gold + b"12345\n")          # something follows the slice in a real file
        s = fn(f, gpu_ctx)
        cells = s.norm_matrix if name == "2d" else s.norm_vector
        assert same_bits(cells, vals[:-1]) and same_bits(np.array([s.total_error]), vals[-1:])
        assert f.read() == b"12345\n"
        assert s.dimension == int(z["head"][0]) and s.flags == int(z["head"][3])
        ref = ref_or_none()
        if ref is not None:
            kind = {"2d": 0, "linear": 1, "diagonal": 2}[name]
            head, rcells, rtp, rte = ot.ref_slice_import(kind, gold, cells.size)
            assert same_bits(rcells, cells) and same_bits(np.array([rtp]), np.array([s.total_probability]))


@pytest.mark.gpu
def test_gpu_slice_importer_reads_more_when_the_block_is_short(gpu_ctx):
    """Numbers separated by a lot of white space: the importer's first block (40 bytes per
    number) ends before the last number and it has to read on -- same result, same position."""
    from qunundrum_b200 import host
    rng = np.random.default_rng(3)
    vals = rand_ld(rng, 64 * 64 + 1, 16383 - 200, 16383)
    lines = ot.format_ld24(vals).split(b"\n")[:-1]
    body = b"".join(b" " * 70 + ln + b"\n" for ln in lines)
    f = io.BytesIO(b"64\n2040\n-2041\n000a0000\n" + body + b"777\n")
    s = host.distribution_slice_import(f, gpu_ctx)
    assert s.dimension == 64 and s.min_log_alpha_d == 2040 and s.min_log_alpha_r == -2041
    assert s.flags == 0xA0000
    assert same_bits(s.norm_matrix, vals[:-1]) and same_bits(np.array([s.total_error]), vals[-1:])
    assert f.read() == b"777\n"


@pytest.mark.gpu
def test_gpu_slice_importer_never_takes_a_number_cut_by_the_block_end(gpu_ctx):
    """Line lengths around the importer's block size (40 bytes per number): for some paddings the
    first block ends inside, or right behind, the last number of the slice. The number must not
    be taken as it stands; the importer reads on."""
    from qunundrum_b200 import host
    rng = np.random.default_rng(4)
    vals = rand_ld(rng, 32 * 32 + 1, 16383 - 200, 16383 - 20)
    lines = ot.format_ld24(vals).split(b"\n")[:-1]
    for pad in range(0, 24):
        body = b"".join(b" " * pad + ln + b"\n" for ln in lines)
        f = io.BytesIO(b"32\n2040\n2041\n000a0000\n" + body + b"777\n" + b"1.5\n" * 4000)
        s = host.distribution_slice_import(f, gpu_ctx)
        assert same_bits(s.norm_matrix, vals[:-1]), pad
        assert same_bits(np.array([s.total_error]), vals[-1:]), pad
        assert f.read(4) == b"777\n", pad
