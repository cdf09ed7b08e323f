"""Parity of the CUDA path (through the C ABI) with the reference -- needs a B200.

 * every cell of every committed golden slice (tests/golden/slices.npz, produced by
   the UNMODIFIED reference): <= 1e-9 relative per cell, <= 1e-12 on the slice's
   probability mass, total_error <= 1e-9 relative, flags exact;
 * the same against the reference itself run on this box when oracle/_ref travelled;
 * the fused kernel against the plain one-thread-per-cell kernels;
 * at BASELINE.json's full size (the 3362-slice m = 2048 set at dimension 128):
   determinism, batch-composition invariance, Richardson identity, captured mass.
"""
import numpy as np
import pytest

import qunundrum_b200 as qb
from qunundrum_b200 import shard
from tests.conftest import golden_slices, ref_or_none
from tests.util import CELL_RTOL, assert_slice_matches, cell_errors, group_by

pytestmark = pytest.mark.gpu

G = golden_slices()
G2D = group_by([g for g in G if g.meta["kind"] == "2d"], ["m", "s", "l", "D", "richardson", "method", "d", "r"])
GLIN = group_by([g for g in G if g.meta["kind"] == "lin"], ["m", "s", "l", "D", "richardson", "target", "d", "r"])
GDIAG = group_by([g for g in G if g.meta["kind"] == "diag"], ["m", "s", "l", "sigma", "D", "richardson", "d", "r"])


def _run_plan(plan, algo):
    import torch
    plan.set_algorithm(algo)
    cells = torch.empty(max(1, plan.cells), dtype=torch.float64, device="cuda")
    summ = torch.empty(max(1, plan.n * 8), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan.run(cells.data_ptr(), summ.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    tp, te, fl = plan.finish(summ.cpu().numpy()[:plan.n * 8])
    return cells.cpu().numpy()[:plan.cells].reshape(plan.n, -1), tp, te, fl


def test_extension_is_loaded_and_device_is_blackwell(gpu_ctx):
    assert qb.lib().qb200_device_count() >= 1
    assert gpu_ctx.measure_fp64_peak() > 1e12


@pytest.mark.parametrize("key", list(G2D), ids=lambda k: f"m{k[0]}-s{k[1]}-D{k[3]}-R{k[4]}-meth{k[5]}")
def test_2d_golden(gpu_ctx, key):
    m, s, l, D, rich, method, d, r = key
    gs = G2D[key]
    P = qb.Parameters(m, s, int(d), int(r), l=0 if s else l)
    cells, tp, te, fl = gpu_ctx.slice2d_batch(P, method, bool(rich), D,
                                              [g.meta["a_d"] for g in gs], [g.meta["a_r"] for g in gs])
    for i, g in enumerate(gs):
        assert_slice_matches(cells[i], tp[i], te[i], fl[i], g)


@pytest.mark.parametrize("key", [k for k in G2D if k[4] == 1 and k[3] % 32 == 0 and k[5] != 1],
                         ids=lambda k: f"m{k[0]}-s{k[1]}-D{k[3]}-meth{k[5]}")
def test_2d_fused_and_plain_agree_with_golden(gpu_ctx, key):
    m, s, l, D, rich, method, d, r = key
    gs = G2D[key]
    P = qb.Parameters(m, s, int(d), int(r))
    plan = gpu_ctx.plan2d(P, method, True, D, [g.meta["a_d"] for g in gs], [g.meta["a_r"] for g in gs])
    assert plan.algorithm == 2, "the fused kernel must be the default for these shapes"
    outs = {algo: _run_plan(plan, algo) for algo in (1, 2)}
    for algo, (cells, tp, te, fl) in outs.items():
        for i, g in enumerate(gs):
            assert_slice_matches(cells[i], tp[i], te[i], fl[i], g)
    assert cell_errors(outs[2][0], outs[1][0]) <= 1e-10
    plan.close()


@pytest.mark.parametrize("key", list(GLIN), ids=lambda k: f"m{k[0]}-s{k[1]}-D{k[3]}-R{k[4]}-t{k[5]}")
def test_linear_golden(gpu_ctx, key):
    m, s, l, D, rich, target, d, r = key
    gs = GLIN[key]
    P = qb.Parameters(m, s, int(d), int(r))
    cells, tp, fl = gpu_ctx.slice1d_batch(P, target, bool(rich), D, [g.meta["a"] for g in gs])
    for i, g in enumerate(gs):
        assert_slice_matches(cells[i], tp[i], 0, fl[i], g)


@pytest.mark.parametrize("key", list(GDIAG), ids=lambda k: f"m{k[0]}-sigma{k[3]}-D{k[4]}")
def test_diagonal_golden(gpu_ctx, key):
    m, s, l, sigma, D, rich, d, r = key
    gs = GDIAG[key]
    P = qb.Diagonal_Parameters(m, sigma, s, int(d), int(r), eta_bound=25)
    cells, tp, fl = gpu_ctx.slice1d_batch(P, 2, bool(rich), D, [g.meta["a"] for g in gs],
                                          [g.meta["eta"] for g in gs])
    for i, g in enumerate(gs):
        assert_slice_matches(cells[i], tp[i], 0, fl[i], g)


def test_reference_entry_points(gpu_ctx):
    """The six entry points with the reference's calling convention (host mirror)."""
    g = next(x for x in G if x.meta["name"].startswith("2d/c1/130_129"))
    P = qb.Parameters(g.meta["m"], g.meta["s"], g.d, g.r)
    sl = qb.Distribution_Slice(g.meta["D"])
    sl.flags = 0x00000100 | 0x00040000  # other bits survive, stale method bits do not
    qb.distribution_slice_compute_richardson(sl, P, 0, 130, 129, ctx=gpu_ctx)
    assert sl.flags == (0x00000100 | g.flags) and (sl.min_log_alpha_d, sl.min_log_alpha_r) == (130, 129)
    assert cell_errors(sl.norm_matrix, g.cells) <= CELL_RTOL
    assert sl.norm_matrix.dtype == np.longdouble
    assert abs(float(sl.total_probability - np.sum(sl.norm_matrix))) < 1e-15

    g = next(x for x in G if x.meta["name"] == "2d/c2s/2040_2041")      # single pass
    sl = qb.Distribution_Slice(g.meta["D"])
    qb.distribution_slice_compute(sl, qb.Parameters(2048, 1, g.d, g.r), 0, 2040, 2041, ctx=gpu_ctx)
    assert sl.flags == g.flags == qb.SLICE_FLAGS_METHOD_SIMPSON
    assert cell_errors(sl.norm_matrix, g.cells) <= CELL_RTOL

    g = next(x for x in G if x.meta["name"] == "lin/c1/t0/-130")
    ls = qb.Linear_Distribution_Slice(g.meta["D"])
    qb.linear_distribution_slice_compute_richardson(ls, qb.Parameters(128, 2, g.d, g.r), 0, -130, ctx=gpu_ctx)
    assert cell_errors(ls.norm_vector, g.cells) <= CELL_RTOL and ls.total_error == 0 and ls.min_log_alpha == -130
    g = next(x for x in G if x.meta["name"] == "lin/single/t1/127")
    ls = qb.Linear_Distribution_Slice(g.meta["D"])
    qb.linear_distribution_slice_compute(ls, qb.Parameters(128, 2, g.d, g.r), 1, 127, ctx=gpu_ctx)
    assert cell_errors(ls.norm_vector, g.cells) <= CELL_RTOL and ls.flags == g.flags

    g = next(x for x in G if x.meta["name"] == "diag/m128/126_-1")
    ds = qb.Diagonal_Distribution_Slice(g.meta["D"])
    qb.diagonal_distribution_slice_compute_richardson(
        ds, qb.Diagonal_Parameters(128, 5, 1, g.d, g.r, eta_bound=25), 126, -1, ctx=gpu_ctx)
    assert cell_errors(ds.norm_vector, g.cells) <= CELL_RTOL and ds.eta == -1 and ds.min_log_alpha_r == 126

    with pytest.raises(qb.CriticalError, match="[Uu]nknown method"):
        qb.distribution_slice_compute_richardson(qb.Distribution_Slice(32), P, 7, 130, 129, ctx=gpu_ctx)
    with pytest.raises(qb.CriticalError, match="Unknown target"):
        qb.linear_distribution_slice_compute(qb.Linear_Distribution_Slice(32), P, 3, 128, ctx=gpu_ctx)


def test_against_reference_on_this_box(gpu_ctx):
    """oracle/_ref travels with the snapshot: fresh coordinates, not in the fixtures."""
    ref = ref_or_none()
    if ref is None:
        pytest.skip("oracle/_ref/libqref.so not present")
    rng = np.random.default_rng(7)
    for (m, s, D) in ((2048, 1, 32), (128, 2, 32), (3072, 4, 32), (512, 3, 20)):
        d, r = ref.deterministic_d_r(m)
        P, RP = qb.Parameters(m, s, d, r), ref.RefParameters(m, s, d, r)
        ad = [int(x) for x in rng.integers(m - 12, m + 11, 3) * rng.choice([-1, 1], 3)]
        ar = [int(x) for x in rng.integers(m - 12, m + 11, 3)]
        cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 0, True, D, ad, ar)
        for i in range(3):
            R = ref.distribution_slice_compute(RP, D, ad[i], ar[i])
            assert cell_errors(cells[i], R.cells) <= CELL_RTOL, (m, ad[i], ar[i])
            assert abs(float(tp[i] - R.total_probability)) <= 1e-12
            assert abs(float((te[i] - R.total_error) / R.total_error)) <= 1e-9
            assert int(fl[i]) == R.flags
        a1 = [int(x) for x in rng.integers(m - 25, m + 11, 3) * rng.choice([-1, 1], 3)]
        for target in (1,) if m > 1024 else (0, 1):
            c1, tp1, f1 = gpu_ctx.slice1d_batch(P, target, True, 64, a1)
            for i in range(3):
                R = ref.linear_distribution_slice_compute(RP, 64, a1[i], target)
                assert cell_errors(c1[i], R.cells) <= CELL_RTOL and abs(float(tp1[i] - R.total_probability)) <= 1e-12


# ---- full-size properties --------------------------------------------------------------

def _t2d_params():
    import random
    rnd = random.Random(20482048)
    m = 2048
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    return qb.Parameters(m, 1, d, r)


def test_full_size_properties(gpu_ctx):
    P = _t2d_params()
    coords = shard.enumerate_2d(2048)
    ad = [c[0] for c in coords]
    ar = [c[1] for c in coords]
    D = 128
    plan = gpu_ctx.plan2d(P, 0, True, D, ad, ar)
    assert plan.cells == 3362 * D * D and plan.algorithm == 2
    c_a, tp_a, te_a, fl_a = _run_plan(plan, 2)
    c_b, tp_b, te_b, fl_b = _run_plan(plan, 2)
    # determinism: bit-identical re-run
    assert np.array_equal(c_a, c_b) and np.array_equal(tp_a, tp_b) and np.array_equal(te_a, te_b)
    # the slice's total is the sum of its cells
    assert np.max(np.abs(c_a.sum(axis=1) - tp_a.astype(np.float64))) < 1e-15
    # captured mass: alpha_r > 0 half of the distribution
    mass = float(tp_a.sum())
    assert 0.4999 < mass < 0.5, mass
    assert np.all(fl_a == (qb.SLICE_FLAGS_METHOD_SIMPSON | qb.SLICE_FLAGS_METHOD_RICHARDSON))
    assert np.all(te_a > 0) and np.all(np.isfinite(c_a))
    # batch-composition invariance (bit-exact) and fused == plain on a sample
    sample = list(range(0, len(coords), 97))
    sub = gpu_ctx.plan2d(P, 0, True, D, [ad[i] for i in sample], [ar[i] for i in sample])
    c_s, tp_s, te_s, _ = _run_plan(sub, 2)
    assert np.array_equal(c_s, c_a[sample]) and np.array_equal(tp_s, tp_a[sample])
    c_p, tp_p, te_p, _ = _run_plan(sub, 1)
    assert cell_errors(c_s, c_p) <= 1e-10
    assert np.max(np.abs((tp_s - tp_p).astype(np.float64))) <= 1e-14
    assert np.max(np.abs(((te_s - te_p) / te_p).astype(np.float64))) <= 1e-10
    # Richardson identity from two single passes (src/distribution_slice_compute_richardson.cpp:47-64)
    k = sample[3]
    co, _, _, _ = gpu_ctx.slice2d_batch(P, 0, False, D, [ad[k]], [ar[k]])
    fi, _, _, _ = gpu_ctx.slice2d_batch(P, 0, False, 2 * D, [ad[k]], [ar[k]])
    f = fi[0].reshape(2 * D, 2 * D)
    four = f[0::2, 0::2] + f[0::2, 1::2] + f[1::2, 0::2] + f[1::2, 1::2]
    rich = 2 * four.reshape(-1) - co[0]
    assert cell_errors(c_a[k], rich) <= 1e-10
    plan.close()
    sub.close()


def test_sigma_optimal_matches_reference_on_this_box(gpu_ctx):
    """-sigma-optimal: the parallel fixed-point solution of the reference's serial walk."""
    ref = ref_or_none()
    if ref is None:
        pytest.skip("oracle/_ref/libqref.so not present")
    for (m, s, D, coords) in ((2048, 30, 16, [(2050, 2049), (-2047, 2046), (2056, 2057)]),
                              (256, 3, 12, [(256, 255), (-259, 258)])):
        d, r = ref.deterministic_d_r(m)
        P, RP = qb.Parameters(m, s, d, r), ref.RefParameters(m, s, d, r)
        ad, ar = [c[0] for c in coords], [c[1] for c in coords]
        cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 1, True, D, ad, ar)
        for i in range(len(coords)):
            R = ref.distribution_slice_compute(RP, D, ad[i], ar[i], method=1)
            assert cell_errors(cells[i], R.cells) <= CELL_RTOL, (m, coords[i])
            assert abs(float(tp[i] - R.total_probability)) <= 1e-12
            assert abs(float((te[i] - R.total_error) / R.total_error)) <= 1e-9
            assert int(fl[i]) == R.flags


def test_replayed_plans_are_cuda_graphs_with_the_same_results(gpu_ctx):
    """A plan run again into the same buffers replays its step as one CUDA graph (the class kernels on
    side streams, summaries by ticket): the same cells and summaries bit for bit as the eager first
    run, also after the buffers change, for the heuristic and the sigma-optimal (large l) methods."""
    import torch
    P = _t2d_params()
    coords = shard.enumerate_2d(2048)[::40]
    ad, ar = [c[0] for c in coords], [c[1] for c in coords]
    for method in (0, 1):
        plan = gpu_ctx.plan2d(P, method, True, 64, ad, ar)
        st = torch.cuda.Stream()
        bufs = [(torch.zeros(plan.cells, dtype=torch.float64, device="cuda"),
                 torch.zeros(plan.n * 8, dtype=torch.float64, device="cuda")) for _ in range(2)]
        outs, counts = [], []
        for k in (0, 0, 0, 1, 1, 0, 0):                   # eager, capture + replay, replay, new buffers ...
            c, sm = bufs[k]
            c.zero_()
            sm.zero_()
            torch.cuda.synchronize()
            l0 = gpu_ctx.launch_count
            plan.run(c.data_ptr(), sm.data_ptr(), st.cuda_stream)
            torch.cuda.synchronize()
            counts.append(gpu_ctx.launch_count - l0)
            outs.append((c.cpu().numpy().copy(), sm.cpu().numpy().copy()))
        for c, sm in outs[1:]:
            assert np.array_equal(c, outs[0][0]) and np.array_equal(sm, outs[0][1])
        assert len(set(counts)) == 1 and counts[0] >= 3, counts     # the replays account for their kernels
        tp, te, fl = plan.finish(outs[-1][1])
        assert np.all(tp > 0) and np.all(te > 0)
        plan.close()


def test_sigma_optimal_closed_form_walk_equals_the_iteration(gpu_ctx):
    """Large l (>= 256): the sigma-optimal walk is a prefix minimum in closed form (k_so_fast) over
    the quick method's cells. Against the general fixed-point iteration (QB200_SO_FAST=0): cells to
    1e-10 (the iteration evaluates every point in double-double), total_error to 1e-9, flags equal;
    and against the reference itself where oracle/_ref is on the box."""
    import os
    ref = ref_or_none()
    cases = [(_t2d_params(), 2048, 1, 32, [(2048, 2048), (-2049, 2048), (2058, 2057), (2018, 2058), (-2058, 2018), (2040, 2041)]),
             (None, 512, 1, 32, [(512, 512), (-510, 515), (520, 521)]),
             (None, 1024, 3, 64, [(1024, 1025), (-1030, 1020)])]
    worst = [0.0, 0.0]
    for P, m, s, D, coords in cases:
        if P is None:
            from oracle import restate as rs
            d, r = rs.deterministic_d_r(m)
            P = qb.Parameters(m, s, d, r)
        ad, ar = [c[0] for c in coords], [c[1] for c in coords]
        l0 = gpu_ctx.launch_count
        cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 1, True, D, ad, ar)
        fast_launches = gpu_ctx.launch_count - l0
        os.environ["QB200_SO_FAST"] = "0"
        try:
            l0 = gpu_ctx.launch_count
            c0, tp0, te0, fl0 = gpu_ctx.slice2d_batch(P, 1, True, D, ad, ar)
            slow_launches = gpu_ctx.launch_count - l0
        finally:
            del os.environ["QB200_SO_FAST"]
        assert fast_launches <= 8 < slow_launches, (fast_launches, slow_launches)
        assert cell_errors(cells, c0) <= 1e-10
        assert np.max(np.abs((tp - tp0).astype(np.float64))) <= 1e-14
        rel = float(np.max(np.abs(((te - te0) / te0).astype(np.float64))))
        assert rel <= 1e-9 and np.array_equal(fl, fl0), (m, s, rel)
        worst[0] = max(worst[0], rel)
        if ref is not None and m <= 1024:
            RP = ref.RefParameters(m, s, P.d, P.r)
            for i in range(min(2, len(coords))):
                R = ref.distribution_slice_compute(RP, D, ad[i], ar[i], method=1)
                assert cell_errors(cells[i], R.cells) <= CELL_RTOL
                assert abs(float(tp[i] - R.total_probability)) <= 1e-12
                rr = abs(float((te[i] - R.total_error) / R.total_error))
                assert rr <= 1e-9 and int(fl[i]) == R.flags
                worst[1] = max(worst[1], rr)
    print(f"sigma-optimal closed form: total_error vs iteration {worst[0]:.2e}, vs reference {worst[1]:.2e}")


def test_empty_and_ragged_batches(gpu_ctx):
    P = _t2d_params()
    cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 0, True, 32, [], [])
    assert cells.shape == (0, 1024) and len(tp) == 0
    # one slice, odd dimension, single pass and Richardson agree with the plain identity
    c1, tp1, _, _ = gpu_ctx.slice2d_batch(P, 2, True, 7, [2048], [2047])
    assert c1.shape == (1, 49) and abs(float(tp1[0]) - c1.sum()) < 1e-16
    c2, tp2, f2 = gpu_ctx.slice1d_batch(P, 1, True, 1, [2048])
    assert c2.shape == (1, 1) and float(tp2[0]) == c2[0, 0]


@pytest.mark.gpu
def test_reference_unit_tests_linear_and_diagonal_totals(gpu_ctx):
    """The reference's own slice-level unit tests, through the CUDA path and the reference-named
    entry points: test_linear_distribution() (src/test/test_linear_distribution.cpp: m = 128,
    s = 1, deterministic d and r, dimension 2048, 16 offsets x 2 signs x 2 targets) and
    test_diagonal_distribution() (src/test/test_diagonal_distribution.cpp: sigma = 5, 9 offsets x
    7 eta x 2 signs) against the Mathematica totals they quote, with their tolerance
    (test_cmp_ld, src/test/test_common.cpp:75-93)."""
    import json
    import os
    import qunundrum_b200 as qb
    from oracle import restate as rs
    from tests.conftest import GOLDEN
    totals = json.load(open(os.path.join(GOLDEN, "mathematica_totals.json")))

    def close(a, b, tol):
        a, b = float(a), float(b)
        return a > 0 and b > 0 and abs(a - b) / min(a, b) <= tol

    m = 128
    d, r = rs.deterministic_d_r(m)
    P = qb.Parameters(m, 1, d, r)
    for target in (qb.LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D, qb.LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_R):
        vals = totals["linear"][target]["values"]
        for i, off in enumerate(range(-5, 11)):
            for sign in (1, -1):
                s = qb.Linear_Distribution_Slice(2048)
                qb.linear_distribution_slice_compute_richardson(s, P, target, sign * (m + off), ctx=gpu_ctx)
                tol = 1e-4 if (target == 1 and off >= 10) else 1e-6
                assert close(s.total_probability, vals[i], tol), (target, off, sign)
                assert s.flags & qb.SLICE_FLAGS_METHOD_RICHARDSON and s.min_log_alpha == sign * (m + off)
    DP = qb.Diagonal_Parameters(m, 5, 1, d, r, eta_bound=25)
    pos, neg = totals["diagonal"][0]["values"], totals["diagonal"][1]["values"]
    offsets, etas = list(range(-5, 4)), [0, 1, -1, 2, -2, 25, -25]
    for i in range(63):
        off, eta = offsets[i % 9], etas[i // 9]
        for sign, want in ((1, pos[i]), (-1, neg[i])):
            s = qb.Diagonal_Distribution_Slice(2048)
            qb.diagonal_distribution_slice_compute_richardson(s, DP, sign * (m + off), eta, ctx=gpu_ctx)
            assert close(s.total_probability, want, 1e-6), (off, eta, sign)
            assert s.eta == eta
