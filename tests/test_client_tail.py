"""SURVEY.md section 8(f) #2 -- what the generator's client and server do to the slices after the
integration, on the device: distribution_slice_copy_scale (src/distribution_slice.cpp:230-264),
linear_distribution_init_collapse_d / _r (src/linear_distribution.cpp:152-324), the export of a
resident distribution, and the single-launch one-dimensional integrator.

CPU part (`-m "not gpu"`): the __host__ __device__ arithmetic of csrc/client_math.cuh (the CPU twin
of tests/hostsim) against the reference's own functions in oracle/_ref -- bit for bit.
GPU part: the kernels against the same oracles, and the BENCH configuration itself (m = 2048,
s = 1, D = 128 / 256 / 512) against slices computed by the unmodified reference
(tests/golden/bench_slices.npz, written by tests/golden/make_bench_golden.py)."""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, ref_or_none
from tests.util import CELL_RTOL, cell_errors

LD = np.longdouble


def _bits(a):
    a = np.ascontiguousarray(a, dtype=LD)
    return a.view(np.uint8).reshape(-1, 16)[:, :10]


def _same_bits(a, b):
    return np.array_equal(_bits(a), _bits(b))


def _random_cells(rng, n, lo=-12, hi=0):
    c = (rng.random(n).astype(LD) + rng.random(n).astype(LD) * LD(2) ** -50) * LD(10.0) ** rng.integers(lo, hi, n).astype(LD)
    c[::9] *= LD(-1e-2)       # Richardson leaves negative cells
    return c


# ---- CPU: the twin of the kernels' arithmetic against the reference --------------------------

def test_x87_division_by_small_integers_matches_the_fpu():
    from tests import hostsim as hs
    rng = np.random.default_rng(5)
    for q in (1, 2, 3, 4, 5, 6, 7, 12, 96, 1000003, 4294967295):
        a = _random_cells(rng, 400, -300, 300)
        for x in a:
            assert hs.x87_div_u32(x, q) == x / LD(q), (x, q)
    for v in (0.0, 1.5, -3.25e-300, 5e-324, 2.2e-308, 1e308):
        assert hs.x87_from_double(v) == LD(v)


def test_copy_scale_twin_matches_reference():
    ref = ref_or_none()
    if ref is None:
        pytest.skip("oracle/_ref/libqref.so not present")
    from tests import hostsim as hs
    rng = np.random.default_rng(6)
    for (D, store) in ((8, 4), (16, 4), (32, 16), (12, 4), (64, 64), (64, 1)):
        cells = _random_cells(rng, D * D).astype(np.float64)
        want, _ = ref.distribution_slice_copy_scale(cells.astype(LD), D, 0, store)
        assert _same_bits(hs.copy_scale(cells, D, store), want), (D, store)


def _mixed_distribution(ref, rng):
    m = 128
    d, r = ref.deterministic_d_r(m)
    dims, a, b, cells = [], [], [], []
    for ad in (126, 127, -127, 128):
        for ar in (125, 126, -128):
            D = 16 if (ad + ar) % 3 else 32
            if ad == 128 and ar == 126:
                D = 8
            dims.append(D)
            a.append(ad)
            b.append(ar)
            cells.append(_random_cells(rng, D * D))
    return m, d, r, dims, a, b, cells


def test_collapse_twin_matches_reference():
    ref = ref_or_none()
    if ref is None:
        pytest.skip("oracle/_ref/libqref.so not present")
    from tests import hostsim as hs
    m, d, r, dims, a, b, cells = _mixed_distribution(ref, np.random.default_rng(7))
    rd = ref.RefDistribution(2, ref.RefParameters(m, 2, d, r), dims, a, b, np.concatenate(cells))
    for axis in (0, 1):
        coords, vec, _ = rd.collapse(axis)
        key = a if axis == 0 else b
        assert list(coords) == list(dict.fromkeys(key))          # first-appearance order
        for k, c in enumerate(coords):
            src = [cells[i] for i in range(len(dims)) if key[i] == c]
            assert _same_bits(hs.collapse(axis, max(dims), src), vec[k]), (axis, c)


def test_bench_goldens_are_committed():
    meta = json.load(open(os.path.join(GOLDEN, "bench_slices_meta.json")))
    z = np.load(os.path.join(GOLDEN, "bench_slices.npz"))
    assert meta["m"] == 2048 and len(meta["slices"]) == 16
    assert sorted({q["tag"] for q in meta["slices"]}) == ["c0", "c1", "c2", "so", "up1024", "up512"]
    for q in meta["slices"]:
        n = (q["scale_to"] or q["D"]) ** 2
        assert z[q["name"] + "/cells_hi"].size == n


# ---- GPU ------------------------------------------------------------------------------------------

def _bench_goldens():
    meta = json.load(open(os.path.join(GOLDEN, "bench_slices_meta.json")))
    z = np.load(os.path.join(GOLDEN, "bench_slices.npz"))
    import random
    rnd = random.Random(meta["seed"])
    m = meta["m"]
    r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    out = []
    for q in meta["slices"]:
        cells = z[q["name"] + "/cells_hi"].astype(LD) + z[q["name"] + "/cells_lo"].astype(LD)
        q = dict(q, cells=cells, tp=LD(q["tp_hi"]) + LD(q["tp_lo"]), te=np.ldexp(LD(q["te_mant"]), q["te_exp"]))
        out.append(q)
    return m, meta["s"], d, r, out


def _check(q, cells, tp, te, fl, extra_flags=0):
    e = cell_errors(cells, q["cells"])
    dtp = abs(float(LD(tp) - q["tp"]))
    dte = abs(float((LD(te) - q["te"]) / q["te"]))
    assert e <= CELL_RTOL and dtp <= 1e-12 and dte <= 1e-9, (q["name"], e, dtp, dte)
    assert int(fl) | extra_flags == q["flags"], (q["name"], hex(int(fl)), hex(q["flags"]))
    return e, dtp, dte


@pytest.mark.gpu
def test_bench_configuration_against_the_reference(gpu_ctx):
    """The bench workload at the dimensions the generator uses: 13 slices at D = 128 across the fused
    kernel's three classes / both signs of alpha_d in ONE batch, an upgraded slice at D = 256, one
    at D = 512 scaled to 256 on the device, and the sigma-optimal method at D = 128."""
    import qunundrum_b200 as qb
    m, s, d, r, G = _bench_goldens()
    P = qb.Parameters(m, s, d, r)
    worst = [0.0, 0.0, 0.0]

    def upd(t):
        for i in range(3):
            worst[i] = max(worst[i], t[i])

    g128 = [q for q in G if q["D"] == 128 and q["method"] == 0]
    plan = gpu_ctx.plan2d(P, 0, True, 128, [q["a_d"] for q in g128], [q["a_r"] for q in g128])
    assert plan.algorithm == 2, "the bench configuration must run on the fused kernel"
    plan.close()
    cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 0, True, 128, [q["a_d"] for q in g128], [q["a_r"] for q in g128])
    for i, q in enumerate(g128):
        upd(_check(q, cells[i], tp[i], te[i], fl[i]))
    q = next(q for q in G if q["tag"] == "up512")
    cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 0, True, 256, [q["a_d"]], [q["a_r"]])
    upd(_check(q, cells[0], tp[0], te[0], fl[0]))
    q = next(q for q in G if q["tag"] == "up1024")
    sc, tp, te, fl = gpu_ctx.slice2d_batch_scaled(P, 0, True, 512, 256, [q["a_d"]], [q["a_r"]])
    upd(_check(q, sc[0], tp[0], te[0], fl[0], extra_flags=0x00100000))     # SLICE_FLAGS_SCALED
    # ... and the scaling itself is the reference's, bit for bit, on the unscaled doubles
    full, _, _, _ = gpu_ctx.slice2d_batch(P, 0, True, 512, [q["a_d"]], [q["a_r"]])
    from tests import hostsim as hs
    assert _same_bits(sc[0], hs.copy_scale(full[0], 512, 256))
    ref = ref_or_none()
    if ref is not None:
        want, _ = ref.distribution_slice_copy_scale(full[0].astype(LD), 512, 0, 256)
        assert _same_bits(sc[0], want)
    q = next(q for q in G if q["tag"] == "so")
    cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 1, True, 128, [q["a_d"]], [q["a_r"]])
    upd(_check(q, cells[0], tp[0], te[0], fl[0]))
    print(f"bench configuration vs reference: worst cell {worst[0]:.2e}, mass {worst[1]:.2e}, "
          f"total_error {worst[2]:.2e}")


@pytest.mark.gpu
def test_copy_scale_on_device_is_bit_exact(gpu_ctx):
    import qunundrum_b200 as qb
    from tests import hostsim as hs
    ref = ref_or_none()
    d, r = hs_params(256)
    P = qb.Parameters(256, 2, d, r)
    ad, ar = [256, -257, 258, 250], [255, 256, 258, 251]
    for D, store in ((64, 16), (32, 32), (48, 16), (64, 1)):
        full, tpf, tef, flf = gpu_ctx.slice2d_batch(P, 0, True, D, ad, ar)
        sc, tp, te, fl = gpu_ctx.slice2d_batch_scaled(P, 0, True, D, store, ad, ar)
        assert np.array_equal(tp, tpf) and np.array_equal(te, tef) and np.array_equal(fl, flf)
        for i in range(len(ad)):
            assert _same_bits(sc[i], hs.copy_scale(full[i], D, store)), (D, store, i)
            if ref is not None:
                want, _ = ref.distribution_slice_copy_scale(full[i].astype(LD), D, 0, store)
                assert _same_bits(sc[i], want)
    with pytest.raises(qb.CriticalError):
        gpu_ctx.slice2d_batch_scaled(P, 0, True, 48, 32, ad, ar)


def hs_params(m):
    from oracle import restate as rs
    return rs.deterministic_d_r(m)


@pytest.mark.gpu
def test_collapse_and_export_of_a_resident_distribution(gpu_ctx):
    """linear_distribution_init_collapse_d / _r on the device: bit-identical vectors, same slice
    order; the export of the resident slices: the bytes of the per-slice exporter."""
    import qunundrum_b200 as qb
    from tests import hostsim as hs
    ref = ref_or_none()
    rng = np.random.default_rng(11)
    m = 128
    d, r = hs_params(m)
    P = qb.Parameters(m, 2, d, r)
    # real slices at mixed dimensions (what the dimension heuristic leaves behind), mirrored
    dist = qb.Distribution(m)
    for D, coords in ((16, [(126, 125), (127, 125), (-127, 126), (128, 126)]), (32, [(128, 128), (129, 128), (-128, 127)]),
                      (8, [(120, 121)])):
        cells, tp, te, fl = gpu_ctx.slice2d_batch(P, 0, True, D, [c[0] for c in coords], [c[1] for c in coords])
        for i, (a, b) in enumerate(coords):
            for sa, sb in ((a, b), (-a, -b)):
                sl = qb.Distribution_Slice(D, sa, sb, norm_matrix=cells[i].astype(LD))
                sl.total_probability, sl.total_error, sl.flags = tp[i], te[i], int(fl[i])
                dist.insert_slice(sl)
    # ... and a few synthetic ones with negative / tiny cells
    for k in range(3):
        sl = qb.Distribution_Slice(16, 126 + k, -125, norm_matrix=_random_cells(rng, 256, -40, -2))
        sl.total_probability = sl.norm_matrix.sum()
        dist.insert_slice(sl)
    dist.sort_slices()
    res = qb.Resident(dist.slices, gpu_ctx)
    try:
        for axis in (0, 1):
            coords, vec, tot = res.collapse(axis)
            key = [int(s.min_log_alpha_d if axis == 0 else s.min_log_alpha_r) for s in dist.slices]
            assert list(coords) == list(dict.fromkeys(key))
            for k, c in enumerate(coords):
                src = [s.norm_matrix for s, kk in zip(dist.slices, key) if kk == c]
                assert _same_bits(vec[k], hs.collapse(axis, 32, src)), (axis, c)
            if ref is not None:
                rd = ref.RefDistribution(2, ref.RefParameters(m, 2, d, r), [s.dimension for s in dist.slices],
                                         [s.min_log_alpha_d for s in dist.slices],
                                         [s.min_log_alpha_r for s in dist.slices],
                                         np.concatenate([s.norm_matrix for s in dist.slices]),
                                         [s.total_probability for s in dist.slices])
                rc, rv, rt = rd.collapse(axis)
                assert list(rc) == list(coords) and _same_bits(rv, vec) and _same_bits(rt, tot)
        # the reference-named entry points
        lin = qb.linear_distribution_init_collapse_d(dist, ctx=gpu_ctx)
        assert len(lin.slices) == len(set(s.min_log_alpha_d for s in dist.slices))
        assert abs(float(sum(s.norm_vector.sum() for s in lin.slices) - sum(s.norm_matrix.sum() for s in dist.slices))) < 1e-15
        # export from the device copy == the per-slice exporter (itself byte-identical to libc)
        n = len(dist.slices)
        texts = res.format(0, n)
        for i, s in enumerate(dist.slices):
            assert texts[i] == gpu_ctx.text_format(s.norm_matrix, s.total_error), i
        assert res.format(3, 2) == texts[3:5] and res.format(n - 1, 1) == texts[-1:]
        assert res.format(0, 0) == []
        # look-ahead: the next batch is formatted while the caller still holds this one
        first = res.format(0, 4, prefetch_next=3)
        assert first == texts[0:4]
        assert res.format(4, 3) == texts[4:7] and first == texts[0:4]
        assert res.format(2, 5, prefetch_next=2) == texts[2:7] and res.format(0, 1) == texts[0:1]   # a prefetch left unused
    finally:
        res.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["linear_d", "linear_r", "diagonal"])
def test_single_launch_1d_kernel_equals_the_plain_path(gpu_ctx, kind):
    """k_fused1d (one launch per distribution, 5 evaluations per cell) against k_vals1d / k_cells1d /
    k_final1d (6 D + 2 point values through memory): the same cells and summaries, bit for bit,
    Richardson and single pass, power-of-two and ragged dimensions, one slice and many."""
    import torch
    import qunundrum_b200 as qb
    d, r = hs_params(128)
    if kind == 2:
        P = qb.Diagonal_Parameters(128, 5, 1, d, r, eta_bound=2)
        a = [120 + i for i in range(12)] * 2
        eta = [0] * 12 + [-2] * 12
    else:
        P = qb.Parameters(128, 2, d, r)
        a = [98 + i for i in range(41)] + [-128, -130]
        eta = None
    for D, rich in ((2048, True), (100, True), (129, False), (1, True)):
        plan = gpu_ctx.plan1d(P, kind, rich, D, a, eta)
        assert plan.algorithm == 2 and plan.launches == 1
        out = {}
        for algo in (2, 1, 2):
            plan.set_algorithm(algo)
            cells = torch.zeros(plan.cells, dtype=torch.float64, device="cuda")
            summ = torch.zeros(len(a) * 8, dtype=torch.float64, device="cuda")
            st = torch.cuda.Stream()
            plan.run(cells.data_ptr(), summ.data_ptr(), st.cuda_stream)
            torch.cuda.synchronize()
            got = (cells.cpu().numpy(), summ.cpu().numpy())
            if algo in out:      # the fused kernel again: tickets were left at zero, same bits
                assert np.array_equal(out[algo][0], got[0]) and np.array_equal(out[algo][1], got[1])
            out[algo] = got
        assert np.array_equal(out[1][0], out[2][0]), (kind, D, rich)
        # the slice totals: double-double tree sums over 256- vs 128-cell blocks, equal to ~1e-31
        t1 = out[1][1].reshape(-1, 8)[:, 0].astype(LD) + out[1][1].reshape(-1, 8)[:, 1].astype(LD)
        t2 = out[2][1].reshape(-1, 8)[:, 0].astype(LD) + out[2][1].reshape(-1, 8)[:, 1].astype(LD)
        assert np.all(np.abs(t1 - t2) <= LD(2) ** -62 * np.abs(t1))
        plan.close()
