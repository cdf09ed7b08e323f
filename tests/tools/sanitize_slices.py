"""Small invocations of every integrator / sampler kernel for compute-sanitizer:

    compute-sanitizer --tool memcheck python tests/tools/sanitize_slices.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import qunundrum_b200 as qb  # noqa: E402
from tests.conftest import golden_slices  # noqa: E402
from tests.test_sampler import GOLD, Gold  # noqa: E402

ctx = qb.Context(0)
G = golden_slices()
gs = [g for g in G if g.meta["name"].startswith("2d/c2/")][:3]
k = gs[0].meta
P = qb.Parameters(k["m"], k["s"], gs[0].d, gs[0].r)
a = [g.meta["a_d"] for g in gs]
b = [g.meta["a_r"] for g in gs]
for method in (0, 2):
    for rich in (True, False):
        ctx.slice2d_batch(P, method, rich, k["D"], a, b)          # fused (Richardson) / plain
ctx.slice2d_batch(P, 0, True, 24, a, b)                           # plain: dimension not a multiple of 32
d, r = gs[0].d, gs[0].r
ctx.slice2d_batch(P, 1, True, 16, a[:1], b[:1])                   # sigma-optimal kernels
for kind in (0, 1):
    ctx.slice1d_batch(P, kind, True, 256, [k["m"], -k["m"] - 1])
PD = qb.Diagonal_Parameters(k["m"], 5, 1, d, r, eta_bound=2)
ctx.slice1d_batch(PD, 2, True, 256, [k["m"]], [1])
for name in ("2d", "lin"):
    g = Gold(np.load(GOLD), name)
    s = qb.Sampler(g.distribution(qb), ctx)
    s.tau_estimate(3, 40, g.words)
    s.set_force_exact(True)
    s.tau_estimate(2, 10, g.words)
    s.close()
# round 2: single-launch one-dimensional kernel (ragged dimension, plain path beside it), copy_scale,
# the resident distribution (collapse to both marginals, export with look-ahead)
for kind, PP, eta in ((0, P, None), (1, P, None), (2, PD, [0, 1, -1])):
    aa = [k["m"] - 1, k["m"], -k["m"] - 2]
    plan = ctx.plan1d(PP, kind, True, 100, aa, eta)
    import torch  # noqa: E402
    cells = torch.zeros(plan.cells, dtype=torch.float64, device="cuda")
    summ = torch.zeros(len(aa) * 8, dtype=torch.float64, device="cuda")
    for algo in (2, 1):
        plan.set_algorithm(algo)
        plan.run(cells.data_ptr(), summ.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    plan.close()
sc, tp, te, fl = ctx.slice2d_batch_scaled(P, 0, True, 64, 16, a, b)
dist = qb.Distribution(k["m"])
for D in (16, 32):
    c2, tp2, te2, fl2 = ctx.slice2d_batch(P, 0, True, D, a, b)
    for i in range(len(a)):
        for sa, sb in ((a[i], b[i]), (-a[i], -b[i])):
            sl = qb.Distribution_Slice(D, sa, sb, norm_matrix=c2[i].astype(np.longdouble))
            sl.total_probability, sl.total_error = tp2[i], te2[i]
            dist.insert_slice(sl)
res = qb.Resident(dist.slices, ctx)
for axis in (0, 1):
    res.collapse(axis)
res.format(0, 4, prefetch_next=3)
res.format(4, 3)
res.close()
print("sanitize_slices: all kernels ran; launches:", ctx.launch_count)
ctx.close()
