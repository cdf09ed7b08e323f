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
print("sanitize_slices: all kernels ran; launches:", ctx.launch_count)
ctx.close()
