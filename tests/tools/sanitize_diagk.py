"""Small invocations of the diagonal k sampler's kernels for compute-sanitizer:

    compute-sanitizer --tool memcheck python tests/tools/sanitize_diagk.py

Every golden case (m = 128 ... 4096, l = 13 ... 2048, m not a multiple of 32, r shorter than m bits;
partial tiles: 72 samples in a tile of 128), both walks, the tau reduction and h.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import qunundrum_b200 as qb  # noqa: E402
from tests.test_diagk import GOLD, check_gold  # noqa: E402

ctx = qb.Context(0)
only = sys.argv[1:]
for g in GOLD:
    if only and g.name not in only:
        continue
    S = qb.DiagonalKSampler(qb.Diagonal_Parameters(g.m, g.sigma, 0, g.d, g.r, eta_bound=25, l=g.l), ctx)
    check_gold(g, S)
    n, rows = 6, 48   # the random rows (the edge rows behind them include j < |eta| 2^(m+sigma) / r)
    S.tau_estimate(n, rows // n, g.J[:rows], g.eta[:rows], np.minimum(g.pivot[:rows], np.longdouble(0.9)), 50, 25)
    S.approx_h(np.array([[0.25, 0.0], [3.5, 1e-17], [-7.25, 0.0]]))
    S.close()
print("sanitize_diagk: all kernels ran; launches:", ctx.launch_count)
ctx.close()
