"""Cold-start cost of the exact sampler path: first and later calls, m = 2048 (diagonal parameters)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import qunundrum_b200 as qb  # noqa: E402
from oracle import restate as rs  # noqa: E402

t0 = time.perf_counter()
ctx = qb.Context(0)
print(f"context {time.perf_counter() - t0:.3f} s")
m, sigma, D = 2048, 5, 2048
d, r = rs.deterministic_d_r(m)
r |= 1
P = qb.Diagonal_Parameters(m=m, sigma=sigma, s=1, d=d, r=r)
t0 = time.perf_counter()
ex = qb.ExactSampler(P, D, 0, ctx)
print(f"qb200_exact_create {time.perf_counter() - t0:.3f} s")
t0 = time.perf_counter()
dk = qb.DiagonalKSampler(P, ctx)
print(f"qb200_diagk_create {time.perf_counter() - t0:.3f} s")
g = np.random.default_rng(1)
n = 8000
regs, off = [], 0
t0 = time.perf_counter()
for _ in range(n):
    e = int(g.integers(m - 30, m + 3))
    reg = int(g.integers(0, D))
    nb, st = ex.region_bytes(e, reg, D)
    regs.append((e, reg, D, off, nb))
    off += nb
print(f"{n} x region_bytes {time.perf_counter() - t0:.3f} s")
stream = g.bytes(off)
G = qb.pack_regions(regs)
eta = np.zeros(n, dtype=np.int32)
piv = g.random(n).astype(np.longdouble)
for i in range(4):
    t0 = time.perf_counter()
    qb.diagonal_sample_drawn(dk, ex, G, None, stream, eta, piv, 1000, want_k=False)
    print(f"qb200_diagk_sample_drawn call {i}: {time.perf_counter() - t0:.4f} s")
