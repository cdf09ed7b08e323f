"""Latency of one-slice calls through the synchronous C ABI (what the drop-in does per slice)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import qunundrum_b200 as qb
random.seed(20482048); m = 2048
r = 2 ** (m - 1) + 1 + random.randrange(2 ** (m - 1) - 1); d = r // 2 + random.randrange(r // 2)
P = qb.Parameters(m, 1, d, r)
ctx = qb.Context(0)
for D in (128, 256, 512):
    out = np.empty((1, D * D))
    coords = [(2040 + i % 10, 2038 + i % 7) for i in range(200)]
    for a, b in coords[:20]:
        ctx.slice2d_batch(P, 0, True, D, [a], [b], out=out)
    t0 = time.perf_counter()
    for a, b in coords:
        ctx.slice2d_batch(P, 0, True, D, [a], [b], out=out)
    dt = (time.perf_counter() - t0) / len(coords)
    print(f"D={D}: {dt * 1e6:.1f} us per single-slice call ({D * D / dt:.3e} cells/s)")
if os.environ.get("QB200_TIMING"):
    pass
