"""Run a few passes over the T2D set (for ncu / quick timing).

    python tests/tools/prof_t2d.py [steps] [D] [share]     # share N: the slices a rank of N holds (i mod N == 0)
"""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import qunundrum_b200 as qb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
random.seed(20482048)
m = 2048
r = 2 ** (m - 1) + 1 + random.randrange(2 ** (m - 1) - 1)
d = r // 2 + random.randrange(r // 2)
P = qb.Parameters(m, 1, d, r)
share = int(sys.argv[3]) if len(sys.argv) > 3 else 1
from qunundrum_b200 import shard
coords = shard.enumerate_2d(m)[::share] if share > 1 else [(sd * a, b) for a in range(2018, 2059) for b in range(2018, 2059) for sd in (1, -1)]
ctx = qb.Context(0)
plan = ctx.plan2d(P, 0, True, D, [c[0] for c in coords], [c[1] for c in coords])
cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
summ = torch.empty(plan.n * 8, dtype=torch.float64, device="cuda")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
for _ in range(3):
    plan.run(cells.data_ptr(), summ.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    plan.run(cells.data_ptr(), summ.data_ptr(), ts.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"D={D} slices={len(coords)} ms/step {ms:.4f} cells/s {plan.cells / ms * 1e3:.4e} launches/step {plan.launches}")
