"""Where does the end-to-end time go? (development tool, GPU box)"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import qunundrum_b200 as qb
from qunundrum_b200 import shard
import bench

ctx = qb.Context(0)
d, r = bench.synthetic_d_r(20482048)
P = qb.Parameters(2048, 1, d, r)
coords = shard.enumerate_2d(2048)
a_d = np.array([c[0] for c in coords], dtype=np.int32); a_r = np.array([c[1] for c in coords], dtype=np.int32)
n = len(coords); D = 128
L = qb.lib()
nbytes = n * D * D * 8
hptr = L.qb200_host_alloc(nbytes)
h = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(n, D * D))
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("full sync call ms", t(lambda: ctx.slice2d_batch(P, 0, True, D, a_d, a_r, out=h)))
def mk():
    p = ctx.plan2d(P, 0, True, D, a_d, a_r); p.close()
print("plan create+destroy ms", t(mk))
dev = torch.empty(n * D * D, dtype=torch.float64, device="cuda")
ht = torch.from_numpy(h.reshape(-1))
print("is_pinned", ht.is_pinned())
print("torch D2H 440MB into qb200_host_alloc buffer ms", t(lambda: ht.copy_(dev, non_blocking=True)))
hp = torch.empty(n * D * D, dtype=torch.float64, pin_memory=True)
ms = t(lambda: hp.copy_(dev, non_blocking=True))
print("torch D2H 440MB into torch pinned ms", ms, "GB/s", nbytes / ms / 1e6)
hpg = torch.empty(n * D * D, dtype=torch.float64)
print("torch D2H pageable ms", t(lambda: hpg.copy_(dev)))
print("python arg prep ms", t(lambda: (P._c(), np.ascontiguousarray(a_d, dtype=np.int32)), 20))
