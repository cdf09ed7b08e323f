"""Development check run on the GPU box: plain vs fused vs reference, and a first timing."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import qunundrum_b200 as qb
from oracle import ref

out = {}
ctx = qb.Context(0)
out["fp64_peak_tflops"] = ctx.measure_fp64_peak() / 1e12
print("fp64 peak TF/s", out["fp64_peak_tflops"], flush=True)


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b) / np.abs(b)))


def run_plan(plan, algo):
    plan.set_algorithm(algo)
    cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
    summ = torch.empty(plan.n * 8, dtype=torch.float64, device="cuda")
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        plan.run(cells.data_ptr(), summ.data_ptr(), s_.cuda_stream)
    torch.cuda.synchronize()
    tp, te, fl = plan.finish(summ.cpu().numpy())
    return cells.cpu().numpy().reshape(plan.n, -1), tp, te, fl


have_ref = ref.available()
print("ref available", have_ref, flush=True)
res = []
for (m, s, D, coords) in [(2048, 1, 32, [(2048, 2048), (-2049, 2048), (2058, 2058), (2030, 2040), (2018, 2018)]),
                          (128, 2, 32, [(130, 129), (128, 128), (-127, 126), (138, 138)]),
                          (3072, 4, 32, [(3075, 3074), (-3060, 3082)])]:
    d, r = ref.deterministic_d_r(m) if have_ref else (None, None)
    P = qb.Parameters(m, s, d, r)
    RP = ref.RefParameters(m, s, d, r)
    ad = [c[0] for c in coords]
    ar = [c[1] for c in coords]
    for method in (0, 2):
        plan = ctx.plan2d(P, method, True, D, ad, ar)
        c1, tp1, te1, fl1 = run_plan(plan, 1)
        c2, tp2, te2, fl2 = run_plan(plan, 2)
        for i, (a, b) in enumerate(coords):
            R = ref.distribution_slice_compute(RP, D, a, b, method=method)
            rc = R.cells.astype(np.float64)
            row = dict(m=m, s=s, method=method, coord=(a, b),
                       plain_vs_ref=rel(c1[i], rc), fused_vs_ref=rel(c2[i], rc),
                       fused_vs_plain=rel(c2[i], c1[i]),
                       tp_ref=float(R.total_probability),
                       dtp_plain=float(tp1[i] - R.total_probability),
                       dtp_fused=float(tp2[i] - R.total_probability),
                       te_rel_plain=float(abs(te1[i] - R.total_error) / R.total_error) if R.total_error != 0 else float(te1[i]),
                       te_rel_fused=float(abs(te2[i] - R.total_error) / R.total_error) if R.total_error != 0 else float(te2[i]),
                       flags=(int(fl1[i]), int(fl2[i]), int(R.flags)))
            print(row, flush=True)
            res.append(row)
        plan.close()
out["parity2d"] = res

res = []
D = 64
for (m, s, coords) in [(2048, 1, [2040, 2050, -2048]), (128, 2, [-130, 100, 128, 138])]:
    d, r = ref.deterministic_d_r(m)
    P = qb.Parameters(m, s, d, r)
    RP = ref.RefParameters(m, s, d, r)
    for kind in (0, 1):
        cells, tp, fl = ctx.slice1d_batch(P, kind, True, D, coords)
        for i, a in enumerate(coords):
            if kind == 0 and m == 2048 and i > 0:
                continue
            R = ref.linear_distribution_slice_compute(RP, D, a, kind)
            row = dict(kind=kind, m=m, a=a, rel=rel(cells[i], R.cells), dtp=float(tp[i] - R.total_probability),
                       flags=(int(fl[i]), int(R.flags)))
            print(row, flush=True)
            res.append(row)
    for sigma, eta in ((5, 0), (5, -3), (12, 2)):
        DP = qb.Diagonal_Parameters(m, sigma, s, d, r, eta_bound=25)
        RDP = ref.RefDiagonalParameters(m, sigma, s, d, r, eta_bound=25)
        cells, tp, fl = ctx.slice1d_batch(DP, 2, True, D, coords, [eta] * len(coords))
        for i, a in enumerate(coords):
            R = ref.diagonal_distribution_slice_compute(RDP, D, a, eta)
            row = dict(kind=2, m=m, sigma=sigma, eta=eta, a=a, rel=rel(cells[i], R.cells),
                       dtp=float(tp[i] - R.total_probability))
            print(row, flush=True)
            res.append(row)
out["parity1d"] = res

# ---- timing: the T2D set ------------------------------------------------------
import random
random.seed(20482048)
m = 2048
r = 2 ** (m - 1) + 1 + random.randrange(2 ** (m - 1) - 1)
d = r // 2 + random.randrange(r // 2)
P = qb.Parameters(m, 1, d, r)
coords = [(sd * a, b) for a in range(2018, 2059) for b in range(2018, 2059) for sd in (1, -1)]
ad = [c[0] for c in coords]
ar = [c[1] for c in coords]
D = 128
t0 = time.time()
plan = ctx.plan2d(P, 0, True, D, ad, ar)
out["plan_create_s"] = time.time() - t0
print("plan create", out["plan_create_s"], "algo", plan.algorithm, "cells", plan.cells, flush=True)
cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
summ = torch.empty(plan.n * 8, dtype=torch.float64, device="cuda")
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
st = tstream.cuda_stream
for algo in (2,):
    plan.set_algorithm(algo)
    for _ in range(3):
        plan.run(cells.data_ptr(), summ.data_ptr(), st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        plan.run(cells.data_ptr(), summ.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[f"t2d_ms_algo{algo}"] = ms
    out[f"t2d_cells_per_s_algo{algo}"] = plan.cells / (ms * 1e-3)
    print("algo", algo, "ms/step", ms, "cells/s", plan.cells / (ms * 1e-3), flush=True)
tp, te, fl = plan.finish(summ.cpu().numpy())
out["t2d_total_mass"] = float(tp.sum())
print("total mass over T2D", float(tp.sum()), "flags", set(int(x) for x in fl), flush=True)
# fused vs plain on a sample of the big batch
sample = list(range(0, len(coords), 211))
plan_s = ctx.plan2d(P, 0, True, D, [ad[i] for i in sample], [ar[i] for i in sample])
c1, tp1, te1, fl1 = run_plan(plan_s, 1)
c2, tp2, te2, fl2 = run_plan(plan_s, 2)
full = cells.cpu().numpy().reshape(plan.n, -1)
den = np.maximum(np.abs(c1), 1e-300)
out["t2d_fused_vs_plain_max_rel"] = float(np.max(np.abs(c2 - c1) / den))
out["t2d_batch_vs_sample_max_abs"] = float(np.max(np.abs(full[sample] - c2)))
print("fused vs plain (sample)", out["t2d_fused_vs_plain_max_rel"], "batch consistency", out["t2d_batch_vs_sample_max_abs"], flush=True)
print("te rel fused vs plain", float(np.max(np.abs((te2 - te1) / te1))), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1, default=str)
print("DONE", flush=True)
