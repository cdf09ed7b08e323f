import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import qunundrum_b200 as qb
from oracle import text as ot
n = 1 << 25
rng = np.random.default_rng(3)
mant = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64) | np.uint64(1 << 63)
se = rng.integers(16383 - 400, 16384 - 14, size=n).astype(np.uint16)   # all e-style
vals = ot.ld_from_fields(mant, se)
ctx = qb.Context(0)
d_in = torch.from_numpy(vals.view(np.uint8)).cuda()
cap = 33 * n
d_text = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
def run(tag):
    for _ in range(3):
        ctx.text_format_device(0, d_in.data_ptr(), n, d_text.data_ptr(), cap, d_len.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.text_format_device(0, d_in.data_ptr(), n, d_text.data_ptr(), cap, d_len.data_ptr(), ts.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag}: {ms:.3f} ms  {n/ms*1e3:.3e} values/s")
run("chained (all e-style data)")
ctx.text_set_force_exact(2)
run("no chain")
