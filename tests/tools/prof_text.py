"""Time the text exporter kernel on device-resident values (for ncu / quick timing).

    python tests/tools/prof_text.py [steps] [log2_n] [kind: cells|random]

cells: the cells of the bench workload's slices (real exported data); random: random x87
values in [1e-120, 1).
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import qunundrum_b200 as qb
from oracle import text as ot

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << log2n
rng = np.random.default_rng(3)
mant = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64) | np.uint64(1 << 63)
se = rng.integers(16383 - 400, 16384, size=n).astype(np.uint16)
vals = ot.ld_from_fields(mant, se)
ctx = qb.Context(0)
d_in = torch.from_numpy(vals.view(np.uint8)).cuda()
cap = 33 * n
d_text = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
for _ in range(3):
    ctx.text_format_device(qb.host.TEXT_X87, d_in.data_ptr(), n, d_text.data_ptr(), cap,
                           d_len.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ctx.text_format_device(qb.host.TEXT_X87, d_in.data_ptr(), n, d_text.data_ptr(), cap,
                           d_len.data_ptr(), ts.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
L = int(d_len.item())
print(f"n={n} text {L} B ({L / n:.2f} B/value) ms/step {ms:.4f} values/s {n / ms * 1e3:.4e} "
      f"algorithmic GB/s {(16 * n + L) / ms * 1e-6:.1f}")
# spot check against libc on the first 100k values
k = 100000
t = bytes(d_text[:L].cpu().numpy().tobytes())
want = ot.format_ld24(vals[:k])
assert t[:len(want)] == want, "device text differs from libc"
t0 = time.time(); ot.format_ld24(vals[:2_000_000]); t1 = time.time()
print(f"libc fprintf on this host: {2_000_000 / (t1 - t0):.3e} values/s (1 core)")
# host API, one slice's worth (65537 values)
v1 = vals[:65537].copy()
ctx.text_format(v1[:-1], v1[-1])
t0 = time.time()
for _ in range(50):
    ctx.text_format(v1[:-1], v1[-1])
t1 = time.time()
print(f"host API, 65536 cells + total_error per call: {(t1 - t0) / 50 * 1e3:.3f} ms/call")
# ---- importer: the text just written, back to values -------------------------------
d_vals = torch.empty(2 * n, dtype=torch.int64, device="cuda")
d_info = torch.zeros(8, dtype=torch.int64, device="cuda")
for _ in range(2):
    ctx.text_parse_device(d_text.data_ptr(), L, n, d_vals.data_ptr(), d_info.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    ctx.text_parse_device(d_text.data_ptr(), L, n, d_vals.data_ptr(), d_info.data_ptr(), ts.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
info = d_info.cpu().numpy()
assert info[0] == n and info[1] == 0, info
assert torch.equal(d_vals.view(torch.uint8).view(-1, 16)[:, :10], d_in.view(-1, 16)[:, :10]), "round trip differs"
print(f"parse: ms/step {ms:.4f} values/s {n / ms * 1e3:.4e} algorithmic GB/s {(16 * n + L) / ms * 1e-6:.1f} "
      f"(round trip bit-exact)")
t0 = time.time(); ot.parse_ld(t[:len(want)], k); t1 = time.time()
print(f"libc fscanf on this host: {k / (t1 - t0):.3e} values/s (1 core)")
text1 = ctx.text_format(v1[:-1], v1[-1])
t0 = time.time()
for _ in range(50):
    ctx.text_parse(text1, 65537)
t1 = time.time()
print(f"host API parse, 65537 numbers per call: {(t1 - t0) / 50 * 1e3:.3f} ms/call")
