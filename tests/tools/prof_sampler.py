"""One tau-estimation pass on the bench distribution (for ncu / timing):

    python tests/tools/prof_sampler.py            # JSON of bench.py's `tau` section (no CPU baseline)
    ncu --set full --clock-control none --import-source on -k regex:k_sample -c 2 \
        -o gpurun_out/sampler python tests/tools/prof_sampler.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import qunundrum_b200 as qb  # noqa: E402
from qunundrum_b200 import shard  # noqa: E402

ctx = qb.Context(0)
d, r = bench.synthetic_d_r(20482048)
P = qb.Parameters(bench.M, bench.S, d, r, bench.T_PARAM)
coords = shard.enumerate_2d(bench.M)
a_d = np.array([c[0] for c in coords], dtype=np.int32)
a_r = np.array([c[1] for c in coords], dtype=np.int32)
cells, tp, te, fl = ctx.slice2d_batch(P, 0, True, bench.DIM, a_d, a_r)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
out = bench.tau_section(ctx, qb, torch, stream, cells, tp, coords, hbm,
                        cpu_baseline="--cpu" in sys.argv)
print(json.dumps(out))
