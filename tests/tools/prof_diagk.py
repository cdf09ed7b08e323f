"""Throughput of the diagonal k sampler (bench.py's `diagk` section on its own, for timing and ncu):

    python tests/tools/prof_diagk.py [m sigma l n] [--ref]       # JSON line
    ncu --set full --clock-control none --import-source on -k regex:k_diagk$ -c 1 \
        -o gpurun_out/diagk python tests/tools/prof_diagk.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import qunundrum_b200 as qb  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
m, sigma, l, n = (int(a) for a in args) if len(args) == 4 else (2048, 5, 2048, 2048 * 148)
ctx = qb.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
print(json.dumps(bench.diagk_section(ctx, qb, torch, stream, cpu_baseline="--ref" in sys.argv,
                                     m=m, sigma=sigma, l=l, n=n)))
