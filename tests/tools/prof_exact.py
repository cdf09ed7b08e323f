"""Throughput of the exact samplers (bench.py's `exact` section on its own, for timing and ncu):

    python tests/tools/prof_exact.py [m l dimension n] [--ref]       # JSON line
    ncu --set full --clock-control none --import-source on -k regex:k_exact_jk$ -c 1 \
        -o gpurun_out/exact python tests/tools/prof_exact.py
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
m, l, dim, n = (int(a) for a in args) if len(args) == 4 else (2048, 2048, 256, 1024 * 148)
ctx = qb.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
print(json.dumps(bench.exact_section(ctx, qb, torch, stream, cpu_baseline="--ref" in sys.argv,
                                     m=m, l=l, dim=dim, n=n)))
