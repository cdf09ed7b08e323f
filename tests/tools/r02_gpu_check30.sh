#!/bin/bash
# Round 2: golden vectors of the exact samplers on the GPU, smoke.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c30_tests.txt 2>&1
tail -3 gpurun_out/c30_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c30_smoke.txt 2>&1
tail -2 gpurun_out/c30_smoke.txt
