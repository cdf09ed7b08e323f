#!/bin/bash
# Round 2: the diagonal drop-in suites with the final integration binaries.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_diagk.py tests/test_estimate_runs_end_to_end.py -x -q -m gpu -k "dropin or diagonal" > gpurun_out/c31_tests.txt 2>&1
tail -3 gpurun_out/c31_tests.txt
QB200_DROPIN_STATS=1 timeout 300 python tests/tools/tau_diagonal_timing.py > gpurun_out/c31_tau_diagonal.txt 2>&1
tail -1 gpurun_out/c31_tau_diagonal.txt | cut -c1-900
