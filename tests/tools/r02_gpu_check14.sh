#!/bin/bash
# Source-level ncu captures of the text kernels and of k_so_fast on the bench workload.
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_text' -c 6 -o gpurun_out/c14_text python tests/tools/prof_text.py 2 22 > gpurun_out/c14_ncu_text.log 2>&1
tail -5 gpurun_out/c14_ncu_text.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_so_fast' -c 1 -o gpurun_out/c14_so_fast python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text --no-tau --no-saturation --no-other-scaling > gpurun_out/c14_ncu_so_fast.log 2>&1
tail -3 gpurun_out/c14_ncu_so_fast.log
python tests/tools/ncu_summary.py gpurun_out/c14_so_fast.ncu-rep gpurun_out/c14_so_fast_ncu_full > /dev/null 2>&1
python tests/tools/ncu_summary.py gpurun_out/c14_text.ncu-rep gpurun_out/c14_text_ncu_full > /dev/null 2>&1
ls -la gpurun_out/c14*
