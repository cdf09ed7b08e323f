#!/bin/bash
# k_diagk with the quotient formulation (s rho, one Barrett division, low bits of s D') and the shortened r j.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_diagk.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c19_tests.txt 2>&1
tail -3 gpurun_out/c19_tests.txt
timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c19_prof_diagk.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c19_prof_diagk.txt | head -1
QB200_DIAGK_FULL_PRODUCT=1 timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c19_prof_diagk_full.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c19_prof_diagk_full.txt | head -1
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_diagk.py > gpurun_out/c19_sanitizer_diagk.txt 2>&1
tail -3 gpurun_out/c19_sanitizer_diagk.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/c19_diagk python tests/tools/prof_diagk.py > gpurun_out/c19_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c19_diagk.ncu-rep gpurun_out/c19_diagk_ncu_full > /dev/null 2>&1
grep -E "duration|issue_active|warp_instructions|stall_long|dram" gpurun_out/c19_diagk_ncu_full.txt
timeout 300 python tests/tools/tau_diagonal_timing.py > gpurun_out/c19_tau_diagonal.txt 2>&1
tail -2 gpurun_out/c19_tau_diagonal.txt | cut -c1-700
