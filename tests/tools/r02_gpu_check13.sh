#!/bin/bash
# Round 2: k_diagk with the linear passes eight limbs at a time; k_sample per thread again.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sampler.py tests/test_diagk.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c13_tests.txt 2>&1
tail -4 gpurun_out/c13_tests.txt
timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c13_prof_sampler.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c13_prof_sampler.txt | head -1
timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c13_prof_diagk.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c13_prof_diagk.txt | head -1
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_diagk.py > gpurun_out/c13_sanitizer_diagk.txt 2>&1
tail -3 gpurun_out/c13_sanitizer_diagk.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/c13_diagk python tests/tools/prof_diagk.py > gpurun_out/c13_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c13_diagk.ncu-rep gpurun_out/c13_diagk_ncu_full > /dev/null 2>&1
grep -E "duration|issue_active|warp_instructions|stall_long" gpurun_out/c13_diagk_ncu_full.txt
