#!/bin/bash
# k_sample with the elements kept as doubles for the quick pass; A/B against QB200_SAMPLER_DOUBLES=0.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sampler.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c15_tests.txt 2>&1
tail -3 gpurun_out/c15_tests.txt
timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c15_prof_sampler.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c15_prof_sampler.txt | head -1
QB200_SAMPLER_DOUBLES=0 timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c15_prof_sampler_nodoubles.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c15_prof_sampler_nodoubles.txt | head -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/c15_sampler python tests/tools/prof_sampler.py > gpurun_out/c15_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c15_sampler.ncu-rep gpurun_out/c15_sampler_ncu_full > /dev/null 2>&1
grep -E "duration|issue_active|warp_instructions|stall_long|dram" gpurun_out/c15_sampler_ncu_full.txt
