#!/bin/bash
# Round 2: k_sample with the warp-cooperative block fetch, k_diagk with ld.shared constants and 8 CTAs/SM.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sampler.py tests/test_diagk.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c12_tests.txt 2>&1
tail -4 gpurun_out/c12_tests.txt
timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c12_prof_sampler.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c12_prof_sampler.txt | head -1
timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c12_prof_diagk.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c12_prof_diagk.txt | head -1
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_slices.py > gpurun_out/c12_sanitizer_slices.txt 2>&1
tail -3 gpurun_out/c12_sanitizer_slices.txt
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_diagk.py > gpurun_out/c12_sanitizer_diagk.txt 2>&1
tail -3 gpurun_out/c12_sanitizer_diagk.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/c12_sampler python tests/tools/prof_sampler.py > gpurun_out/c12_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c12_sampler.ncu-rep gpurun_out/c12_sampler_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/c12_diagk python tests/tools/prof_diagk.py > gpurun_out/c12_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c12_diagk.ncu-rep gpurun_out/c12_diagk_ncu_full > /dev/null 2>&1
