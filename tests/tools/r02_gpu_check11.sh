#!/bin/bash
# A/B of compile-time variants (tests/tools/kernel_variants.py) of k_diagk and k_sample, one process
# each; then the sampler / diagk / dropin tests with the default build.
set -x
mkdir -p gpurun_out
for v in diagk_occ6 diagk_occ8 diagk_occ9 diagk_occ10 diagk_occ12 diagk_occ8_unroll8 diagk_occ8_unroll2; do
  QB200_LIB=$PWD/qunundrum_b200/_variants/lib_$v.so timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c11_$v.txt 2>&1
  echo "$v: $(grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c11_$v.txt | head -1)"
done
for v in sample_occ6 sample_occ7 sample_occ8 sample_occ10 sample_occ12; do
  QB200_LIB=$PWD/qunundrum_b200/_variants/lib_$v.so timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c11_$v.txt 2>&1
  echo "$v: $(grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c11_$v.txt | head -1)"
done
timeout 1500 python -m pytest tests/test_sampler.py tests/test_diagk.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c11_tests.txt 2>&1
tail -4 gpurun_out/c11_tests.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/c11_sampler python tests/tools/prof_sampler.py > gpurun_out/c11_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c11_sampler.ncu-rep gpurun_out/c11_sampler_ncu_full > /dev/null 2>&1
