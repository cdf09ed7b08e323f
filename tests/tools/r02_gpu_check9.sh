#!/bin/bash
# Round 2, after the convergent k_sample, the padded mul_columns of k_diagk and the leaner k_so_fast.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sampler.py tests/test_diagk.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c9_tests_a.txt 2>&1
tail -4 gpurun_out/c9_tests_a.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_client_tail.py -x -q -m gpu -k "sigma or optimal or bench" > gpurun_out/c9_tests_b.txt 2>&1
tail -4 gpurun_out/c9_tests_b.txt
timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c9_prof_sampler.txt 2>&1
tail -5 gpurun_out/c9_prof_sampler.txt
timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c9_prof_diagk.txt 2>&1
tail -5 gpurun_out/c9_prof_diagk.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c9_bench_1gpu.json 2> gpurun_out/c9_bench_1gpu.err
tail -c 300 gpurun_out/c9_bench_1gpu.json; tail -3 gpurun_out/c9_bench_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/c9_sampler python tests/tools/prof_sampler.py > gpurun_out/c9_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c9_sampler.ncu-rep gpurun_out/c9_sampler_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/c9_diagk python tests/tools/prof_diagk.py > gpurun_out/c9_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c9_diagk.ncu-rep gpurun_out/c9_diagk_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_so_fast' -c 1 -o gpurun_out/c9_so_fast python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "closed_form" > gpurun_out/c9_ncu_so_fast.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c9_so_fast.ncu-rep gpurun_out/c9_so_fast_ncu_full > /dev/null 2>&1
ls -la gpurun_out | tail -15
