#!/bin/bash
# Round 2, first GPU pass: new tests, bench (N = 1), generator timing, launch list + ncu capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_client_tail.py tests/test_gpu_parity.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/c1_tests_a.txt 2>&1
tail -5 gpurun_out/c1_tests_a.txt
timeout 900 python -m pytest tests/test_generators_end_to_end.py -x -q -m gpu -k "prefetching or matches_reference" > gpurun_out/c1_tests_b.txt 2>&1
tail -5 gpurun_out/c1_tests_b.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_1gpu.json 2> gpurun_out/c1_bench_1gpu.err
tail -c 600 gpurun_out/c1_bench_1gpu.json; tail -5 gpurun_out/c1_bench_1gpu.err
rm -f gpurun_out/generate_timing.json
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client > gpurun_out/c1_gen_a.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 256 --tag dim256_2clients > gpurun_out/c1_gen_b.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 0 --tag heuristic_2clients > gpurun_out/c1_gen_c.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --prefetch 0 --tag dim256_1client_noprefetch > gpurun_out/c1_gen_d.txt 2>&1
grep -h "generate_wall_s\|tag" gpurun_out/c1_gen_*.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text --no-tau > gpurun_out/c1_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:k_fused2d -c 3 -o gpurun_out/c1_fused2d python tests/tools/prof_t2d.py 1 > gpurun_out/c1_ncu_fused2d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused1d|k_collapse|k_scale_x87" -c 8 -o gpurun_out/c1_new_kernels python -m pytest tests/test_client_tail.py -x -q -m gpu > gpurun_out/c1_ncu_new.log 2>&1
ls -la gpurun_out/ | tail -30
