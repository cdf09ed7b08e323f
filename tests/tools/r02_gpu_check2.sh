#!/bin/bash
# Round 2, second GPU pass: the fixes of pass 1 (batch-invariant plans, pageable prefetch cache,
# eager server context, persistent k_diagk), profile refresh, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generators_end_to_end.py -x -q -m gpu -k "prefetching" > gpurun_out/c2_tests_prefetch.txt 2>&1
tail -5 gpurun_out/c2_tests_prefetch.txt
timeout 900 python -m pytest tests/test_diagk.py tests/test_gpu_parity.py tests/test_client_tail.py -x -q -m gpu > gpurun_out/c2_tests_a.txt 2>&1
tail -3 gpurun_out/c2_tests_a.txt
# profile of the hot kernel for this source, then the bench that quotes it
timeout 600 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:k_fused2d -c 3 -o gpurun_out/c2_fused2d python tests/tools/prof_t2d.py 1 > gpurun_out/c2_ncu_fused2d.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c2_fused2d.ncu-rep gpurun_out/r02_fused2d_ncu_full > /dev/null 2>&1
python tests/tools/fused2d_latest.py gpurun_out/r02_fused2d_ncu_full.json > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench_1gpu.json 2> gpurun_out/c2_bench_1gpu.err
tail -c 400 gpurun_out/c2_bench_1gpu.json; tail -5 gpurun_out/c2_bench_1gpu.err
rm -f gpurun_out/generate_timing.json
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client > gpurun_out/c2_gen_a.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 256 --tag dim256_2clients > gpurun_out/c2_gen_b.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 0 --tag heuristic_2clients > gpurun_out/c2_gen_c.txt 2>&1
GLIBC_TUNABLES=glibc.malloc.mmap_threshold=1073741824:glibc.malloc.top_pad=268435456 timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 256 --tag dim256_2clients_malloc_tunables > gpurun_out/c2_gen_d.txt 2>&1
grep -h "generate_wall_s\|tag" gpurun_out/c2_gen_*.txt
# the diagonal k sampler with per-CTA scratch
timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c2_diagk.json 2> gpurun_out/c2_diagk.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/c2_diagk python tests/tools/prof_diagk.py > gpurun_out/c2_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c2_diagk.ncu-rep gpurun_out/r02_diagk_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused1d" -c 3 -o gpurun_out/c2_fused1d python -m pytest tests/test_client_tail.py -x -q -m gpu -k single_launch > gpurun_out/c2_ncu_fused1d.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c2_fused1d.ncu-rep gpurun_out/r02_fused1d_ncu_full > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text --no-tau > gpurun_out/c2_bench_under_ncu.log 2>&1
ls -la gpurun_out/ | tail -40
