#!/bin/bash
# Round 2, sixth GPU pass: lean fused path (merged prologue, in-kernel summaries) A/B + tests.
set -x
mkdir -p gpurun_out
for v in 1 0; do
  QB200_FUSED_LEAN=$v timeout 300 python tests/tools/prof_t2d.py 30 128 > gpurun_out/c6_t2d_lean$v.txt 2>&1
  QB200_FUSED_LEAN=$v timeout 300 python tests/tools/prof_t2d.py 10 256 >> gpurun_out/c6_t2d_lean$v.txt 2>&1
  QB200_FUSED_LEAN=$v timeout 300 python tests/tools/prof_t2d.py 30 128 >> gpurun_out/c6_t2d_lean$v.txt 2>&1
done
cat gpurun_out/c6_t2d_lean1.txt gpurun_out/c6_t2d_lean0.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_client_tail.py tests/test_dropin_gpu.py tests/test_fuzz_reference.py -x -q -m gpu > gpurun_out/c6_tests_a.txt 2>&1
tail -4 gpurun_out/c6_tests_a.txt
QB200_FUSED_LEAN=0 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/c6_tests_b.txt 2>&1
tail -3 gpurun_out/c6_tests_b.txt
timeout 900 python -m pytest tests/test_generators_end_to_end.py -x -q -m gpu -k "matches_reference or prefetching" > gpurun_out/c6_tests_c.txt 2>&1
tail -3 gpurun_out/c6_tests_c.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-text --no-tau --no-sections --no-cpu-baseline > gpurun_out/c6_bench_1gpu.json 2> gpurun_out/c6_bench_1gpu.err
tail -c 700 gpurun_out/c6_bench_1gpu.json
timeout 600 compute-sanitizer --tool memcheck python tests/tools/sanitize_slices.py > gpurun_out/c6_sanitizer.txt 2>&1; tail -3 gpurun_out/c6_sanitizer.txt
timeout 600 compute-sanitizer --tool racecheck python tests/tools/sanitize_slices.py > gpurun_out/c6_racecheck.txt 2>&1; tail -3 gpurun_out/c6_racecheck.txt
