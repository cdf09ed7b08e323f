#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_client_tail.py -x -q -m gpu > gpurun_out/c8_tests.txt 2>&1
tail -5 gpurun_out/c8_tests.txt
rm -f gpurun_out/c8_share.txt
for share in 8 4 2 1; do
  for g in 1 0; do
    echo "share=$share graphs=$g" >> gpurun_out/c8_share.txt
    QB200_GRAPHS=$g timeout 300 python tests/tools/prof_t2d.py 200 128 $share >> gpurun_out/c8_share.txt 2>&1
  done
done
echo "share=8 graphs=1 lean=0" >> gpurun_out/c8_share.txt
QB200_FUSED_LEAN=0 timeout 300 python tests/tools/prof_t2d.py 200 128 8 >> gpurun_out/c8_share.txt 2>&1
echo "share=1 graphs=1 lean=0" >> gpurun_out/c8_share.txt
QB200_FUSED_LEAN=0 timeout 300 python tests/tools/prof_t2d.py 200 128 1 >> gpurun_out/c8_share.txt 2>&1
echo "share=8 graphs=1 D=256" >> gpurun_out/c8_share.txt
timeout 300 python tests/tools/prof_t2d.py 100 256 8 >> gpurun_out/c8_share.txt 2>&1
cat gpurun_out/c8_share.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-text --no-tau --no-sections --no-cpu-baseline > gpurun_out/c8_bench_1gpu.json 2> gpurun_out/c8_bench_1gpu.err
tail -c 500 gpurun_out/c8_bench_1gpu.json; tail -3 gpurun_out/c8_bench_1gpu.err
