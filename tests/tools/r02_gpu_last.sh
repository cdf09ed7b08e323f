#!/bin/bash
# Round 2: what the driver runs at round end -- the whole gpu suite, smoke(), bench.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_gpu_tests_full.txt 2>&1
tail -5 gpurun_out/r02_gpu_tests_full.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
tail -3 gpurun_out/r02_smoke.txt
timeout 900 python bench.py > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err
tail -c 300 gpurun_out/l_bench_default.json; tail -3 gpurun_out/l_bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/l_bench_ref.json 2>&1
tail -c 300 gpurun_out/l_bench_ref.json
