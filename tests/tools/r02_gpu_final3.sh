#!/bin/bash
# Round 2, third final pass (after the exact samplers): full gpu suite, smoke, default bench, reference arm,
# launch list.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/h_gpu_tests_full.txt 2>&1
tail -4 gpurun_out/h_gpu_tests_full.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.txt 2>&1
tail -2 gpurun_out/h_smoke.txt
timeout 900 python bench.py > gpurun_out/h_bench_default.json 2> gpurun_out/h_bench_default.err
tail -c 300 gpurun_out/h_bench_default.json; tail -3 gpurun_out/h_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/h_bench_reference.json 2> gpurun_out/h_bench_reference.err
tail -c 300 gpurun_out/h_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_under_ncu.log 2>&1
python tests/tools/launch_summary.py gpurun_out/h_launches.csv > gpurun_out/h_bench_launches_summary.txt 2>&1
