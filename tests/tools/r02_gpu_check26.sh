#!/bin/bash
# Round 2: complete launch list of a bench run (all sections), and the 2-GPU bench line after the last changes.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_under_ncu.log 2>&1
python tests/tools/launch_summary.py gpurun_out/i_launches.csv > gpurun_out/i_bench_launches_summary.txt 2>&1
tail -5 gpurun_out/i_bench_launches_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/i_bench_2gpu.json 2> gpurun_out/i_bench_2gpu.err
tail -c 400 gpurun_out/i_bench_2gpu.json; tail -3 gpurun_out/i_bench_2gpu.err
