#!/bin/bash
# Round 2: bench.py at N = 1, 2, 4, 8 on one 8-GPU box (what the driver's SCALE run does).
set -x
mkdir -p gpurun_out
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s${N}_bench.json 2> gpurun_out/s${N}_bench.err
  tail -c 300 gpurun_out/s${N}_bench.json; tail -2 gpurun_out/s${N}_bench.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-text --no-tau --no-sections > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -c 300 gpurun_out/s1_bench.json
