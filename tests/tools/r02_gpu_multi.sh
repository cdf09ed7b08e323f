#!/bin/bash
# Round 2: N-GPU pass (N = number of GPUs of the box): D2H scaling microbenchmark, bench.py strong /
# weak / saturation, the reference generator as a farm with one client per GPU.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/tools/d2h_scaling.py > gpurun_out/m${N}_d2h.json 2> gpurun_out/m${N}_d2h.err
tail -c 1500 gpurun_out/m${N}_d2h.json
timeout 300 python tests/tools/d2h_scaling.py > gpurun_out/m${N}_d2h_single.json 2> gpurun_out/m${N}_d2h_single.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
tail -c 1200 gpurun_out/m${N}_bench.json; tail -5 gpurun_out/m${N}_bench.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-text --no-tau --no-sections > gpurun_out/m${N}_bench_n1.json 2> gpurun_out/m${N}_bench_n1.err
tail -c 300 gpurun_out/m${N}_bench_n1.json
rm -f gpurun_out/generate_timing.json
timeout 300 python tests/tools/generate_timing.py --clients $N --dim 256 --tag farm_dim256_${N}clients_${N}gpus > gpurun_out/m${N}_gen_a.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients $N --dim 0 --tag farm_heuristic_${N}clients_${N}gpus > gpurun_out/m${N}_gen_b.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 0 --tag farm_heuristic_1client > gpurun_out/m${N}_gen_c.txt 2>&1
QB200_PIN_VISIBLE=0 timeout 300 python tests/tools/generate_timing.py --clients $N --dim 0 --tag farm_heuristic_${N}clients_all_gpus_visible > gpurun_out/m${N}_gen_d.txt 2>&1
cp gpurun_out/generate_timing.json gpurun_out/m${N}_generate_timing.json
grep -h "generate_wall_s\|tag" gpurun_out/m${N}_gen_*.txt
