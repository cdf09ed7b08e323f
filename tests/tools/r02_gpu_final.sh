#!/bin/bash
# Round 2, final 1-GPU pass: profiles for the final sources, bench, generator timing, sanitizer.
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:k_fused2d -c 3 -o gpurun_out/f_fused2d python tests/tools/prof_t2d.py 1 > gpurun_out/f_ncu_fused2d.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/f_fused2d.ncu-rep gpurun_out/r02_fused2d_ncu_full > /dev/null 2>&1
python tests/tools/fused2d_latest.py gpurun_out/r02_fused2d_ncu_full.json > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err
tail -c 300 gpurun_out/f_bench_1gpu.json; tail -3 gpurun_out/f_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
tail -c 300 gpurun_out/f_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_under_ncu.log 2>&1
python tests/tools/launch_summary.py gpurun_out/f_launches.csv > gpurun_out/r02_bench_launches_summary.txt 2>&1
rm -f gpurun_out/generate_timing.json
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client > gpurun_out/f_gen_a.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client_again > gpurun_out/f_gen_a2.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 256 --tag dim256_2clients > gpurun_out/f_gen_b.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 0 --tag heuristic_1client > gpurun_out/f_gen_c.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 0 --tag heuristic_2clients > gpurun_out/f_gen_d.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --prefetch 0 --tag dim256_1client_noprefetch > gpurun_out/f_gen_e.txt 2>&1
grep -h "generate_wall_s\|tag" gpurun_out/f_gen_*.txt
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_slices.py > gpurun_out/r02_compute_sanitizer_slices.txt 2>&1
tail -4 gpurun_out/r02_compute_sanitizer_slices.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/f_sampler python tests/tools/prof_sampler.py > gpurun_out/f_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/f_sampler.ncu-rep gpurun_out/r02_sampler_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused1d" -c 3 -o gpurun_out/f_fused1d python -m pytest tests/test_client_tail.py -x -q -m gpu -k single_launch > gpurun_out/f_ncu_fused1d.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/f_fused1d.ncu-rep gpurun_out/r02_fused1d_ncu_full > /dev/null 2>&1
python - <<'PY' > gpurun_out/f_min_cell.txt 2>&1
import numpy as np, sys
sys.path.insert(0, '.')
import bench, qunundrum_b200 as qb
from qunundrum_b200 import shard
ctx = qb.Context(0)
d, r = bench.synthetic_d_r(20482048)
P = qb.Parameters(2048, 1, d, r, 30)
coords = shard.enumerate_2d(2048)
cells, tp, te, fl = ctx.slice2d_batch(P, 0, True, 128, [c[0] for c in coords], [c[1] for c in coords])
a = np.abs(cells)
print("smallest non-zero |cell| of the T2D distribution:", a[a > 0].min(), "zeros:", int((a == 0).sum()),
      "smallest slice total:", float(tp.min()), "min |cell| / slice total:", float((a.min(axis=1) / tp.astype(float)).min()))
PY
cat gpurun_out/f_min_cell.txt
