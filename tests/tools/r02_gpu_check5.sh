#!/bin/bash
# Round 2, fifth GPU pass: closed-form sigma-optimal walk, warp-per-slice final kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_client_tail.py tests/test_sampler.py -x -q -m gpu -s > gpurun_out/c5_tests_a.txt 2>&1
tail -6 gpurun_out/c5_tests_a.txt; grep "sigma-optimal closed form\|bench configuration" gpurun_out/c5_tests_a.txt
timeout 900 python -m pytest tests/test_generators_end_to_end.py -x -q -m gpu -k "matches_reference or prefetching" > gpurun_out/c5_tests_b.txt 2>&1
tail -3 gpurun_out/c5_tests_b.txt
timeout 300 python tests/tools/prof_t2d.py 20 128 > gpurun_out/c5_t2d.txt 2>&1; cat gpurun_out/c5_t2d.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-text --no-tau > gpurun_out/c5_bench_1gpu.json 2> gpurun_out/c5_bench_1gpu.err
tail -3 gpurun_out/c5_bench_1gpu.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/c5_bench_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms %.4f frac %s" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["frac"]))
print(json.dumps(b["sections"]["sigma_optimal"], indent=1)[:1500])
print(json.dumps(b["saturation"])[:600])
PY
