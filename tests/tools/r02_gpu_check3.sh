#!/bin/bash
# Round 2, third GPU pass: speculation fix, delayed eager server context, look-ahead export,
# bucketed sampler (A/B), generator timing.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler.py tests/test_client_tail.py -x -q -m gpu > gpurun_out/c3_tests_a.txt 2>&1
tail -4 gpurun_out/c3_tests_a.txt
timeout 900 python -m pytest tests/test_generators_end_to_end.py tests/test_estimate_runs_end_to_end.py -x -q -m gpu -k "prefetching or estimate" > gpurun_out/c3_tests_b.txt 2>&1
tail -4 gpurun_out/c3_tests_b.txt
QB200_SAMPLER_BUCKETS=1 timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c3_sampler_buckets.json 2> gpurun_out/c3_sampler_buckets.err
QB200_SAMPLER_BUCKETS=0 timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c3_sampler_draworder.json 2> gpurun_out/c3_sampler_draworder.err
python - <<'PY'
import json
for f in ("c3_sampler_buckets", "c3_sampler_draworder"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "%.4g samples/s" % d["value"], d["ms"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
rm -f gpurun_out/generate_timing.json
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client > gpurun_out/c3_gen_a.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 256 --tag dim256_2clients > gpurun_out/c3_gen_b.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 2 --dim 0 --tag heuristic_2clients > gpurun_out/c3_gen_c.txt 2>&1
timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 0 --tag heuristic_1client > gpurun_out/c3_gen_d.txt 2>&1
QB200_EAGER_DELAY_MS=0 timeout 300 python tests/tools/generate_timing.py --clients 1 --dim 256 --tag dim256_1client_eager_at_once > gpurun_out/c3_gen_e.txt 2>&1
grep -h "generate_wall_s\|tag" gpurun_out/c3_gen_*.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sample_cells|k_sample_slices" -c 2 -o gpurun_out/c3_sampler python tests/tools/prof_sampler.py > gpurun_out/c3_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c3_sampler.ncu-rep gpurun_out/r02_sampler_ncu_full > /dev/null 2>&1
ls -la gpurun_out | tail -20
