#!/bin/bash
# Round 2: k_exact_alpha_staged (coalesced staging through shared memory) against the per-thread kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c25_tests_exact.txt 2>&1
tail -3 gpurun_out/c25_tests_exact.txt
QB200_EXACT_STAGED=0 timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c25_tests_exact_unstaged.txt 2>&1
tail -3 gpurun_out/c25_tests_exact_unstaged.txt
for v in 1 0; do
  QB200_EXACT_STAGED=$v timeout 300 python tests/tools/prof_exact.py > gpurun_out/c25_prof_exact_staged$v.txt 2> gpurun_out/c25_prof_exact_staged$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c25_prof_exact_staged$v.txt").read().strip().splitlines()[-1])
print("staged=$v", {k: d[k] for k in ("value","ms_k_exact_alpha","ms_k_exact_jk")})
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_exact_alpha_staged$' -c 1 -o gpurun_out/c25_exact_alpha python tests/tools/prof_exact.py > gpurun_out/c25_ncu_exact_alpha.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c25_exact_alpha.ncu-rep gpurun_out/c25_exact_alpha_ncu_full > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_exact.py -x -q -m gpu -k "exact_arithmetic or bytes_to_k or small" > gpurun_out/c25_sanitizer_exact.txt 2>&1
tail -3 gpurun_out/c25_sanitizer_exact.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_exact.py -x -q -m gpu -k "exact_arithmetic" > gpurun_out/c25_racecheck_exact.txt 2>&1
tail -3 gpurun_out/c25_racecheck_exact.txt
