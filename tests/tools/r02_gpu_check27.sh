#!/bin/bash
# Round 2: k_exact_jk eight columns per pass against four.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c27_tests_exact.txt 2>&1
tail -3 gpurun_out/c27_tests_exact.txt
for v in 8 4 8 4; do
  QB200_EXACT_COLUMNS=$v timeout 300 python tests/tools/prof_exact.py > gpurun_out/c27_prof_exact_cols$v.txt 2> gpurun_out/c27_prof_exact_cols$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c27_prof_exact_cols$v.txt").read().strip().splitlines()[-1])
print("columns=$v", {k: d[k] for k in ("value","ms_k_exact_alpha","ms_k_exact_jk")})
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_exact_jk' -c 1 -o gpurun_out/c27_exact_jk python tests/tools/prof_exact.py > gpurun_out/c27_ncu_exact_jk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c27_exact_jk.ncu-rep gpurun_out/c27_exact_jk_ncu_full > /dev/null 2>&1
