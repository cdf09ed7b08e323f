#!/bin/bash
# Round 2: the exact samplers after the small-|log alpha| support and the single-pass k_exact_alpha.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c24_tests_exact.txt 2>&1
tail -3 gpurun_out/c24_tests_exact.txt
timeout 900 python -m pytest tests/test_estimate_runs_end_to_end.py tests/test_diagk.py -x -q -m gpu -k "diagonal or dropin" > gpurun_out/c24_tests_diag.txt 2>&1
tail -3 gpurun_out/c24_tests_diag.txt
timeout 300 python tests/tools/prof_exact.py --ref > gpurun_out/c24_prof_exact.txt 2> gpurun_out/c24_prof_exact.err
tail -c 1200 gpurun_out/c24_prof_exact.txt; tail -3 gpurun_out/c24_prof_exact.err
timeout 300 python tests/tools/tau_diagonal_timing.py > gpurun_out/c24_tau_diagonal.txt 2>&1
tail -2 gpurun_out/c24_tau_diagonal.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_exact_alpha$' -c 1 -o gpurun_out/c24_exact_alpha python tests/tools/prof_exact.py > gpurun_out/c24_ncu_exact_alpha.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c24_exact_alpha.ncu-rep gpurun_out/c24_exact_alpha_ncu_full > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_exact.py -x -q -m gpu -k "exact_arithmetic or bytes_to_k or small" > gpurun_out/c24_sanitizer_exact.txt 2>&1
tail -4 gpurun_out/c24_sanitizer_exact.txt
