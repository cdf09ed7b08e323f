#!/bin/bash
# Round 2: the exact samplers end to end -- parity tests, the diagonal tau drop-in over them, bench section,
# ncu captures of the two kernels, sanitizer, smoke.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c23_tests_exact.txt 2>&1
tail -3 gpurun_out/c23_tests_exact.txt
timeout 900 python -m pytest tests/test_diagk.py tests/test_estimate_runs_end_to_end.py -x -q -m gpu > gpurun_out/c23_tests_diagk.txt 2>&1
tail -3 gpurun_out/c23_tests_diagk.txt
timeout 300 python tests/tools/prof_exact.py --ref > gpurun_out/c23_prof_exact.txt 2> gpurun_out/c23_prof_exact.err
tail -c 1500 gpurun_out/c23_prof_exact.txt; tail -3 gpurun_out/c23_prof_exact.err
timeout 300 python tests/tools/tau_diagonal_timing.py > gpurun_out/c23_tau_diagonal.txt 2>&1
tail -3 gpurun_out/c23_tau_diagonal.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_exact_jk$' -c 1 -o gpurun_out/c23_exact_jk python tests/tools/prof_exact.py > gpurun_out/c23_ncu_exact_jk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c23_exact_jk.ncu-rep gpurun_out/c23_exact_jk_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_exact_alpha$' -c 1 -o gpurun_out/c23_exact_alpha python tests/tools/prof_exact.py > gpurun_out/c23_ncu_exact_alpha.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c23_exact_alpha.ncu-rep gpurun_out/c23_exact_alpha_ncu_full > /dev/null 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c23_smoke.txt 2>&1
tail -2 gpurun_out/c23_smoke.txt
