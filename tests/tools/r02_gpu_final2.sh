#!/bin/bash
# Round 2, second final pass (after the sampler / diagk / sigma-optimal kernel changes): full gpu suite, smoke,
# default bench, launch list, profiles of the changed kernels, sanitizer.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/g_gpu_tests_full.txt 2>&1
tail -4 gpurun_out/g_gpu_tests_full.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.txt 2>&1
tail -2 gpurun_out/g_smoke.txt
timeout 900 python bench.py > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err
tail -c 300 gpurun_out/g_bench_default.json; tail -3 gpurun_out/g_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_under_ncu.log 2>&1
python tests/tools/launch_summary.py gpurun_out/g_launches.csv > gpurun_out/g_bench_launches_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/g_sampler python tests/tools/prof_sampler.py > gpurun_out/g_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/g_sampler.ncu-rep gpurun_out/g_sampler_ncu_full > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diagk$' -c 1 -o gpurun_out/g_diagk python tests/tools/prof_diagk.py > gpurun_out/g_ncu_diagk.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/g_diagk.ncu-rep gpurun_out/g_diagk_ncu_full > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_slices.py > gpurun_out/g_sanitizer_slices.txt 2>&1
tail -3 gpurun_out/g_sanitizer_slices.txt
timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_diagk.py > gpurun_out/g_sanitizer_diagk.txt 2>&1
tail -3 gpurun_out/g_sanitizer_diagk.txt
timeout 300 python tests/tools/tau_diagonal_timing.py > gpurun_out/g_tau_diagonal.txt 2>&1
tail -3 gpurun_out/g_tau_diagonal.txt
