#!/bin/bash
# Round 2: last confirmation of the sampler suites and smoke with the final sources.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_exact.py tests/test_diagk.py tests/test_estimate_runs_end_to_end.py tests/test_sampler.py -x -q -m gpu > gpurun_out/c29_tests.txt 2>&1
tail -3 gpurun_out/c29_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c29_smoke.txt 2>&1
tail -2 gpurun_out/c29_smoke.txt
timeout 300 python tests/tools/prof_exact.py --ref > gpurun_out/c29_prof_exact.txt 2> gpurun_out/c29_prof_exact.err
tail -c 600 gpurun_out/c29_prof_exact.txt
