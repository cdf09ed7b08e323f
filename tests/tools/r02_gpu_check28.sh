#!/bin/bash
# Round 2: the exact sampler tests after the last changes (mirror of src/sample.h, eight-column products).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py tests/test_diagk.py -x -q -m gpu > gpurun_out/c28_tests.txt 2>&1
tail -3 gpurun_out/c28_tests.txt
