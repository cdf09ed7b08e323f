#!/bin/bash
# Round 2: the exact samplers on the GPU (parity tests, sanitizer on the same tests' small cases).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exact.py -x -q -m gpu > gpurun_out/c22_tests.txt 2>&1
tail -15 gpurun_out/c22_tests.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_exact.py -x -q -m gpu -k "exact_arithmetic or bytes_to_k" > gpurun_out/c22_sanitizer_exact.txt 2>&1
tail -4 gpurun_out/c22_sanitizer_exact.txt
