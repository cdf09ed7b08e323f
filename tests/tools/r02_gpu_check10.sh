#!/bin/bash
# A/B of compile-time variants of k_diagk (tests/tools/diagk_variants.py), one process each.
set -x
mkdir -p gpurun_out
for v in base unroll8 lds lds_unroll8 lds_unroll8_occ5 lds_occ8; do
  QB200_LIB=$PWD/qunundrum_b200/_variants/lib_$v.so timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c10_diagk_$v.txt 2>&1
  echo "$v: $(grep -o '"value": [0-9.]*, "unit": "samples/s", "ms": [0-9.]*' gpurun_out/c10_diagk_$v.txt | head -1)"
done
