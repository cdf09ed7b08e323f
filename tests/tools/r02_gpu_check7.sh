#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/c7_share.txt
for share in 8 4 2; do
  for v in 1 0; do
    for o in 1 0; do
      echo "share=$share lean=$v overlap=$o" >> gpurun_out/c7_share.txt
      QB200_FUSED_LEAN=$v QB200_OVERLAP_CLASSES=$o timeout 300 python tests/tools/prof_t2d.py 200 128 $share >> gpurun_out/c7_share.txt 2>&1
    done
  done
done
cat gpurun_out/c7_share.txt
