#!/bin/bash
# Where the wall clock of estimate_runs_diagonal_distribution (drop-in flavour, m = 64) goes with 16 and
# with 4 client ranks on one GPU: QB200_DROPIN_STATS of the first ranks.  bash tests/tools/e2e_diag_ranks.sh
B=$PWD/integration/_build
T=$(mktemp -d); cd $T; mkdir distributions
QB200_DEVICE=0 QB200_TEXT_DEVICE=0 $B/minimpirun -np 3 $B/gpu/generate_diagonal_distribution -det -dim 128 -eta-bound 8 64 6 2 > gen.log 2>&1 || tail -5 gen.log
ls distributions
for np in 17 5; do
  mkdir run$np; cd run$np; ln -s ../distributions distributions
  t0=$(date +%s%N)
  QB200_DEVICE=0 QB200_TEXT_DEVICE=0 QB200_DROPIN_STATS=1 $B/minimpirun -np $np $B/gpu/estimate_runs_diagonal_distribution distributions/diagonal-distribution-det-dim-128-m-64-sigma-6-s-2.txt > out.log 2> err.log
  t1=$(date +%s%N)
  echo "np=$np wall $(( (t1 - t0) / 1000000 )) ms"
  grep "^m:" out.log; grep "drop-in" err.log | head -3
  cd ..
done
