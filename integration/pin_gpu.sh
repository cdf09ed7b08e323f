#!/bin/sh
# One visible GPU per MPI rank: client k sees only GPU (k - 1) mod N, the server (rank 0, which
# exports through the text drop-in) GPU 0. A process that sees a single device initialises
# CUDA several times faster on an 8-GPU box than one that maps all eight.
#   mpirun -np 9 integration/pin_gpu.sh generate_distribution ...
r=${QB200_MINIMPI_RANK:-${OMPI_COMM_WORLD_LOCAL_RANK:-${SLURM_LOCALID:-0}}}
n=${QB200_GPUS:-$(nvidia-smi -L | wc -l)}
if [ "$r" -gt 0 ]; then g=$(( (r - 1) % n )); else g=0; fi
CUDA_VISIBLE_DEVICES=$g QB200_DEVICE=0 exec "$@"
