#!/bin/bash
# A/B build of the library with the tc4 timing ablations compiled in (results are wrong by construction when an
# ablation is switched on through LRR_ABL_BITS / LRR_ABL_STREAM / LRR_ABL_CONTIG):
#   scratch/build_abl.sh && LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so LRR_ABL_BITS=2 scratch/sustain.sh noMMA
cd "$(dirname "$0")/.."
mkdir -p scratch/abl
SRC="$(ls hail_b200/csrc/*.cu | tr '
' ' ')"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
nvcc $FLAGS -DLRR_TC4_ABLATIONS=1 -DLRR_TUNING=1 "$@" -o scratch/abl/tc4_abl.so $SRC && ls -la scratch/abl
