#!/bin/bash
# A/B builds of the sweep kernel with ablation switches (timing only; results are wrong by construction)
cd "$(dirname "$0")/.."
SRC="hail_b200/csrc/abi.cu hail_b200/csrc/pack.cu hail_b200/csrc/fp64_kernel.cu hail_b200/csrc/stats_epilogue.cu hail_b200/csrc/tc_kernel.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
build() { name=$1; shift; nvcc $FLAGS "$@" -o scratch/abl/$name.so $SRC & }
build no_bpanel -DLRR_ABL_NO_BPANEL=1
build no_geno -DLRR_ABL_NO_GENO=1
build no_mma -DLRR_ABL_NO_MMA=1
build no_sttm -DLRR_ABL_NO_STTM=1
wait
build no_unpack -DLRR_ABL_NO_UNPACK=1
build mma_j2 -DLRR_ABL_MMA_J=2
build no_bp_geno -DLRR_ABL_NO_BPANEL=1 -DLRR_ABL_NO_GENO=1
build mma_n16 -DLRR_ABL_MMA_N=16
wait
ls -la scratch/abl
