#!/bin/bash
# Builds the CUDA library in-tree (sm_100a only).  Used by __graft_entry__.build().
#   megapath-nano_b200/libmpn_ssw.so        batched C ABI (include/mpn_ssw_batch.h) + legacy per-pair ABI (include/ssw.h)
#   megapath-nano_b200/realign/libssw.so    the name the reference's ctypes callers load (pyssw.py:4, build.sh:2 of the reference)
set -e
cd "$(dirname "$0")"
mkdir -p megapath-nano_b200/realign
NVCC_FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
nvcc $NVCC_FLAGS -o megapath-nano_b200/libmpn_ssw.so megapath-nano_b200/csrc/engine.cu megapath-nano_b200/csrc/ssw_abi.cu
cp megapath-nano_b200/libmpn_ssw.so megapath-nano_b200/realign/libssw.so
