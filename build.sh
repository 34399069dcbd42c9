#!/bin/bash
# Builds the CUDA library in-tree (sm_100a only).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -o megapath-nano_b200/libmpn_ssw.so megapath-nano_b200/csrc/engine.cu
