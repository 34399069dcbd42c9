#!/bin/bash
# Builds the CUDA library in-tree (sm_100a only).  Used by __graft_entry__.build().
#   megapath-nano_b200/libmpn_ssw.so        batched C ABI (include/mpn_ssw_batch.h) + legacy per-pair ABI (include/ssw.h)
#   megapath-nano_b200/realign/libssw.so    the name the reference's ctypes callers load (pyssw.py:4, build.sh:2 of the reference)
#   megapath-nano_b200/realign/realigner    the name realign_illumina_reads.py:29 loads (realign_reads / free_memory + the ssw_* symbols)
set -e
cd "$(dirname "$0")"
mkdir -p megapath-nano_b200/realign
#   megapath-nano_b200/realign/debruijn_graph   the name realign_illumina_reads.py:30 loads (get_consensus / free_memory), host only
g++ -O2 -std=c++17 -fPIC -Wall -shared -pthread -o megapath-nano_b200/realign/debruijn_graph megapath-nano_b200/csrc/debruijn_assemble.cpp
[ "$1" = "dbg" ] && exit 0
NVCC_FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $MPN_EXTRA_NVCC_FLAGS"
OBJ=build/obj${MPN_BUILD_TAG:+_$MPN_BUILD_TAG}
OUT=${MPN_SSW_OUT:-megapath-nano_b200/libmpn_ssw.so}
mkdir -p $OBJ
pids=()
for f in engine pool ssw_abi strip_inst_a strip_inst_b strip_inst_c strip_inst_d; do
    nvcc $NVCC_FLAGS -c -o $OBJ/$f.o megapath-nano_b200/csrc/$f.cu &
    pids+=($!)
done
# host-only sources: C++ front end (include/ssw_cpp.h) and region realigner (include/realigner.h)
for f in ssw_cpp_layer realign_region; do
    g++ -O2 -std=c++17 -fPIC -Wall -c -o $OBJ/$f.o megapath-nano_b200/csrc/$f.cpp &
    pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJ/engine.o $OBJ/pool.o $OBJ/ssw_abi.o $OBJ/strip_inst_a.o $OBJ/strip_inst_b.o $OBJ/strip_inst_c.o $OBJ/strip_inst_d.o $OBJ/ssw_cpp_layer.o $OBJ/realign_region.o
[ -n "$MPN_SSW_OUT" ] && exit 0
cp megapath-nano_b200/libmpn_ssw.so megapath-nano_b200/realign/libssw.so
# the reference builds `realigner` as a shared object without suffix (README.md:45 of the reference); same file, all symbols
cp megapath-nano_b200/libmpn_ssw.so megapath-nano_b200/realign/realigner
