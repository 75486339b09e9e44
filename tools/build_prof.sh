#!/bin/bash
# Profiling build (TQB_PROFILE_SWITCHES: TQB_DBG bits 8/16/32 skip arithmetic / tile loads / tile stores of the chain
# sweeps; results are wrong with any bit set).  Only the lean + default TMA kernels are needed, so this is the quick
# lean-only build plus the reductions.  Use with TQB_LIB=tools/libtqb_prof.so.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DTQB_LEAN_ONLY -DTQB_PROFILE_SWITCHES"
nvcc $F -c tyxonq_b200/csrc/tqb_tile.cu -o /tmp/prof_tile.o &
[ -f tyxonq_b200/build/tqb_reduce.cu.o ] || python -m tyxonq_b200.build
wait
nvcc -shared -o tools/libtqb_prof.so /tmp/prof_tile.o tyxonq_b200/build/tqb_reduce.cu.o -lcudart
echo built tools/libtqb_prof.so
