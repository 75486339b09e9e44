#!/bin/bash
# Quick SASS inspection of the lean tile kernel (no GPU needed): compiles tqb_tile.cu with only the lean
# instantiations, disassembles with line info, and prints per-source-line static instruction counts.
set -e
cd "$(dirname "$0")/.."
OUT=${OUT:-/tmp/sass_lean}
mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTQB_LEAN_ONLY -Xptxas=-v -cubin \
  tyxonq_b200/csrc/tqb_tile.cu -o $OUT/lean.cubin 2> $OUT/ptxas.log || { tail -30 $OUT/ptxas.log; exit 1; }
grep -A2 "lean_kernel" $OUT/ptxas.log | grep -v "^--" | grep "spill\|Used" || true
nvdisasm -g $OUT/lean.cubin > $OUT/lean.dis
echo "disassembly: $OUT/lean.dis"
