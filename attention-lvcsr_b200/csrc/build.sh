#!/bin/bash
# Build liblvsr_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I."
OBJS=()
for f in gemm gemm_tc bigru bigru_bwd attention decoder dec_scan train search api; do
  stale=0
  for h in kernels.h common.cuh attention_row.cuh model.h train_kernels.cuh ../../include/lvsr_b200.h; do
    if [ "$h" -nt "$f.o" ]; then stale=1; fi
  done
  if [ ! -f "$f.o" ] || [ "$f.cu" -nt "$f.o" ] || [ $stale = 1 ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS "$@" -c "$f.cu" -o "$f.o"
  fi
  OBJS+=("$f.o")
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o liblvsr_b200.so "${OBJS[@]}" -lcudart
echo "built $(pwd)/liblvsr_b200.so"
