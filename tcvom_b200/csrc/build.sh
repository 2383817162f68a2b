#!/bin/bash
# Builds libtcvom_b200.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
OUT=../lib
mkdir -p $OUT
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
SRCS=$(ls *.cu)
OBJS=""
pids=()
for s in $SRCS; do
  o=$OUT/${s%.cu}.o
  OBJS="$OBJS $o"
  if [ ! -f $o ] || [ $s -nt $o ] || [ common.cuh -nt $o ] || [ tc_common.cuh -nt $o ] || [ tc_epilogue.cuh -nt $o ] || [ fba_body.h -nt $o ] || [ ../../include/tcvom_b200.h -nt $o ]; then
    $NVCC $FLAGS ${EXTRA_FLAGS} -c $s -o $o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT/libtcvom_b200.so $OBJS -gencode arch=compute_100a,code=sm_100a
echo built $OUT/libtcvom_b200.so
