#!/bin/bash
# Builds libsimq.so for sm_100a (in-tree; the .so travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p build
pids=()
for f in api kernels_elem conv_fma conv_umma; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ kernels.h -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ ../../include/simq.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libsimq.so build/api.o build/kernels_elem.o build/conv_fma.o build/conv_umma.o -lcudart
echo "built $(cd .. && pwd)/libsimq.so"
