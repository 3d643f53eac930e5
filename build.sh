#!/bin/bash
# Builds libb200dsp.so (the product: CUDA kernels + C-ABI) in-tree for sm_100a.
# nvcc cross-compiles without a GPU; the .so travels to the GPU box with the snapshot.
set -e
cd "$(dirname "$0")"
SRC=gstreamer_b200/csrc
OUT=gstreamer_b200/libb200dsp.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-ffp-contract=off -Xptxas -v"
mkdir -p build
OBJS=""
for f in $SRC/*.cu $SRC/*.cpp; do
  o=build/$(basename "$f").o
  if [ ! -f "$o" ] || [ -n "$(find $SRC include -newer "$o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' -o -name '*.cpp' \) | head -1)" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS -x cu -c "$f" -o "$o" 2> "build/$(basename "$f").ptxas.log" || { cat "build/$(basename "$f").ptxas.log"; exit 1; }
  fi
  OBJS="$OBJS $o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS -cudart static
echo "built $OUT"
