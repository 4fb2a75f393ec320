#!/bin/bash
# Development aid: builds libtracer variants with different -D flags for wavefront.cu into msc-futhark-ray-tracer_b200/variants/
# (run a tool against one with tools/run_with_lib.py <lib.so> <script> [args]).  usage: tools/build_variants.sh name1 "-Dflags1" name2 "-Dflags2" ...
set -e
cd "$(dirname "$0")/../msc-futhark-ray-tracer_b200"
make -s > /dev/null
mkdir -p variants build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc $NVFLAGS $flags -c csrc/wavefront.cu -o build/variants/wavefront_$name.o 2> build/variants/$name.ptxas.log
  /usr/local/cuda/bin/nvcc $ARCH -shared -ccbin /usr/bin/g++ -o variants/libtracer_$name.so build/abi.o build/variants/wavefront_$name.o build/lbvh.o
  echo "$name: $(grep -A2 '7k_shadeILi256' build/variants/$name.ptxas.log | grep -E 'registers' | head -1)"
done
