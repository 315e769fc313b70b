#!/bin/bash
# Build a variant of libhybridq_b200.so with extra -D flags (compile-time experiments):
#   tools/build_variant.sh NAME -DHQ_K1D_BLOCKS=3 ...   ->  hybridq_b200/lib/variants/libhybridq_b200_NAME.so
# Use it with  HYBRIDQ_B200_LIB=hybridq_b200/lib/variants/libhybridq_b200_NAME.so python tools/...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=hybridq_b200/lib/variants; obj=$out/obj_$name
mkdir -p $obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -Xptxas -v -split-compile 0"
for f in hq_kernels hq_abi; do
  $NVCC $FLAGS "$@" -c hybridq_b200/csrc/$f.cu -o $obj/$f.o 2> $obj/$f.ptxas.log || (cat $obj/$f.ptxas.log; exit 1)
done
/usr/bin/g++ -O2 -std=c++17 -fPIC "$@" -c hybridq_b200/csrc/hq_plan.cpp -o $obj/hq_plan.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $out/libhybridq_b200_$name.so $obj/hq_kernels.o $obj/hq_abi.o $obj/hq_plan.o -cudart static -Xlinker --exclude-libs,ALL
echo built $out/libhybridq_b200_$name.so
