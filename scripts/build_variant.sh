#!/bin/bash
# A/B variants of one kernel file: build_variants/<name>.so = the library with <file>.cu recompiled with extra flags.
#   scripts/build_variant.sh pcs_kernels q2 "-DLB_QROWS=2 -DLB_Q_MINBLOCKS=4" [name "flags" ...]
#   LUMINAIR_B200_LIB=build_variants/q2.so python scripts/time_stage.py ...
# (build_variants/ is git- and gpurun-ignored: copy a variant somewhere under scripts/ubench/ to take it to the GPU box.)
set -e
file=$1; shift
cd "$(dirname "$0")/../luminair_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../build_variants
objs=""
for o in capi cfft merkle pcs_kernels air_kernels trace_kernels prover; do
  [ "$o" = "$file" ] || objs="$objs $o.o"
done
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $flags -c $file.cu -o /tmp/${file}_$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build_variants/$name.so $objs /tmp/${file}_$name.o -lcudart
done
ls ../../build_variants
