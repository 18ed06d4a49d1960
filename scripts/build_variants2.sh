#!/bin/bash
# build_variants2.sh name "flags" [name "flags" ...]
set -e
cd "$(dirname "$0")/../luminair_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../build_variants
args=("$@")
for ((i=0;i<${#args[@]};i+=2)); do
  n=${args[i]}; f=${args[i+1]}
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $f -c cfft.cu -o /tmp/cfft_$n.o &
done
wait
for ((i=0;i<${#args[@]};i+=2)); do
  n=${args[i]}
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build_variants/$n.so capi.o /tmp/cfft_$n.o merkle.o pcs_kernels.o air_kernels.o prover.o -lcudart
done
