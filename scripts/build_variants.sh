#!/bin/bash
# Build A/B variants of the CFFT kernels: build_variants/v<N>.so  (usage: build_variants.sh "0 1 2 3" [extra nvcc flags])
set -e
cd "$(dirname "$0")/../luminair_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../build_variants
for v in $1; do
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DLB_BFLY_VARIANT=$v $2 -c cfft.cu -o /tmp/cfft_v$v.o &
done
wait
for v in $1; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build_variants/v$v.so capi.o /tmp/cfft_v$v.o merkle.o pcs_kernels.o air_kernels.o prover.o -lcudart
done
ls -la ../../build_variants
