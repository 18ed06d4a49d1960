#!/bin/bash
# A/B variants of pcs_kernels.cu: build_variants/pcs_<name>.so  (usage: build_pcs_variants.sh name "-DLB_QROWS=2 ..." [name flags]...)
set -e
cd "$(dirname "$0")/../luminair_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../build_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $flags -c pcs_kernels.cu -o /tmp/pcs_$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build_variants/pcs_$name.so capi.o cfft.o merkle.o /tmp/pcs_$name.o air_kernels.o trace_kernels.o prover.o -lcudart
  cuobjdump -res-usage ../../build_variants/pcs_$name.so 2>/dev/null | grep -A1 "quotients_kernelILi2E" | grep REG | head -1
done
ls ../../build_variants
