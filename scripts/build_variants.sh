#!/bin/bash
# Builds A/B variants of libktn.so (ktn_kernels.cu and ktn_comm.cu are recompiled with the flags) into build/variants/ for scripts/ab_time.py.
# usage: scripts/build_variants.sh name1 "-DFLAG=..." name2 "-D..." ...
set -e
cd "$(dirname "$0")/../katana.jl_b200/csrc"
make -s >/dev/null
mkdir -p ../../build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -ccbin g++ $flags -Xptxas -v -c -o ../../build/variants/k_$name.o ktn_kernels.cu 2> ../../build/variants/k_$name.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -ccbin g++ $flags -c -o ../../build/variants/c_$name.o ktn_comm.cu
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker -Bsymbolic -o ../../build/variants/libktn_$name.so ktn_api.o ../../build/variants/c_$name.o ../../build/variants/k_$name.o ktn_compile.o ktn_synth.o -lcudart -ldl
  echo "$name: $flags :: $(grep -A2 'ktn_family_kernelILi1' ../../build/variants/k_$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -B1 'ktn_family_kernelILi1' ../../build/variants/k_$name.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
