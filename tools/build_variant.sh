#!/bin/bash
# Development aid: build libb200math_<tag>.so with extra -D flags applied to the BLS12-381 kernels only
# (the other objects come from the regular `make lib` build).   usage: tools/build_variant.sh <tag> [-DFLAG ...]
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=mathlib_b200/csrc/build/var_$tag
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" \
     -c mathlib_b200/csrc/kernels_bls381.cu -o $out/kernels_bls381.o 2> $out/ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o mathlib_b200/libb200math_$tag.so mathlib_b200/csrc/build/abi.o \
     mathlib_b200/csrc/build/kernels_bn254.o $out/kernels_bls381.o mathlib_b200/csrc/build/kernels_bls377.o -lcudart
grep -A2 "vm_pairing_kernelINS_6BLS381ELi2" $out/ptxas.log | grep -E "Used|stack" | head -4
