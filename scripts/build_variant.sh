#!/bin/bash
# usage: scripts/build_variant.sh <tag> <file.cu> [-DNAME=VALUE ...]
# A second build of libssgpu.so with ONE source compiled under extra defines -> statespacer_b200/libssgpu_<tag>.so
# (A/B runs of kernel variants on the GPU box: SSGPU_LIB=statespacer_b200/libssgpu_<tag>.so python bench.py ...)
set -e
cd "$(dirname "$0")/.."
tag=$1; src=$2; shift 2
C=statespacer_b200/csrc
mkdir -p $C/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC "$@" -Xptxas -v \
  -c $C/$src -o $C/build/variants/${src%.cu}_$tag.o 2>&1 | grep -A2 "${VARIANT_GREP:-nomatch}" || true
objs=$(ls $C/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o statespacer_b200/libssgpu_$tag.so $objs $C/build/variants/${src%.cu}_$tag.o -gencode arch=compute_100a,code=sm_100a
echo built statespacer_b200/libssgpu_$tag.so
