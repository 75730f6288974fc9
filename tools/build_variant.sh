#!/bin/bash
# Experiment builds of the library: tools/build_variant.sh <name> <source.cu> <extra nvcc flags...>
# recompiles ONE translation unit with extra flags and links it with the production objects into
# zkstark_b200/variants/libzk_<name>.so (git-ignored; travels to the GPU box).  Use with ZK_LIB=... tools/*.py.
set -e
cd "$(dirname "$0")/../zkstark_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p ../variants
obj=/tmp/zkvar_${name}_$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC "$@" -c $src -o $obj
objs=""
for o in tree.o ntt.o capi.o sharded.o benchsupport.o trace.o host_sha256.o host_trace.o verify.o; do
  if [ "$o" = "$(basename $src .cu).o" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libzk_${name}.so $objs -lcudart -ldl
echo built ../variants/libzk_${name}.so
