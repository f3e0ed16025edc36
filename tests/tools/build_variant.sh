#!/bin/bash
# Builds an A/B variant of libgwb200.so with extra nvcc flags: tests/tools/build_variant.sh <name> <flags...>
# -> build/variants/libgwb200_<name>.so (select with GWB200_LIB_PATH)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
O=build/variants/obj_$name
mkdir -p $O
F="-std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -Xcompiler -O2"
pids=()
for s in genomeworks_b200/csrc/*.cu genomeworks_b200/csrc/*.cpp; do
  x=""; [[ $s == *.cpp ]] && x="-x cu"
  nvcc $F "$@" $x -c $s -o $O/$(basename $s).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libgwb200_$name.so $O/*.o -lcudart -lpthread
echo built build/variants/libgwb200_$name.so
