#!/bin/bash
# build an A/B variant of libribca_b200.so with extra nvcc flags.  usage: tools/build_variant.sh <name> <flags...>
# -> multiplexed_image_annotator_b200/build/libribca_<name>.so (load with RIBCA_LIB=...)
set -e
name=$1; shift
P=multiplexed_image_annotator_b200
mkdir -p $P/build/$name
for f in $P/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude "$@" -c $f -o $P/build/$name/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $P/build/libribca_$name.so $P/build/$name/*.o
echo $P/build/libribca_$name.so
