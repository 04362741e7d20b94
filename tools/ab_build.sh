#!/bin/bash
# Developer A/B builds: tools/ab_build.sh <name> <file.cu> [-DFLAG ...] -> csrc/build/ab/<name>.so (the other objects come from the regular build).
# Run with RDM_B200_LIB=retrieval-augmented-diffusion-models_b200/csrc/build/ab/<name>.so
set -e
cd "$(dirname "$0")/../retrieval-augmented-diffusion-models_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c $src -o build/ab/$name.o
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/$name.so $objs build/ab/$name.o -lcuda
echo built build/ab/$name.so
