#!/bin/bash
# usage: build_variant.sh <name> [extra nvcc flags...]  -> gpurun_out/variants/libb2pt_<name>.so  (experiments only)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=variants
mkdir -p $out
C=raytracingtherestofyourlife_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fopenmp"
/usr/local/cuda/bin/nvcc $F "$@" -shared -o $out/libb2pt_$name.so $C/b2pt_kernels.cu $C/b2pt_api.cu $C/b2pt_lbvh.cu -x cu $C/b2pt_scene.cpp
echo built $out/libb2pt_$name.so
