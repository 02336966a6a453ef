#!/bin/bash
# Builds a variant of libissl_cuda.so with extra nvcc flags into crackling_b200/lib_ab/<name>/ for A/B timing:
#   tools/ab_build.sh l256 -DISSL_BLOCK_POLICY=3
#   ISSL_CUDA_LIB=crackling_b200/lib_ab/l256/libissl_cuda.so python tools/score_once.py
# (the directory is git-ignored; the variants travel to the GPU box with the snapshot)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=crackling_b200/lib_ab/$name
mkdir -p "$out"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fopenmp,-Wall -Iinclude -Icrackling_b200/csrc --expt-relaxed-constexpr"
$NVCC $FLAGS "$@" -c -o "$out/issl_device.o" crackling_b200/csrc/issl_device.cu
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libissl_cuda.so" crackling_b200/lib/issl_host.o crackling_b200/lib/issl_multi.o \
      "$out/issl_device.o" crackling_b200/lib/issl_sites.o -Xcompiler -fopenmp -lgomp -lpthread
rm -f "$out/issl_device.o"
echo "built $out/libissl_cuda.so ($*)"
