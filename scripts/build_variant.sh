#!/bin/bash
# Build an experimental copy of the coder library for A/B runs on the GPU box (selected with NS_CODER_LIB):
#   scripts/build_variant.sh NAME [extra nvcc flags...]   ->   gpurun_bin/libns_NAME.so   (git-ignored, travels with gpurun)
name=$1; shift
cd "$(dirname "$0")/.." && mkdir -p gpurun_bin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o gpurun_bin/libns_$name.so neuralsteganography_b200/csrc/ns_coder.cu neuralsteganography_b200/csrc/ns_codecs.cu 2>&1 | grep -E "error|spill" | head -20
ls -la gpurun_bin/libns_$name.so
