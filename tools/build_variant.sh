#!/bin/bash
# Build an experimental variant of libaad_b200.so: tools/build_variant.sh NAME "-DFLAG=1 ..."
# -> aad_b200/exp/libaad_NAME.so (git-ignored; run with AAD_B200_LIBRARY=aad_b200/exp/libaad_NAME.so)
set -e
cd "$(dirname "$0")/../aad_b200/csrc"
name=$1; shift
mkdir -p ../exp build
nvcc $* -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I. \
  -c aad_kernels.cu -o ../exp/aad_kernels_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../exp/libaad_$name.so ../exp/aad_kernels_$name.o \
  build/aad_gpu.o build/aad_encoder.o build/aad_decoder.o build/aad_wav.o -lpthread -lm
rm -f ../exp/aad_kernels_$name.o
echo built aad_b200/exp/libaad_$name.so
