#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG ...]   -> bpp_b200/variants/libbppgpu_<name>.so (select with BPPGPU_LIB=...)
name=$1; shift
mkdir -p bpp_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --cudart static "$@" \
  -o bpp_b200/variants/libbppgpu_$name.so bpp_b200/csrc/engine.cu
