#!/bin/bash
# Developer helper: build an A/B variant of the library with extra nvcc flags (select it with POMCP_B200_LIB).
#   scripts/build_variant.sh probe -DPOMCP_PROBE   ->  scripts/_ab_probe.so
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ -shared \
  -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function "$@" -o scripts/_ab_$name.so pomcp.jl_b200/csrc/abi.cu -ldl -lcudart \
  && echo scripts/_ab_$name.so
