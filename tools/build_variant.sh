#!/bin/bash
# usage: tools_build_variant.sh <name> <extra nvcc flags...>  -> sc-a-loam_b200/csrc/variants/libs2m_<name>.so
set -e
cd "$(dirname "$0")/../sc-a-loam_b200/csrc"
name=$1; shift
mkdir -p variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O3"
nvcc $F "$@" -c s2m_kernels.cu -o variants/k_$name.o
nvcc $F "$@" -c s2m_api.cu -o variants/a_$name.o
nvcc $F "$@" -c s2m_fx.cu -o variants/f_$name.o
nvcc $F -shared -o variants/libs2m_$name.so variants/k_$name.o variants/a_$name.o variants/f_$name.o -lcudart
echo built variants/libs2m_$name.so
