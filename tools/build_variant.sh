#!/bin/bash
# Kernel-tuning helper: build a variant of the library with only the bench instantiation
# (Roe + Venkatakrishnan + conservative, 1 quadrature point) and extra -D flags, for A/B runs on the GPU box:
#   tools/build_variant.sh NAME [-DPYH_MARCH_MINB=3 ...]   ->  gpurun_variants/libpyh_NAME.so
#   PYH_LIB_PATH=gpurun_variants/libpyh_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants
nvcc -shared -Xcompiler -fPIC -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -DPYH_ONLY_ROE_VENKAT_CONS "$@" \
  -I include -o gpurun_variants/libpyh_${name}.so pyhype_b200/csrc/pyh_api.cu pyhype_b200/csrc/pyh_march_nq1.cu \
  pyhype_b200/csrc/pyh_march_nq2.cu pyhype_b200/csrc/pyh_march_nq3.cu
cuobjdump -res-usage gpurun_variants/libpyh_${name}.so 2>/dev/null | grep -A1 "k_stage_marchILi[0-9]ELi[0-9]ELi[0-9]ELi1" | grep REG
