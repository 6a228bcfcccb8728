#!/bin/bash
# Full library (every instantiation) with extra -D flags, for running the whole GPU suite against a variant:
#   tools/build_full_variant.sh NAME [-DPYH_FOLD_POW2=1 ...]   ->  gpurun_variants/libpyh_full_NAME.so
#   PYH_LIB_PATH=$PWD/gpurun_variants/libpyh_full_NAME.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants build/var_$name
FLAGS="-Xcompiler -fPIC -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -I include"
pids=()
for u in pyh_api pyh_march_nq1 pyh_march_nq2 pyh_march_nq3; do
  nvcc -c $FLAGS "$@" -o build/var_$name/$u.o pyhype_b200/csrc/$u.cu & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -Wno-deprecated-gpu-targets -o gpurun_variants/libpyh_full_${name}.so build/var_$name/*.o
cuobjdump -res-usage gpurun_variants/libpyh_full_${name}.so 2>/dev/null | grep -A1 "k_stage_marchILi0ELi0ELi0ELi1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' '; echo
