#!/bin/bash
# round 2, GPU call 1: whole GPU suite with the shipped build (first hardware run of tests/test_named_configs.py), then the A/B of
# every CPU-validated kernel variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_call1_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_call1_pytest.txt 2>&1
tail -3 gpurun_out/r02_call1_pytest.txt
export PYH_VARIANT_PARITY=0
tools/variant_bench.sh base tight b2x2 b2x4 b1x2_b2x2 dearly1 dearly2 dearly2_b2x2 geomfirst geomfirst_dearly1 hcert hcert_minmax ldg1 ldg2 ldg2_b2x2_dearly2 uniform "base PYH_MARCH_NT=64" "base PYH_MARCH_TYS=128" "base PYH_MARCH_TYS=32" base > gpurun_out/r02_variants.txt 2>&1
cat gpurun_out/r02_variants.txt
