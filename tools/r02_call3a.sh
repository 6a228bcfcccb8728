#!/bin/bash
# round 2, GPU call 3a (1 GPU): new GPU tests (device fills, jet_hlle abort, tile interface), combined kernel variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call3a_pytest.txt 2>&1
tail -5 gpurun_out/r02_call3a_pytest.txt
export PYH_VARIANT_PARITY=1
tools/variant_bench.sh base "base PYH_NO_FUSED_DT=1" c1 c2 c3 c4 c5 base > gpurun_out/r02_variants_combined.txt 2>&1
cat gpurun_out/r02_variants_combined.txt
