#!/bin/bash
# Build the bench-only (Roe + Venkatakrishnan + conservative, 1 quadrature point) libraries of every opt-in kernel variant
# that is validated on the CPU (tests/test_host_twin.py, tests/test_kernel_twin.py) but not yet measured, and print the
# gpurun line that A/Bs them.  A variant that wins must then pass the whole GPU suite as a FULL library
# (tools/build_full_variant.sh NAME flags; PYH_LIB_PATH=... python -m pytest tests -m gpu) before it becomes the default.
set -e
cd "$(dirname "$0")/.."
tools/build_variant.sh base
tools/build_variant.sh tight -DPYH_LEAN_CHECKS=1 -DPYH_COLD_HOOKS=1
tools/build_variant.sh b2x2 -DPYH_UNROLL_B2=2
tools/build_variant.sh b2x4 -DPYH_UNROLL_B2=4
tools/build_variant.sh b1x2_b2x2 -DPYH_UNROLL_B1=2 -DPYH_UNROLL_B2=2
tools/build_variant.sh dearly1 -DPYH_D_EARLY=1
tools/build_variant.sh dearly2 -DPYH_D_EARLY=2
tools/build_variant.sh dearly2_b2x2 -DPYH_D_EARLY=2 -DPYH_UNROLL_B2=2
tools/build_variant.sh geomfirst -DPYH_B_GEOM_FIRST=1
tools/build_variant.sh geomfirst_dearly1 -DPYH_B_GEOM_FIRST=1 -DPYH_D_EARLY=1
tools/build_variant.sh hcert -DPYH_HARTEN_CERT=1
tools/build_variant.sh hcert_minmax -DPYH_HARTEN_CERT=1 -DPYH_MINMAX_NET=1
tools/build_variant.sh uniform -DPYH_UNIFORM_SHORTCUT=1
tools/build_variant.sh uniform_tight -DPYH_UNIFORM_SHORTCUT=1 -DPYH_LEAN_CHECKS=1 -DPYH_COLD_HOOKS=1
echo
echo "gpurun --timeout 600 -- 'tools/variant_bench.sh base tight b2x2 b2x4 b1x2_b2x2 dearly1 dearly2 dearly2_b2x2 geomfirst geomfirst_dearly1 hcert hcert_minmax uniform uniform_tight \"base PYH_MARCH_NT=64\" \"geomfirst PYH_MARCH_NT=64\" base > gpurun_out/variants.txt 2>&1; \\"
echo "  for v in base uniform; do PYH_LIB_PATH=\$PWD/gpurun_variants/libpyh_\$v.so python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --ic smooth | tail -1 >> gpurun_out/variants_smooth.jsonl; done'"
