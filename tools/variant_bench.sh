#!/bin/bash
# A/B kernel variants on the GPU box: each line "NAME ENV..." runs bench.py with PYH_LIB_PATH=gpurun_variants/libpyh_NAME.so
# usage (on the box): tools/variant_bench.sh "base" "base PYH_MARCH_TYS=128" "minb3" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  out=$(env "$@" PYH_LIB_PATH=gpurun_variants/libpyh_${name}.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1)
  echo "$spec :: $(echo "$out" | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print("value %.4g ms/step %.3f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
except Exception as e: print("FAILED", e)')"
done
