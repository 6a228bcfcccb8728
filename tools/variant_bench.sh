#!/bin/bash
# A/B kernel variants on the GPU box: each line "NAME ENV..." runs bench.py with PYH_LIB_PATH=gpurun_variants/libpyh_NAME.so
# usage (on the box): tools/variant_bench.sh "base" "base PYH_MARCH_TYS=128" "minb3" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  out=$(env "$@" PYH_LIB_PATH=gpurun_variants/libpyh_${name}.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-named --sustain-steps 0 2>&1 | tail -1)
  echo "$spec :: $(echo "$out" | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print("value %.4g ms/step %.3f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
except Exception as e: print("FAILED", e)')"
  # bit parity of the variant on explosion_multi at its named size, initial condition -> t_final (1604 steps, fingerprints of
  # the unmodified reference; the bench-only libraries hold exactly this scheme); set PYH_VARIANT_PARITY=0 to skip
  if [ "${PYH_VARIANT_PARITY:-1}" != 0 ]; then
    env "$@" PYH_LIB_PATH=$PWD/gpurun_variants/libpyh_${name}.so python -m pytest tests/test_named_configs.py -m gpu -k em -q 2>&1 | tail -1 | sed "s/^/    named-size parity: /"
  fi
done
