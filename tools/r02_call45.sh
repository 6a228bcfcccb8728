#!/bin/bash
# round 2, GPU call 45 (1 GPU): recon kernel prefetches its geometry into the L1 ahead of the PDL wait and the tile staging -- A/B against the previous build
# block table read ahead of the PDL wait in all three kernels -- A/B against the previous build on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call45
export PYH_SPLIT=1
for rep in 1 2; do
for cfg in explosion_multi dmr; do
  for lib in prev new; do
    if [ $lib = prev ]; then export PYH_LIB_PATH=$PWD/gpurun_variants/libpyh_prev.so; else unset PYH_LIB_PATH; fi
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_${lib}.json 2> ${O}_${cfg}_${lib}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_${lib}.json").read().strip().splitlines()[-1])
print("$cfg $lib", "value %.4g ms/step %.4f parity %s" % (d["value"], d["ms_per_step"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
done
