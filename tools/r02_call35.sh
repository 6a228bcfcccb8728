#!/bin/bash
# round 2, GPU call 35 (1 GPU): occupancy targets of the split stage's kernels -- flux 128 x {6, 8} (80 / 64 registers), recon 256 x {3, 4} (80 / 64)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call35
export PYH_SPLIT=1
for cfg in explosion_multi dmr; do
  for spec in "6 3" "8 3" "6 4" "8 4"; do
    set -- $spec
    PYH_SPLIT_FLUX_MINB=$1 PYH_SPLIT_RECON_MINB=$2 timeout 300 python bench.py --config $cfg > ${O}_${cfg}_f$1_r$2.json 2> ${O}_${cfg}_f$1_r$2.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_f$1_r$2.json").read().strip().splitlines()[-1])
print("$cfg flux minb=$1 recon minb=$2", "value %.4g ms/step %.4f parity %s" % (d["value"], d["ms_per_step"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
