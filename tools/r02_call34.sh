#!/bin/bash
# round 2, GPU call 34 (1 GPU): occupancy target of the split stage's HLLL flux kernel (launch bounds 128 x {1, 5, 6}: 122 / 96 / 80 registers) on DMR
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call34
export PYH_SPLIT=1
for m in 1 5 6 1 5 6; do
    PYH_SPLIT_FLUX_MINB=$m timeout 300 python bench.py --config dmr > ${O}_dmr_minb${m}.json 2> ${O}_dmr_minb${m}.err
    python - <<PY
import json
d=json.loads(open("${O}_dmr_minb${m}.json").read().strip().splitlines()[-1])
print("dmr flux min blocks=$m", "value %.4g ms/step %.4f parity %s" % (d["value"], d["ms_per_step"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
done
