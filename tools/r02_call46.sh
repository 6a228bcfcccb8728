#!/bin/bash
# round 2, GPU call 46 (1 GPU): bench.py --config lines of the two named configurations on the final build (default = measured stage path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in explosion_multi dmr; do
    timeout 300 python bench.py --config $cfg > gpurun_out/r02_call46_${cfg}.json 2> gpurun_out/r02_call46_${cfg}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r02_call46_${cfg}.json").read().strip().splitlines()[-1])
print("$cfg", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s e2e %.4g frac %.4f" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
done
