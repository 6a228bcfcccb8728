#!/bin/bash
# round 2, GPU call 29 (1 GPU): programmatic dependent launch between the kernels of the split stage -- GPU suite, A/B on the named configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call29
timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -2 ${O}_pytest.txt
export PYH_SPLIT=1
for cfg in explosion_multi dmr; do
  for nopdl in 0 1; do
    if [ $nopdl = 1 ]; then export PYH_NO_PDL=1; else unset PYH_NO_PDL; fi
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_nopdl${nopdl}.json 2> ${O}_${cfg}_nopdl${nopdl}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_nopdl${nopdl}.json").read().strip().splitlines()[-1])
print("$cfg no_pdl=$nopdl", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
