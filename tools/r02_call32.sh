#!/bin/bash
# round 2, GPU call 32 (1 GPU): persisting-L2 window over the whole scratch when it fits the carve-out, else over the flux planes alone
# (PYH_NO_L2_PERSIST_FX=1: else nothing) -- A/B on DMR and 8 x 256^2; explosion_multi as the control
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call32
export PYH_SPLIT=1
for cfg in explosion_multi dmr; do
  for nofx in 0 1; do
    if [ $nofx = 1 ]; then export PYH_NO_L2_PERSIST_FX=1; else unset PYH_NO_L2_PERSIST_FX; fi
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_nofx${nofx}.json 2> ${O}_${cfg}_nofx${nofx}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_nofx${nofx}.json").read().strip().splitlines()[-1])
print("$cfg no_fx_window=$nofx", d.get("stage_path"), "value %.4g ms/step %.4f parity %s" % (d["value"], d["ms_per_step"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
for nofx in 0 1; do
    if [ $nofx = 1 ]; then export PYH_NO_L2_PERSIST_FX=1; else unset PYH_NO_L2_PERSIST_FX; fi
    timeout 300 python bench.py --block 256 --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 4 > ${O}_ws256_nofx${nofx}.json 2> ${O}_ws256_nofx${nofx}.err
    python - <<PY
import json
d=json.loads(open("${O}_ws256_nofx${nofx}.json").read().strip().splitlines()[-1])
print("ws 8 x 256^2 split=1 no_fx_window=$nofx value %.4g ms/step %.4f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
PY
done
