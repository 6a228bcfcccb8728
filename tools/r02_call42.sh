#!/bin/bash
# round 2, GPU call 42 (1 GPU): what the measured stage path is on the weak-scaling mesh at smaller block sizes, final build
# (informational: bench.py now reports config.stage_path / stage_path_tuning_ms on its main line), plus the default line's new fields
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call42
for b in 128 256 512 700; do
    timeout 300 python bench.py --block $b --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 4 > ${O}_ws${b}.json 2> ${O}_ws${b}.err
    python - <<PY
import json
d=json.loads(open("${O}_ws${b}.json").read().strip().splitlines()[-1])
print("ws 8 x $b^2 value %.4g ms/step %.4f" % (d["value"], d["ms_per_step"]), d["config"]["stage_path"], d["config"]["stage_path_tuning_ms"])
PY
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 2 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default size', d['value'], d['config']['stage_path'], d['config']['stage_path_tuning_ms'])"
