#!/bin/bash
# round 2, GPU call 28 (1 GPU): stage-path tuning timed as CUDA-graph launches -- what it picks for the named configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call28
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_streaming.py tests/test_gpu_fill_box.py tests/test_gpu_golden.py -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -2 ${O}_pytest.txt
for cfg in explosion_multi dmr; do
  for rep in 1 2; do
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_${rep}.json 2> ${O}_${cfg}_${rep}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_${rep}.json").read().strip().splitlines()[-1])
print("$cfg run $rep", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
python -c "import __graft_entry__ as g; g.smoke()"
