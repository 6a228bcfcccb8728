#!/bin/bash
# round 2, GPU call 22 (1 GPU): stage path chosen by measurement at the first pyh_run -- GPU suite (default = tuned, plus every case
# forced through each path), default bench line with the named configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call22
timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
python bench.py --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -3 ${O}_bench.err
python - <<PY
import json
d=json.loads(open("${O}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f stage_ms %.4f e2e %.4g sustained %.4g launches %d frac %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["sustained"]["value"], d["gpu_launches"], d["roofline"]["frac"]), d["config"]["setup_s"])
for k,v in d.get("named_configs",{}).items(): print("  ",k, v.get("value"), v.get("ms_per_step"), v.get("parity",{}).get("bit_identical_to_reference"), v.get("stage_path"), v.get("stage_path_tuning_ms"), v.get("error"))
print(d.get("other_schemes",{}).get("hlll_primitive_rk2",{}).get("value"))
PY
