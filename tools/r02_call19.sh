#!/bin/bash
# round 2, GPU call 19 (1 GPU): GPU suite after the NaN-propagating limiter minimum (shockbox abort), default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call19_pytest.txt 2>&1
tail -4 gpurun_out/r02_call19_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_call19_bench.json 2> gpurun_out/r02_call19_bench.err
tail -3 gpurun_out/r02_call19_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_call19_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f stage_ms %.4f e2e %.4g sustained %.4g launches %d frac %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["sustained"]["value"], d["gpu_launches"], d["roofline"]["frac"]), d["config"]["setup_s"])
for k,v in d.get("named_configs",{}).items(): print("  ",k, v.get("value"), v.get("ms_per_step"), v.get("parity",{}).get("bit_identical_to_reference"), v.get("stage_kernel_shape"), v.get("error"))
print(d.get("other_schemes",{}).get("hlll_primitive_rk2",{}).get("value"))
PY
