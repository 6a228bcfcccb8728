#!/bin/bash
# round 2, GPU call 24 (1 GPU): two-kernel stage (flux + update merged, 31 x 7 cells per 32 x 8 thread tile) -- GPU suite, forced
# A/B on the named configurations, warm per-kernel times, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call24
timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
for cfg in explosion_multi dmr; do
  for s in 0 1; do
    PYH_SPLIT=$s timeout 300 python bench.py --config $cfg > ${O}_${cfg}_split${s}.json 2> ${O}_${cfg}_split${s}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_split${s}.json").read().strip().splitlines()[-1])
print("$cfg split=$s", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
for b in 256 512; do
  for s in 0 1; do
    PYH_SPLIT=$s timeout 300 python bench.py --block $b --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 4 > ${O}_ws${b}_split${s}.json 2> ${O}_ws${b}_split${s}.err
    python - <<PY
import json
d=json.loads(open("${O}_ws${b}_split${s}.json").read().strip().splitlines()[-1])
print("ws 8 x $b^2 split=$s value %.4g ms/step %.4f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
PY
  done
done
PYH_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 90 --csv --log-file ${O}_em_split1_warm.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
PYH_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 40 --csv --log-file ${O}_dmr_split1_warm.csv python bench.py --config dmr --steps 30 > /dev/null 2>&1
