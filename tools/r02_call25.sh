#!/bin/bash
# round 2, GPU call 25 (1 GPU): three-kernel stage with the reworked update kernel (one wave of 256-thread blocks, loads first, one
# atomic per block) and the recon kernel's geometry loads ahead of the barrier (A/B: PYH_SPLIT_LATE_GEOM=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call25
timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
export PYH_SPLIT=1
for cfg in explosion_multi dmr; do
  for late in 0 1; do
    if [ $late = 1 ]; then export PYH_SPLIT_LATE_GEOM=1; else unset PYH_SPLIT_LATE_GEOM; fi
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_late${late}.json 2> ${O}_${cfg}_late${late}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_late${late}.json").read().strip().splitlines()[-1])
print("$cfg late_geom=$late", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
unset PYH_SPLIT_LATE_GEOM
for b in 256 512; do
    timeout 300 python bench.py --block $b --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 4 > ${O}_ws${b}_split1.json 2> ${O}_ws${b}_split1.err
    python - <<PY
import json
d=json.loads(open("${O}_ws${b}_split1.json").read().strip().splitlines()[-1])
print("ws 8 x $b^2 split=1 value %.4g ms/step %.4f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 130 --csv --log-file ${O}_em_split1_warm.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 50 --csv --log-file ${O}_dmr_split1_warm.csv python bench.py --config dmr --steps 30 > /dev/null 2>&1
