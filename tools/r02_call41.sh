#!/bin/bash
# round 2, GPU call 33 (1 GPU): evidence for the final build -- GPU suite, default bench line, launch list of the bench command,
# compute-sanitizer memcheck + racecheck on the smoke case through the split stage (PDL + L2 window), smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call41
timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -n 3 ${O}_pytest.txt
python bench.py --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -n 3 ${O}_bench.err
python - <<PY
import json
d=json.loads(open("${O}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f stage_ms %.4f e2e %.4g sustained %.4g launches %d frac %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["sustained"]["value"], d["gpu_launches"], d["roofline"]["frac"]), d["config"]["setup_s"])
for k,v in d.get("named_configs",{}).items(): print("  ",k, v.get("value"), v.get("ms_per_step"), v.get("parity",{}).get("bit_identical_to_reference"), v.get("stage_path"), v.get("stage_path_tuning_ms"), v.get("error"))
print(d.get("other_schemes",{}).get("hlll_primitive_rk2",{}).get("value"))
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv $B > /dev/null 2>&1
PYH_SPLIT=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_memcheck_split1.txt 2>&1; echo "memcheck rc=$?" >> ${O}_memcheck_split1.txt
PYH_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_racecheck_split1.txt 2>&1; echo "racecheck rc=$?" >> ${O}_racecheck_split1.txt
tail -n 3 ${O}_memcheck_split1.txt; tail -n 3 ${O}_racecheck_split1.txt
python -c "import __graft_entry__ as g; g.smoke()"
