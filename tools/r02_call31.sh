#!/bin/bash
# round 2, GPU call 31 (1 GPU): persisting-L2 access-policy window over the split stage's scratch planes -- A/B (PYH_NO_L2_PERSIST=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call31
export PYH_SPLIT=1
for cfg in explosion_multi dmr; do
  for nop in 0 1; do
    if [ $nop = 1 ]; then export PYH_NO_L2_PERSIST=1; else unset PYH_NO_L2_PERSIST; fi
    timeout 300 python bench.py --config $cfg > ${O}_${cfg}_nopersist${nop}.json 2> ${O}_${cfg}_nopersist${nop}.err
    python - <<PY
import json
d=json.loads(open("${O}_${cfg}_nopersist${nop}.json").read().strip().splitlines()[-1])
print("$cfg no_persist=$nop", d.get("stage_path"), "value %.4g ms/step %.4f launches %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parity"]["bit_identical_to_reference"]))
PY
  done
done
unset PYH_NO_L2_PERSIST
for b in 256 512; do
    timeout 300 python bench.py --block $b --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 4 > ${O}_ws${b}_split1.json 2> ${O}_ws${b}_split1.err
    python - <<PY
import json
d=json.loads(open("${O}_ws${b}_split1.json").read().strip().splitlines()[-1])
print("ws 8 x $b^2 split=1 persist value %.4g ms/step %.4f stage_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"]))
PY
done
python -c "
import ctypes
rt = ctypes.CDLL('libcudart.so.12')
v = ctypes.c_int()
for a, n in ((108, 'MaxPersistingL2CacheSize'), (109, 'MaxAccessPolicyWindowSize'), (38, 'L2CacheSize')):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); print(n, v.value)
"
