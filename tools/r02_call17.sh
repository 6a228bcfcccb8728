#!/bin/bash
# round 2, GPU call 17 (1 GPU): rows per strip chosen for an integer number of waves (592 resident thread blocks; 17 column strips x 8 blocks)
cd "$(dirname "$0")/.."
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1"
for spec in "A=1" "PYH_MARCH_TYS=61" "PYH_MARCH_TYS=69" "PYH_MARCH_TYS=79" "PYH_MARCH_TYS=98" "PYH_MARCH_TYS=121" "PYH_MARCH_TYS=52" "A=1"; do
  env $spec $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'value %.4g ms/step %.3f stage_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg']))"
done
