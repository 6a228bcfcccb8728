#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain-steps 0 --e2e-steps 1"
for spec in "A=1" "PYH_NO_PUSH_GHOST=1"; do
  env $spec $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'value %.4g ms/step %.3f stage_ms %.4f launches %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['gpu_launches']), {k:(round(v['value']/1e9,3), round(v['ms_per_step'],4)) for k,v in d.get('named_configs',{}).items()}, round(d.get('other_schemes',{}).get('hlll_primitive_rk2',{}).get('value',0)/1e9,3))"
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
