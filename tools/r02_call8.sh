#!/bin/bash
# round 2, GPU call 8 (1 GPU): cost of the edge / interior split in isolation (no exchange), edge strip height
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1"
for spec in "A=1" "PYH_FORCE_EDGE_SPLIT=1" "PYH_FORCE_EDGE_SPLIT=1 PYH_EDGE_ROWS=2" "PYH_FORCE_EDGE_SPLIT=1 PYH_EDGE_ROWS=8" "PYH_FORCE_EDGE_SPLIT=1 PYH_EDGE_ROWS=16" "A=1"; do
  env $spec $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'value %.4g ms/step %.3f stage_ms %.4f launches %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['gpu_launches']))"
done
