#!/bin/bash
# round 2, GPU call 43 (2 GPUs): last multi-rank sanity check of the final build (BlkDev grew once more after the 8-GPU run):
# 2-rank bench line with its sharded == single-GPU self-check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call43
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29842 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 4 --sustain-steps 0 > ${O}_bench_n2.json 2> ${O}_bench_n2.err
python - <<PY
import json
d=json.loads(open("${O}_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.4g ms/step %.3f selfcheck %s e2e %.4g stage_path %s" % (d["value"], d["ms_per_step"], d.get("selfcheck",{}).get("sharded_equals_single_gpu"), d["e2e"]["value"], d["config"].get("stage_path")))
PY
