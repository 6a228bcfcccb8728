#!/bin/bash
# round 2, GPU call 26 (2 GPUs): multi-rank regression after the split stage path (BlkDev grew a pointer; contexts with remote
# neighbours stay on the fused kernel): multi-rank GPU tests, 2-GPU bench line with its self-check, reference arm under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call26
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -3 ${O}_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29832 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 4 --sustain-steps 100 > ${O}_bench_n2.json 2> ${O}_bench_n2.err
tail -2 ${O}_bench_n2.err
python - <<PY
import json
d=json.loads(open("${O}_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.4g ms/step %.3f selfcheck %s e2e %.4g" % (d["value"], d["ms_per_step"], d.get("selfcheck",{}).get("sharded_equals_single_gpu"), d["e2e"]["value"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29833 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > ${O}_ref_n2.json 2> ${O}_ref_n2.err
tail -c 600 ${O}_ref_n2.json
