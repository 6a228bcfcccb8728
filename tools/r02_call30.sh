#!/bin/bash
# round 2, GPU call 30 (8 GPUs): final build on the full box -- 8-rank bench line with its sharded == single-GPU self-check, N = 1 on
# GPU 0 of the same box for the efficiency figure, multi-rank GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29838 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 4 --sustain-steps 100 > ${O}_bench_n8.json 2> ${O}_bench_n8.err
tail -2 ${O}_bench_n8.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1 > ${O}_bench_n1.json 2> ${O}_bench_n1.err
python - <<PY
import json
d8=json.loads(open("${O}_bench_n8.json").read().strip().splitlines()[-1])
d1=json.loads(open("${O}_bench_n1.json").read().strip().splitlines()[-1])
print("N=8 value %.4g ms/step %.3f selfcheck %s e2e %.4g per-rank kernel ms %s" % (d8["value"], d8["ms_per_step"], d8.get("selfcheck",{}).get("sharded_equals_single_gpu"), d8["e2e"]["value"], d8["roofline"].get("kernel_ms_avg_per_rank")))
print("N=1 value %.4g ms/step %.3f  -> weak-scaling efficiency at 8: %.4f" % (d1["value"], d1["ms_per_step"], d8["value"]/(8*d1["value"])))
PY
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > ${O}_pytest.txt 2>&1
tail -3 ${O}_pytest.txt
