#!/bin/bash
# round 2, GPU call 10 (8 GPUs): sharded parity at 8 ranks, bench N = 8 (+ N = 1 on the same box), reference arm at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "8 or more_ranks" > gpurun_out/r02_call10_pytest.txt 2>&1
tail -5 gpurun_out/r02_call10_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 8 --sustain-steps 100 > gpurun_out/r02_call10_bench_n8.json 2> gpurun_out/r02_call10_bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_call10_bench_n8.json").read().strip().splitlines()[-1])
    print("N=8", "value %.4g ms/step %.3f stage_ms %.4f sustained %.4g e2e %.4g selfcheck %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d.get("sustained",{}).get("value",0), d["e2e"]["value"], d.get("selfcheck",{}).get("sharded_equals_single_gpu")), d["roofline"].get("kernel_ms_avg_per_rank"))
except Exception as e:
    print("N=8 FAILED", e); print(open("gpurun_out/r02_call10_bench_n8.err").read()[-2500:])
PY
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 100 --e2e-steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', 'value %.4g ms/step %.3f stage_ms %.4f sustained %.4g e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['sustained']['value'], d['e2e']['value']))"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29782 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02_call10_ref_n8.json 2> gpurun_out/r02_call10_ref_n8.err
tail -c 700 gpurun_out/r02_call10_ref_n8.json
