#!/bin/bash
# round 2, GPU call 16 (4 GPUs): sharded parity + bench with the push-model ghost refresh (strips written straight into the send buffer)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r02_call16_pytest.txt 2>&1
tail -6 gpurun_out/r02_call16_pytest.txt | cut -c1-300
for n in 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2983$n bench.py --gpus $n --steps 20 --warmup 3 --e2e-steps 4 --sustain-steps 100 > gpurun_out/r02_call16_bench_n$n.json 2> gpurun_out/r02_call16_bench_n$n.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_call16_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", "value %.4g ms/step %.3f stage_ms %.4f sustained %.4g e2e %.4g selfcheck %s launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d.get("sustained",{}).get("value",0), d["e2e"]["value"], d.get("selfcheck",{}).get("sharded_equals_single_gpu"), d["gpu_launches"]))
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/r02_call16_bench_n$n.err").read()[-1500:])
PY
done
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 100 --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', 'value %.4g ms/step %.3f stage_ms %.4f sustained %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['sustained']['value']))"
