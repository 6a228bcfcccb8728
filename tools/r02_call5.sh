#!/bin/bash
# round 2, GPU call 5 (2 GPUs): edge-first overlapped exchange vs blocking exchange, sharded parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r02_call5_pytest.txt 2>&1
tail -8 gpurun_out/r02_call5_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 2 --sustain-steps 100 > gpurun_out/r02_call5_bench_$name.json 2> gpurun_out/r02_call5_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_call5_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.4g ms/step %.3f stage_ms %.4f sustained %.4g selfcheck %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d.get("sustained",{}).get("value",0), d.get("selfcheck",{}).get("sharded_equals_single_gpu")))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r02_call5_bench_$name.err").read()[-1500:])
PY
}
run overlap A=1
run blocking PYH_NO_HALO_OVERLAP=1
run overlap2 A=1
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 100 --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single', 'value %.4g ms/step %.3f stage_ms %.4f sustained %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['sustained']['value']))"
