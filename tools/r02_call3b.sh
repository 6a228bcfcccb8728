#!/bin/bash
# round 2, GPU call 3b (2 GPUs): sharded parity through the C-layer NCCL transport, then both bench arms at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r02_call3b_pytest.txt 2>&1
tail -15 gpurun_out/r02_call3b_pytest.txt
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_call3b_bench_n2.json 2> gpurun_out/r02_call3b_bench_n2.err
tail -c 2500 gpurun_out/r02_call3b_bench_n2.json; tail -5 gpurun_out/r02_call3b_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_call3b_ref_n2.json 2> gpurun_out/r02_call3b_ref_n2.err
tail -c 600 gpurun_out/r02_call3b_ref_n2.json
