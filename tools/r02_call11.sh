#!/bin/bash
# round 2, GPU call 11 (4 GPUs): sharded parity after the tile-plan fix (all four sides remote, both splits active)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r02_call11_pytest.txt 2>&1
tail -30 gpurun_out/r02_call11_pytest.txt | cut -c1-400
