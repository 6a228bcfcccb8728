#!/bin/bash
# round 2, GPU call 12 (8 GPUs): the 8-rank sharded parity run after the tile-plan fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "matches_oracle and 8" > gpurun_out/r02_call12_pytest.txt 2>&1
tail -30 gpurun_out/r02_call12_pytest.txt | cut -c1-300
