#!/bin/bash
# round 2, GPU call 47 (1 GPU): full ncu capture of the final split-stage kernels at explosion_multi's size (one stage)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYH_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:k_split -s 39 -c 3 -f -o gpurun_out/r02_call47_em_split python bench.py --config explosion_multi --steps 30 > gpurun_out/r02_call47_ncu.log 2>&1
ls -la gpurun_out/r02_call47_em_split.ncu-rep
