#!/bin/bash
# round 2, GPU call 15 (1 GPU): ncu full capture of the stage kernel at explosion_multi's own size (why 42 us for 180 k cells)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 40 -c 2 -f -o gpurun_out/r02_call15_em_stage python bench.py --config explosion_multi --steps 30 > gpurun_out/r02_call15_ncu.log 2>&1
ls -la gpurun_out/r02_call15_em_stage.ncu-rep
