#!/bin/bash
# round 2, GPU call 37 (1 GPU): warm per-kernel times of the final split stage at DMR's and explosion_multi's size (ncu --cache-control none)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call37
PYH_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 60 --csv --log-file ${O}_dmr_warm.csv python bench.py --config dmr --steps 30 > /dev/null 2>&1
PYH_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 130 --csv --log-file ${O}_em_warm.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
