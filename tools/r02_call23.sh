#!/bin/bash
# round 2, GPU call 23 (1 GPU): warm-cache per-kernel times at explosion_multi's size (ncu keeps the caches: --cache-control none),
# both stage paths; the step time minus their sum is launch gaps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call23
for s in 0 1; do
  PYH_SPLIT=$s ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 130 --csv --log-file ${O}_em_split${s}_warm.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
done
PYH_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 100 -c 60 --csv --log-file ${O}_dmr_split1_warm.csv python bench.py --config dmr --steps 30 > /dev/null 2>&1
PYH_SPLIT=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 30 --csv --log-file ${O}_dmr_split0_warm.csv python bench.py --config dmr --steps 30 > /dev/null 2>&1
ls -la ${O}*
