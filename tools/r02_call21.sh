#!/bin/bash
# round 2, GPU call 21 (1 GPU): where the three-kernel stage spends its time -- launch list + full ncu capture of its kernels at
# explosion_multi's size and at 8 x 512^2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_call21
export PYH_SPLIT=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 80 --csv --log-file ${O}_em_launches.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_split -s 39 -c 6 -f -o ${O}_em_split python bench.py --config explosion_multi --steps 30 > ${O}_ncu_em.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_split -s 39 -c 3 -f -o ${O}_ws512_split python bench.py --block 512 --steps 4 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1 > ${O}_ncu_ws.log 2>&1
ls -la gpurun_out/*call21*
