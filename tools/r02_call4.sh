#!/bin/bash
# round 2, GPU call 4 (1 GPU): new defaults -- whole GPU suite, bench line, launch list of the EM line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call4_pytest.txt 2>&1
tail -5 gpurun_out/r02_call4_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_call4_bench.json 2> gpurun_out/r02_call4_bench.err
tail -c 1500 gpurun_out/r02_call4_bench.json; tail -3 gpurun_out/r02_call4_bench.err
PYH_NO_FUSED_DT=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1 > gpurun_out/r02_call4_bench_nofuse.json 2>/dev/null
for v in ilp4_minb2 ilp4_minb3 ilp2_minb3; do
  echo "== $v"; PYH_LIB_PATH=$PWD/gpurun_variants/libpyh_$v.so python bench.py --config explosion_multi 2>&1 | tail -1 | cut -c1-400
done
echo "== shipped"; python bench.py --config explosion_multi 2>&1 | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r02_call4_em_launches.csv python bench.py --config explosion_multi --steps 100 > /dev/null 2>&1
