#!/bin/bash
# round 2, GPU call 2 (1 GPU): whole GPU suite on the new library (pyh_run role fix, C-layer comm), FP64 issue-model
# microbenchmarks, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_call2_pytest.txt 2>&1
tail -5 gpurun_out/r02_call2_pytest.txt
tools/fp64_mix > gpurun_out/r02_fp64_mix.txt 2>&1
tools/fp64_latency > gpurun_out/r02_fp64_latency.txt 2>&1
cat gpurun_out/r02_fp64_mix.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_call2_bench.json 2> gpurun_out/r02_call2_bench.err
tail -c 3000 gpurun_out/r02_call2_bench.json; tail -5 gpurun_out/r02_call2_bench.err
