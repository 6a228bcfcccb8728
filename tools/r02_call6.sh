#!/bin/bash
# round 2, GPU call 6 (1 GPU): final-build evidence -- GPU suite, bench line, launch list, full ncu capture of 4 consecutive
# stage launches (= one RK4 step), compute-sanitizer memcheck + racecheck on the smoke case
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call6_pytest.txt 2>&1
tail -4 gpurun_out/r02_call6_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_call6_bench.json 2> gpurun_out/r02_call6_bench.err
tail -c 600 gpurun_out/r02_call6_bench.json; tail -3 gpurun_out/r02_call6_bench.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named --sustain-steps 0 --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_call6_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 8 -c 4 -f -o gpurun_out/r02_call6_stage $B > gpurun_out/r02_call6_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call6_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_call6_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call6_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_call6_racecheck.txt
tail -4 gpurun_out/r02_call6_memcheck.txt gpurun_out/r02_call6_racecheck.txt
