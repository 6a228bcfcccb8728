#!/bin/bash
# round 2, GPU call 44 (1 GPU): the library as rebuilt from clean objects at the end of the round (what ships): GPU suite + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call44_pytest.txt 2>&1
tail -n 2 gpurun_out/r02_call44_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()"
