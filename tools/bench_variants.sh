#!/bin/bash
# usage: tools/bench_variants.sh  -> runs a short bench for every lib in pyhype_b200/lib/variants
for f in pyhype_b200/lib/variants/*.so; do
  for nt in ${NTS:-auto}; do
    if [ "$nt" != "auto" ]; then export PYH_MARCH_NT=$nt; else unset PYH_MARCH_NT; fi
    PYH_LIB_PATH=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$f', 'NT=$nt', '%.3e cell-stage/s' % d['value'], 'stage kernel %.3f ms' % d['roofline']['kernel_ms_avg'])"
  done
done
