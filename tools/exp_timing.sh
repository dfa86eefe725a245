#!/bin/bash
# Profiling aid: build with an experiment switch (-DTC_EXP=n, results are WRONG by design) and print the per-kernel times.
# usage: tools/exp_timing.sh <exp> [<exp> ...]
for e in "$@"; do
  DMDQN_TC_TIMING= DMDQN_TC_EXP_ONLY=$e python -m dmdqn_b200.build --force > /dev/null
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print('exp $e', round(j['value']), {k:round(v['ms']*1000,1) for k,v in j['kernels'].items()})"
done
