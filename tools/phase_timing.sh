#!/bin/bash
# Profiling aid: build with phase stamps (and optionally an experiment), run two bench steps, keep the stamp lines.
# usage: tools/phase_timing.sh <exp 0|1|2> <out log>
set -e
DMDQN_TC_TIMING=1 DMDQN_TC_EXP=$1 python -m dmdqn_b200.build --force > /dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^K3|^K4" | tail -${3:-27} > $2
