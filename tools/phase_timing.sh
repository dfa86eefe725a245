#!/bin/bash
# Profiling aid: build the phase-stamp variant into libdmdqn_b200_timing.so (the product library is untouched), run two
# bench steps against it, keep the stamp lines.   usage: tools/phase_timing.sh <out log> [lines]
set -e
python -m dmdqn_b200.build --timing > /dev/null
DMDQN_PROFILING_LIB=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --blocks none 2>&1 | grep -E "^K3|^K4" | tail -${2:-40} > $1
