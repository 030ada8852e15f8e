#!/bin/bash
# A/B of two BUILDS of the library on one box (fresh boxes differ by +-2 %, up to 7 % on the training workload):
#   1. build the baseline sources (e.g. `git stash; csrc/build.sh; cp protein_redesign_b200/libprd_sm100.so
#      protein_redesign_b200/libprd_sm100_base.so; git stash pop; csrc/build.sh`): the copy travels with the gpurun snapshot;
#   2. gpurun -- 'bash tools/ab_lib.sh [extra bench.py args, e.g. --workload train]'
# The baseline is loaded through PRD_LIB_PATH with the source-hash check switched off (PRD_ALLOW_STALE_LIB=1).
for v in base new base new; do
  if [ $v = base ]; then export PRD_LIB_PATH=$PWD/protein_redesign_b200/libprd_sm100_base.so PRD_ALLOW_STALE_LIB=1; else unset PRD_LIB_PATH PRD_ALLOW_STALE_LIB; fi
  python bench.py --steps 40 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-sample-parallel --no-sustained --no-ragged "$@" 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
