#!/bin/bash
# A/B of an environment switch on the headline bench: tools/ab_bench.sh VAR valueA valueB [extra bench args]
var="$1"; a="$2"; b="$3"; shift 3
for v in "$a" "$b" "$a" "$b"; do
  env "$var=$v" python bench.py --steps 40 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-sample-parallel --no-sustained --no-ragged "$@" 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', 'ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
