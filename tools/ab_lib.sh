for v in base new base new; do
  if [ $v = base ]; then export PRD_LIB_PATH=$PWD/protein_redesign_b200/libprd_sm100_base.so PRD_ALLOW_STALE_LIB=1; else unset PRD_LIB_PATH PRD_ALLOW_STALE_LIB; fi
  python bench.py --steps 40 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-sample-parallel --no-sustained --no-ragged 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
