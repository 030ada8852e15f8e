#!/bin/bash
# Final evidence of round 2 on one B200 (run under gpurun): launch list of one sampling step with the final build, ncu --set
# full of the attention core (both variants) and pair_fc, the default bench line.  Outputs under gpurun_out/r02f/.
set -x
O=gpurun_out/r02f
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_step.csv \
    python bench.py --steps 1 --warmup 3 --profile-eager > $O/launches_step.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv 115 > $O/launches_step.md
PRD_STEP_HINT=1 ncu --set full --clock-control none --import-source on -k regex:"triattn_flash_g4|pair_transition_ws|pair_bias_kernel" -c 8 -f -o $O/core \
    python bench.py --steps 1 --warmup 3 --profile-eager > $O/ncu_core.log 2>&1
python tools/ncu_digest.py $O/core.ncu-rep > $O/core_digest.txt 2>&1
python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
python bench.py --workload train --steps 10 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
for w in config1 config2 config5; do python bench.py --workload $w --steps 30 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/train_breakdown.py > $O/train_breakdown.txt 2>&1
ls -la $O
