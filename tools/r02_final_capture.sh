#!/bin/bash
# Final evidence of round 2 on one B200 (run under gpurun).  Outputs under gpurun_out/r02f/:
#   launch list of one sampling step (ncu duration pass), ncu --set full of the attention core / pair_fc / the GEMM with the
#   new epilogue / the backward attention and dW kernels, the bench lines of every workload, the training-step breakdowns,
#   the GEMM / dW shape table.
set -x
O=gpurun_out/r02f
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_step.csv \
    python bench.py --steps 1 --warmup 3 --profile-eager > $O/launches_step.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv 115 > $O/launches_step.md
PRD_STEP_HINT=1 ncu --set full --clock-control none --import-source on -k regex:"triattn_flash_g4|pair_transition_ws|pair_bias_kernel|trimul_out|trimul_in_t|triattn_out|triattn_proj|coord_head|pair_embed_lut|outer_linear" -c 14 -f -o $O/core \
    python bench.py --steps 1 --warmup 3 --profile-eager > $O/ncu_core.log 2>&1
python tools/ncu_digest.py $O/core.ncu-rep > $O/core_digest.txt 2>&1
if [ "${CAPTURE_BWD_NCU:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:"attn_tc|bw_dw_tc" -c 14 -f -o $O/bwd \
    python tools/train_breakdown.py > $O/ncu_bwd.log 2>&1
python tools/ncu_digest.py $O/bwd.ncu-rep > $O/bwd_digest.txt 2>&1
fi
python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
for m in graph manual autograd; do python bench.py --workload train --train-mode $m --steps 20 --warmup 3 > $O/bench_train_$m.json 2> $O/bench_train_$m.err; done
python bench.py --workload train --train-sizes 2 --steps 20 --warmup 3 > $O/bench_train_n374.json 2> $O/bench_train_n374.err
for w in config1 config2 config5; do python bench.py --workload $w --steps 30 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/train_breakdown.py > $O/train_breakdown.txt 2>&1
python tools/train_breakdown.py --kernels > $O/train_kernels.txt 2>&1
python tools/gemm_shapes.py > $O/gemm_shapes.txt 2>&1
ls -la $O
