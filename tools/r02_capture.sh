#!/bin/bash
# Round-2 evidence capture on one B200 (run under gpurun): launch list of one sampling step, ncu --set full of the kernels
# changed / added this round, every bench workload, the per-op training breakdown.  Outputs under gpurun_out/r02/.
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_step.csv \
    python bench.py --steps 1 --warmup 3 --profile-eager > $O/launches_step.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv 119 > $O/launches_step.md
ncu --set full --clock-control none --import-source on -k regex:"triattn_flash_g4|pair_transition_ws" -c 4 -f -o $O/rows \
    python tools/kprof.py op_triattn_start op_pair_fc --iters 1 > $O/ncu_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"triattn_flash_g4" -c 2 -f -o $O/flash_ragged \
    python tools/kprof.py op_triattn_start --iters 1 --pad 60 > $O/ncu_flash_ragged.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bw_dw_tc|gemm_f16_kernel<128, 5, true>|bw_attn" -c 12 -f -o $O/bwd \
    python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import gpu_cases
print(gpu_cases.case_bwd_triattn(B=2, N=256, pad=9))" > $O/ncu_bwd.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in config1 config2 config5; do python bench.py --workload $w --steps 30 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
for m in graph manual autograd; do python bench.py --workload train --train-mode $m --steps 10 --warmup 3 > $O/bench_train_$m.json 2> $O/bench_train_$m.err; done
python bench.py --workload train --train-mode graph --train-sizes 2 --steps 10 --warmup 3 > $O/bench_train_graph_sizes2.json 2>&1
python tools/train_breakdown.py > $O/train_breakdown.txt 2>&1
tools/ab_bench.sh PRD_PDL 1 0 > $O/pdl_ab.txt 2>&1
ls -la $O
