#!/bin/bash
# Per-kernel durations of one FoldingBlock at the north-star size (ncu launch list, not a bench number).
# Usage (on the GPU box): tools/launch_times.sh [tag]
tag="${1:-x}"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_block_$tag.csv \
  python tools/kprof.py op_block --iters 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_block_$tag.csv 28 | head -30
