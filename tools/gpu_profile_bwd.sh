#!/bin/bash
# ncu --set full of the RoIAlign backward kernel: bench RoI list and BASELINE C5 (16384 RoIs on one map) — under gpurun; plain run first
mkdir -p gpurun_out
CMD="python tools/bench_kernels.py --only bwd1 --reps 2"
$CMD > gpurun_out/bwd1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"roi_bwd_warp" -s 6 -c 4 -f -o gpurun_out/prof_roi_bwd $CMD > gpurun_out/ncu_roi_bwd.log 2>&1
echo "bwd capture rc=$?"
cat gpurun_out/bwd1_plain.log | tail -3
