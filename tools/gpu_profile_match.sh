#!/bin/bash
# device time of lcr_match_boxes_f32 beside torchvision's chain, then ncu --set full of match_kernel (first launches of each case)
mkdir -p gpurun_out
CMD="python tools/match_probe.py"
$CMD > gpurun_out/r02d_match_probe.jsonl 2> gpurun_out/r02d_match_probe.err &&
ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 6 -c 1 -f -o gpurun_out/r02d_prof_match $CMD > gpurun_out/ncu_match.log 2>&1
echo "capture rc=$?"
cat gpurun_out/r02d_match_probe.jsonl; tail -3 gpurun_out/r02d_match_probe.err
