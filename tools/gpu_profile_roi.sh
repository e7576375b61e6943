#!/bin/bash
# ncu --set full of the staged RoIAlign forward kernel on the bench RoI list (under gpurun; plain run first)
mkdir -p gpurun_out
export LCR_ROI_FWD=staged
export LCR_ROI_STAGED_WARPS=${LCR_ROI_STAGED_WARPS:-4}
CMD="python tools/bench_kernels.py --only roi1 --reps 2"
$CMD > gpurun_out/roi1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"roi_fwd_staged" -s 3 -c 1 -f -o gpurun_out/prof_roi_staged $CMD > gpurun_out/ncu_roi_staged.log 2>&1
echo "staged capture rc=$?"
ls -la gpurun_out/*.ncu-rep
