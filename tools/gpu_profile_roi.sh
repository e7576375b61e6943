#!/bin/bash
# ncu --set full of ONE RoIAlign forward variant on the bench RoI list (under gpurun; plain run first)
#   LCR_ROI_FWD=staged|rm|rm1|(unset: warp kernel)  bash tools/gpu_profile_roi.sh
mkdir -p gpurun_out
CMD="python tools/bench_kernels.py --only roi1 --reps 2"
$CMD > gpurun_out/roi1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"roi_fwd_" -s 3 -c 1 -f -o gpurun_out/prof_roi_${LCR_ROI_FWD:-warp} $CMD > gpurun_out/ncu_roi.log 2>&1
echo "capture rc=$?"
tail -1 gpurun_out/roi1_plain.log
