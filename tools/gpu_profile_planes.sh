#!/bin/bash
# ncu --set full of roi_fwd_planes_kernel (K = 50 and K = 1024 on one NCHW map) and of the generic kernel beside it
mkdir -p gpurun_out
CMD="python tools/planes_probe.py"
$CMD > gpurun_out/planes_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"roi_fwd_planes|roi_fwd_generic" -c 8 -f -o gpurun_out/r02c_prof_planes $CMD > gpurun_out/ncu_planes.log 2>&1
echo "capture rc=$?"
cat gpurun_out/planes_plain.log
