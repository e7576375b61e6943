#!/bin/bash
# ncu passes on the GPU box (under gpurun).  Each profiled command line first exits 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# one timed step of the region kernels: 2 sub-batches x 7 matching launches; skip the eager step, the capture warm-up and the warm-up steps
ncu --set full --clock-control none --import-source on \
    -k regex:"roi_fwd_|paste_split|rpn_prefilter|rpn_sortfilter|nms_mask|nms_jacobi" -s 70 -c 14 \
    -f -o gpurun_out/prof_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
