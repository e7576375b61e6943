#!/bin/bash
# ncu passes on the GPU box (under gpurun).  Each profiled command line first exits 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# one whole timed step of the region kernels (8 matching launches per step; skip the 3 warm-up steps)
ncu --set full --clock-control none --import-source on \
    -k regex:"roi_fwd_warp|paste_split|paste_bulk|rpn_prefilter|rpn_sortfilter|nms_resolve|nms_mask|nms_jacobi" -s 24 -c 8 \
    -f -o gpurun_out/prof_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
