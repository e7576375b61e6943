#!/bin/bash
mkdir -p gpurun_out
python tools/bench_kernels.py --reps 10 --only roi,bwd,paste,select,nms,roiexp,overlap > gpurun_out/r01d_kernel_ab.jsonl 2> gpurun_out/r01d_kernel_ab.err; echo "ab rc=$?"
python tests/perf_config_latency.py > gpurun_out/r01d_config_latency.log 2>&1; echo "lat rc=$?"; tail -2 gpurun_out/r01d_config_latency.log | cut -c1-300
python tools/roi_sweep.py --reps 10 > gpurun_out/r01d_roi_sweep.jsonl 2> gpurun_out/r01d_roi_sweep.err; echo "sweep rc=$?"
bash tools/gpu_profile.sh 2>&1 | grep rc=
