#!/bin/bash
# r02b evidence for the RoIAlign forward work (under gpurun): A/B of every forward organisation, list-composition experiments,
# K sweep, the streamed step with each forward kernel, ncu --set full of the default and the persistent kernel.
mkdir -p gpurun_out
python tools/bench_kernels.py --reps 10 --only roi,roiexp > gpurun_out/r02b_roi_ab.jsonl 2> gpurun_out/r02b_roi_ab.err; echo "ab rc=$?"
python tools/roi_mix_exp.py > gpurun_out/r02b_roi_mix_exp.jsonl 2> gpurun_out/r02b_roi_mix_exp.err; echo "mix rc=$?"
python tools/roi_ksweep.py > gpurun_out/r02b_roi_ksweep.jsonl 2> gpurun_out/r02b_roi_ksweep.err; echo "ksweep rc=$?"
for v in default warp team rm; do
  if [ $v = default ]; then unset LCR_ROI_FWD; else export LCR_ROI_FWD=$v; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-python --no-extras > gpurun_out/r02b_bench_fwd_$v.json 2> gpurun_out/r02b_bench_fwd_$v.err; echo "bench $v rc=$?"
done
unset LCR_ROI_FWD
bash tools/gpu_profile_roi.sh; mv gpurun_out/prof_roi_warp.ncu-rep gpurun_out/r02b_prof_roi_default.ncu-rep
LCR_ROI_FWD=team bash tools/gpu_profile_roi.sh; mv gpurun_out/prof_roi_team.ncu-rep gpurun_out/r02b_prof_roi_team.ncu-rep
