#!/bin/bash
# Runs on the GPU box (under gpurun): per-file GPU parity tests (separate processes so that a faulting kernel cannot poison
# the other files), smoke(), then a short bench.  Logs land in gpurun_out/.  The exit status reflects EVERY step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "$name rc=$rc: $(tail -1 gpurun_out/$name.log)"
  [ $rc -ne 0 ] && status=1
done
# the alternative kernels behind the tuning switches are covered inside -m gpu (tests/test_gpu_alternates.py and the
# parametrised RoIAlign tests); the switches also work from the environment (read once at library load):
LCR_SELECT=general LCR_NMS_RESOLVE=serial timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_nms.py tests/test_gpu_pipeline.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_env_select_nms.log 2>&1
rc=$?; echo "env switches (general select, serial NMS resolve) rc=$rc: $(tail -1 gpurun_out/test_env_select_nms.log)"; [ $rc -ne 0 ] && status=1
LCR_ROI_FWD=warp LCR_PASTE=single LCR_TORCH_EXT=0 timeout 600 python -m pytest tests/test_gpu_roi_align.py tests/test_gpu_paste.py tests/test_gpu_pipeline.py tests/test_gpu_fuzz.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_env_roi_paste.log 2>&1
rc=$?; echo "env switches (sample-walk RoIAlign, single-role paste, ctypes call path) rc=$rc: $(tail -1 gpurun_out/test_env_roi_paste.log)"; [ $rc -ne 0 ] && status=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc: $(tail -1 gpurun_out/smoke.log)"; [ $rc -ne 0 ] && status=1
if [ "$1" != "--no-bench" ]; then
  timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  rc=$?; echo "bench rc=$rc"; [ $rc -ne 0 ] && status=1
  tail -c 1500 gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
fi
exit $status
