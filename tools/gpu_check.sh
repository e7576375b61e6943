#!/bin/bash
# Runs on the GPU box (under gpurun): per-file GPU parity tests (separate processes so that a faulting
# kernel cannot poison the other files), smoke(), then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for f in tests/test_gpu_match.py tests/test_gpu_boxes.py tests/test_gpu_select.py tests/test_gpu_nms.py tests/test_gpu_roi_align.py tests/test_gpu_paste.py tests/test_gpu_pipeline.py tests/test_gpu_guards.py tests/test_gpu_properties.py tests/test_gpu_fuzz.py; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "$name rc=$rc: $(tail -1 gpurun_out/$name.log)"
  [ $rc -ne 0 ] && status=1
done
# the alternative kernels behind the tuning switches stay covered
LCR_SELECT=general timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_pipeline.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_alt_select.log 2>&1
echo "alt select (general cluster kernel) rc=$?: $(tail -1 gpurun_out/test_alt_select.log)"
LCR_ROI_FWD=cta LCR_ROI_BWD=cta LCR_PASTE=rows16 timeout 600 python -m pytest tests/test_gpu_roi_align.py tests/test_gpu_paste.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_alt_roi_paste.log 2>&1
echo "alt roi/paste (per-CTA RoIAlign, per-thread paste) rc=$?: $(tail -1 gpurun_out/test_alt_roi_paste.log)"
LCR_NMS_RESOLVE=serial timeout 600 python -m pytest tests/test_gpu_nms.py tests/test_gpu_pipeline.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_alt_nms.log 2>&1
echo "alt nms (serial chunk resolve) rc=$?: $(tail -1 gpurun_out/test_alt_nms.log)"
LCR_TORCH_EXT=0 timeout 600 python -m pytest tests/test_gpu_roi_align.py tests/test_gpu_nms.py tests/test_gpu_pipeline.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_alt_ctypes.log 2>&1
echo "alt call path (ctypes + Python autograd instead of csrc/lcr_torch.so) rc=$?: $(tail -1 gpurun_out/test_alt_ctypes.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?: $(tail -1 gpurun_out/smoke.log)"
if [ "$1" != "--no-bench" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"
  tail -c 3000 gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
fi
exit $status
