mkdir -p gpurun_out
python -m pytest tests/test_gpu_roi_align.py -x -q -m gpu -k "row_major" > gpurun_out/t_roi.log 2>&1; echo "roi tests rc=$?"; tail -3 gpurun_out/t_roi.log
python tools/bench_kernels.py --only roi --reps 8 > gpurun_out/roi_ab.jsonl 2>gpurun_out/roi_ab.err; tail -2 gpurun_out/roi_ab.err
python - <<'PY'
import json
for l in open("gpurun_out/roi_ab.jsonl"):
    d = json.loads(l); print(f"{d['ms']:.3f} {d['frac']:.3f} err={d['rel_err_vs_first']:.1e}  {d['variant']}")
PY
