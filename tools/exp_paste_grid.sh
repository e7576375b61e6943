# experiment (under gpurun): paste as a persistent kernel on FEWER than all SMs (the rest keep 4 RoIAlign CTAs)
mkdir -p gpurun_out
for g in 148 120 100 74 56; do
LCR_PASTE_GRID=$g timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/exp.json 2> gpurun_out/exp.err; python - "$g" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/exp.json").read().strip().splitlines()[-1])
    print(json.dumps({"paste_ctas": int(sys.argv[1]), "value": round(d["value"]), "ms_per_step": round(d["ms_per_step"],3), "paste_isolated_ms": round(d["kernels"]["paste+records"]["ms"],3), "paste_alone_frac": round(d["roofline"]["frac"],3), "roi_isolated_ms": round(d["kernels"]["roi_align_fwd"]["ms"],3)}))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/exp.err").read()[-300:])
PY
done
