# experiment (under gpurun): sub-batches x buffer sets of the streamed pipeline, and paste's shared-memory footprint
mkdir -p gpurun_out
run() { name="$1"; shift; env "$@" timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 $EXTRA > gpurun_out/exp.json 2> gpurun_out/exp.err; python - "$name" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/exp.json").read().strip().splitlines()[-1])
    print(json.dumps({"variant": sys.argv[1], "value": round(d["value"]), "ms_per_step": round(d["ms_per_step"],3), "paste_isolated_ms": round(d["kernels"]["paste+records"]["ms"],3), "roi_isolated_ms": round(d["kernels"]["roi_align_fwd"]["ms"],3), "sum_isolated_ms": round(d["kernels"]["sum_of_isolated_stages_ms"],3)}))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/exp.err").read()[-300:])
PY
}
for c in 1 2 4; do for d in 1 2 3; do EXTRA="--chunks $c --depth $d"; run "chunks$c depth$d" A=1; done; done
EXTRA="--chunks 2 --depth 2"; run "chunks2 depth2 paste single-role 2 CTAs/SM" LCR_PASTE=single
EXTRA="--chunks 1 --depth 2"; run "chunks1 depth2 paste single-role 2 CTAs/SM" LCR_PASTE=single
EXTRA="--chunks 2 --depth 2"; run "chunks2 depth2 paste split 2 CTAs/SM" LCR_PASTE_CTAS=2
