# experiment (under gpurun): the streamed step with the r02b RoIAlign against paste's shared-memory footprint
# (ring slots x zero-buffer size: 107 KB default -> 2 RoIAlign CTAs beside paste per SM; <= 70 KB -> 3)
mkdir -p gpurun_out
run() { name="$1"; shift; env "$@" timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 $EXTRA > gpurun_out/exp.json 2> gpurun_out/exp.err; python - "$name" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/exp.json").read().strip().splitlines()[-1])
    print(json.dumps({"variant": sys.argv[1], "value": round(d["value"]), "ms_per_step": round(d["ms_per_step"],3), "paste_isolated_ms": round(d["kernels"]["paste+records"]["ms"],3), "roi_isolated_ms": round(d["kernels"]["roi_align_fwd"]["ms"],3), "step_frac_of_hbm": round(d["step_roofline"]["frac"],3)}))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/exp.err").read()[-300:])
PY
}
EXTRA=""
run "default (8 slots, 16 KB zeros)" A=1
run "4 slots" LCR_PASTE_SLOTS=4
run "2 slots" LCR_PASTE_SLOTS=2
run "4 slots, 8 KB zeros" LCR_PASTE_SLOTS=4 LCR_PASTE_ZB_KB=8
run "8 slots, 32 KB zeros" LCR_PASTE_ZB_KB=32
run "4 slots, 32 KB zeros" LCR_PASTE_SLOTS=4 LCR_PASTE_ZB_KB=32
run "chunks 1" A=1
EXTRA="--chunks 1"; run "chunks 1, 4 slots" LCR_PASTE_SLOTS=4
EXTRA="--chunks 1"; run "chunks 1" A=1
