run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", round(d["ms_per_step"],3), d["roofline"]["stage_ms_profiled_pass"])
except Exception as e: print("$name failed", e)
PY
}
run carve A=1
run nocarve AIGW_WALK_NO_CARVEOUT=1
run carve_w3 AIGW_WALK_CTAS=3
run carve_w5 AIGW_WALK_CTAS=5
