run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", "$*", round(d["ms_per_step"],3), d["roofline"]["stage_ms_profiled_pass"])
except Exception as e: print("$name failed", e)
PY
}
run base2 A=1
run s2_262k_w3 AIGW_CHAT_STREAMS=2 AIGW_CHAT_SUB=262144 AIGW_WALK_CTAS=3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
