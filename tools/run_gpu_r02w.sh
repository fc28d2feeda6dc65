run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", round(d["ms_per_step"],3), d["roofline"]["stage_ms_profiled_pass"]["walk"])
except Exception as e: print("$name failed", e)
PY
}
for w in 3 4; do run b4_w$w AIGW_WALK_CTAS=$w; done
run b4_s2 AIGW_CHAT_STREAMS=2 AIGW_CHAT_SUB=262144
