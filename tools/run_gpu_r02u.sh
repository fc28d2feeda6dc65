run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", round(d["ms_per_step"],3), d["roofline"]["stage_ms_profiled_pass"]["walk"])
except Exception as e: print("$name failed", e)
PY
}
for v in NONE NO_WIN NO_PF; do for w in 2 3 4 6; do run ${v}_w$w AIGW_B200_SO=/root/repo/variants/lib$v.so AIGW_WALK_CTAS=$w; done; done
