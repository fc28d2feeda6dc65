run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", "$*", round(d["ms_per_step"],3))
except Exception as e: print("$name failed", e)
PY
}
run base A=1
run p131_232 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072
run p65_232 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=65536
run p262_232 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=262144
run p131_332 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3
run p131_333 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=3
run p131_222 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_WALK_CTAS=2
run p131_242 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_WALK_CTAS=4
run p131_322 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3 AIGW_WALK_CTAS=2
run p131_565 AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=5 AIGW_WALK_CTAS=6 AIGW_EMIT_CTAS=5
timeout 600 env AIGW_CHAT_PIPE=1 AIGW_CHAT_SUB=16384 python -m pytest tests/test_chat_gpu.py tests/test_escapes_gpu.py -m gpu -x -q 2>&1 | tail -3
