run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/ov_$name.json 2> gpurun_out/ov_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ov_$name.json")); print("$name", "$*", round(d["ms_per_step"],3))
except Exception as e: print("$name failed", e)
PY
}
run base A=1
run s3_131k_i3e2w3 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2 AIGW_WALK_CTAS=3
run s3_131k_i2e2w4 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=2 AIGW_EMIT_CTAS=2 AIGW_WALK_CTAS=4
run s2_131k_i3e3w3 AIGW_CHAT_STREAMS=2 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=3 AIGW_WALK_CTAS=3
run s4_65k_i2e2w3 AIGW_CHAT_STREAMS=4 AIGW_CHAT_SUB=65536 AIGW_IDX_CTAS=2 AIGW_EMIT_CTAS=2 AIGW_WALK_CTAS=3
run s3_131k AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=131072
run s3_262k_i3e2w3 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=262144 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2 AIGW_WALK_CTAS=3
run s2_262k_w3 AIGW_CHAT_STREAMS=2 AIGW_CHAT_SUB=262144 AIGW_WALK_CTAS=3
