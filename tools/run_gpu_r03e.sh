run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 > gpurun_out/pp_$name.json 2> gpurun_out/pp_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pp_$name.json")); print("$name", round(d["ms_per_step"],3), d["e2e"]["value"], d["e2e"]["pcie_gbs_each_way"], d.get("p50_added_us"))
except Exception as e: print("$name failed", e)
PY
}
run chunk256 A=1
run chunk128 AIGW_CHUNK_MB=128
run chunk64 AIGW_CHUNK_MB=64
run chunk32 AIGW_CHUNK_MB=32
