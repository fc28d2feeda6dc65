set -x
for v in main B walk_2a3b971 walk_ebab585 main2; do
  case $v in main|main2) so="";; *) so=/root/repo/variants/lib$v.so;; esac
  AIGW_B200_SO=$so timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["stage_ms_profiled_pass"], d["clocks"])
except Exception as e: print("$v failed", e)
PY
done
timeout 900 python -m pytest tests/test_escapes_gpu.py tests/test_chat_gpu.py -m gpu -x -q 2>&1 | tail -5
