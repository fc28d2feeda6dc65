set -x
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/ab_split.json 2> gpurun_out/ab_split.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_split.json")); print("split", d["ms_per_step"], d["roofline"]["stage_ms_profiled_pass"])
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
