timeout 600 python -m pytest tests/test_small_batch_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/lat_probe.py 2>&1 | python -c "
import sys,json
d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a:round(b,1) for a,b in v.items()})"
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_fused.json 2> gpurun_out/pp_fused.err; python -c "
import json; d=json.load(open('gpurun_out/pp_fused.json')); print(d['ms_per_step'], d['roofline']['stage_ms_profiled_pass'])"
