timeout 600 python -m pytest tests/test_small_batch_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 2000 --skip-e2e > gpurun_out/pp_fused.json 2> gpurun_out/pp_fused.err; python -c "
import json; d=json.load(open('gpurun_out/pp_fused.json')); print(d['ms_per_step'], d['roofline']['stage_ms_profiled_pass'])"
