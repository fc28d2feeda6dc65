run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --bodies 1000000 --skip-e2e --cpu-sample 2000 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name', round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['roofline']['stage_ms_serialised'])"
}
timeout 600 python -m pytest tests/test_chat_gpu.py tests/test_embeddings_gpu.py tests/test_bedrock_response_gpu.py -x -q -m gpu 2>&1 | tail -3
run base
run w8 AIGW_WALK_VARIANT=1
