set -x
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_chat_gpu.py tests/test_embeddings_gpu.py tests/test_bedrock_response_gpu.py tests/test_sha256_gpu.py tests/test_batcher_gpu.py -x -q -m gpu 2>&1 | tail -15
for sub in 4096 8192 16384 32768 131072; do
  AIGW_CHAT_SUB=$sub timeout 600 python bench.py --steps 3 --warmup 2 --skip-e2e --cpu-sample 20000 > gpurun_out/bench_sub$sub.json 2> gpurun_out/bench_sub$sub.err; tail -c 1500 gpurun_out/bench_sub$sub.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($sub, d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('stage_ms_serialised'))"
done
