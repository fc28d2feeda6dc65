set -x
export AIGW_CHAT_SUB=131072
timeout 600 python -m pytest tests/test_chat_gpu.py tests/test_embeddings_gpu.py tests/test_bedrock_response_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --bodies 1000000 --skip-e2e --cpu-sample 2000 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
bash tools/profile_kernel.sh chat_walk walk_r02a 1
bash tools/profile_kernel.sh chat_index idx_r02b 2
bash tools/profile_kernel.sh chat_emit emit_r02b 2
