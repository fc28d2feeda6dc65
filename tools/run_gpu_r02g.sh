timeout 1200 python -m pytest tests/test_anthropic_response_gpu.py tests/test_bedrock_response_gpu.py tests/test_chat_gpu.py tests/test_embeddings_gpu.py -x -q -m gpu 2>&1 | tail -30
