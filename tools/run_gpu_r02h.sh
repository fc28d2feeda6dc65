timeout 1200 python -m pytest tests/test_chat_gpu.py -x -q -m gpu -k "gemini or anthropic or bedrock or passthrough" 2>&1 | tail -30
