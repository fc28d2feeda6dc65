# /v1/messages error translators (OpenAI + Bedrock) and the whole suite
timeout 700 python -m pytest tests/test_response_error_gpu.py tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r05g.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/pytest_r05g.log
