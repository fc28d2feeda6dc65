# OpenAI passthrough error translator, several contexts in one process, the whole suite
timeout 700 python -m pytest tests/test_multi_context_gpu.py tests/test_response_error_gpu.py tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/pytest_r05h.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/pytest_r05h.log
