# T5 response direction for the OpenAI backend (stream kinds 8 / 9) + the whole suite again (shared headers moved)
timeout 700 python -m pytest tests/test_messages_openai_response_gpu.py tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r05c.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/pytest_r05c.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
