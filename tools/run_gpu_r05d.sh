# /v1/messages on Anthropic behind AWS Bedrock, stream (kind 10) + the whole suite
timeout 700 python -m pytest tests/test_messages_aws_anthropic_stream_gpu.py tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r05d.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/pytest_r05d.log
