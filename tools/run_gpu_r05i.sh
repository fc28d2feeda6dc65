# integer fields with a fraction / exponent -> ParseBody 400; whole suite; accept-rate line of the bench
timeout 700 python -m pytest tests/test_chat_gpu.py tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r05i.log 2>&1; echo "pytest rc $?"; tail -12 gpurun_out/pytest_r05i.log
timeout 100 python tools/diverse_reasons.py 2>&1 | cut -c1-160 | head -6
