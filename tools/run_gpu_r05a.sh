# round 2, session 3: full GPU suite (with the T4 sampling parameters and the /v1/completions kinds), smoke, headline bench line
timeout 700 python -m pytest tests/test_completions_gpu.py tests/test_chat_gpu.py tests -m gpu -q --durations=12 -p no:cacheprovider > gpurun_out/pytest_r05a.log 2>&1; echo "pytest rc $?"; tail -30 gpurun_out/pytest_r05a.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_r05a.json 2> gpurun_out/bench_r05a.err; echo "bench rc $?"; tail -2 gpurun_out/bench_r05a.err; head -c 1800 gpurun_out/bench_r05a.json
