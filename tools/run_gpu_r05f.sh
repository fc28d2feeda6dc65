# /v1/messages error translator (OpenAI backend) + the whole suite + the default bench line (single-call e2e path after the bench.py refactor)
timeout 700 python -m pytest tests/test_response_error_gpu.py tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r05f.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/pytest_r05f.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 400 python bench.py > gpurun_out/bench_r05f.json 2> gpurun_out/bench_r05f.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_r05f.err | cut -c1-400; python -c "
import json; d=json.loads(open('gpurun_out/bench_r05f.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['p50_added_us'])"
