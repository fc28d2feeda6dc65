# 16- / 17-digit floats (shortest_f64.cuh) in the walk: whole suite, the device-resident bench line (walk registers 93 -> 98: check the step time), decline reasons
timeout 400 python -m pytest tests/test_chat_gpu.py tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 120 python bench.py --skip-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('ms_per_step', d['ms_per_step'], d['roofline']['stage_ms_profiled_pass'], d.get('accept_rate'))"
timeout 60 python tools/diverse_reasons.py 2>&1 | cut -c1-100 | head -4
