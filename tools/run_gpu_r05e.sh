# pipelined e2e (two calls in flight through two contexts) vs the serial variant, and a small sweep of lanes / parts
for cfg in "2 8" "1 8" "3 12" "2 16" "4 16"; do set -- $cfg; ( time timeout 300 python bench.py --e2e-lanes $1 --e2e-parts $2 > gpurun_out/bench_r05e_$1_$2.json 2> gpurun_out/bench_r05e_$1_$2.err ) 2>&1 | grep real; tail -1 gpurun_out/bench_r05e_$1_$2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/bench_r05e_$1_$2.json').read().strip().split('\n')[-1]); print('lanes $1 parts $2', d['value'], d['e2e'])"; done
