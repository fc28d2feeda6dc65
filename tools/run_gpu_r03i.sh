for t in 1 2 4 8; do AIGW_HOST_THREADS=$t AIGW_STREAM_PROF=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 1 2> gpurun_out/c4_$t.err > gpurun_out/c4_$t.json; echo "threads $t"; grep stream_chunks gpurun_out/c4_$t.err | tail -2; python -c "
import json; d=json.loads([l for l in open('gpurun_out/c4_$t.json') if l.startswith('{')][-1]); print(d['value'])"; done
