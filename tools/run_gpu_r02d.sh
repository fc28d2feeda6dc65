run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --bodies 1000000 --skip-e2e --cpu-sample 2000 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name', round(d['ms_per_step'],2), round(d['roofline']['frac'],4))"
}
run s1 AIGW_CHAT_STREAMS=1
run s2_128k AIGW_CHAT_STREAMS=2 AIGW_CHAT_SUB=131072
run s3_64k AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=65536
run s3_64k_i3e2 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=65536 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2
run s3_64k_i3e2_w8 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=65536 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2 AIGW_WALK_VARIANT=1
run s4_64k_i2e2_w8 AIGW_CHAT_STREAMS=4 AIGW_CHAT_SUB=65536 AIGW_IDX_CTAS=2 AIGW_EMIT_CTAS=2 AIGW_WALK_VARIANT=1
run s3_128k_i3e2_w8 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=131072 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2 AIGW_WALK_VARIANT=1
run s4_32k_i3e2_w8 AIGW_CHAT_STREAMS=4 AIGW_CHAT_SUB=32768 AIGW_IDX_CTAS=3 AIGW_EMIT_CTAS=2 AIGW_WALK_VARIANT=1
run s3_64k_i4e3 AIGW_CHAT_STREAMS=3 AIGW_CHAT_SUB=65536 AIGW_IDX_CTAS=4 AIGW_EMIT_CTAS=3
