set -x
export AIGW_CHAT_SUB=131072
for v in 0 1 2 3 4 5; do
  AIGW_WALK_VARIANT=$v timeout 600 python bench.py --steps 2 --warmup 2 --bodies 400000 --skip-e2e --cpu-sample 2000 > gpurun_out/bench_wv$v.json 2> gpurun_out/bench_wv$v.err
done
bash tools/profile_kernel.sh chat_index idx_r02a 2
bash tools/profile_kernel.sh chat_emit emit_r02a 2
