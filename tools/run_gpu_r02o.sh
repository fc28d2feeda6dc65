set -x
ncu --set full --clock-control none --import-source on -k regex:chat_ -s 4 -c 4 -f -o gpurun_out/chat_r02b python bench.py --steps 1 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/ncu_full_r02b.log 2>&1
ls -la gpurun_out | tail -4
