set -x
timeout 900 python bench.py > gpurun_out/bench_r02_main.json 2> gpurun_out/bench_r02_main.err; tail -2 gpurun_out/bench_r02_main.err
AIGW_HOST_ZEROCOPY=1 timeout 900 python bench.py --steps 3 --cpu-sample 20000 > gpurun_out/bench_r02_zerocopy.json 2> gpurun_out/bench_r02_zerocopy.err; tail -2 gpurun_out/bench_r02_zerocopy.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; tail -2 gpurun_out/bench_r02_ref.err
timeout 900 python bench.py --config 4 --steps 2 --warmup 3 > gpurun_out/bench_r02_c4.json 2> gpurun_out/bench_r02_c4.err; tail -2 gpurun_out/bench_r02_c4.err
timeout 1200 python bench.py --config 5 --steps 2 --warmup 1 --bodies 400000 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; tail -2 gpurun_out/bench_r02_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chat_ -s 4 -c 4 -f -o gpurun_out/chat_r02 python bench.py --steps 1 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/ncu_full_r02.log 2>&1
ls -la gpurun_out | tail -12
