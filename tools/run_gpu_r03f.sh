set -x
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -2 gpurun_out/bench_r02_final.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err; tail -2 gpurun_out/bench_r02_final_ref.err
timeout 900 python bench.py --config 4 --steps 2 --warmup 3 > gpurun_out/bench_r02_final_c4.json 2> gpurun_out/bench_r02_final_c4.err; tail -2 gpurun_out/bench_r02_final_c4.err
timeout 1200 python bench.py --config 5 --steps 2 --warmup 1 --bodies 400000 > gpurun_out/bench_r02_final_c5.json 2> gpurun_out/bench_r02_final_c5.err; tail -2 gpurun_out/bench_r02_final_c5.err
timeout 600 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/bench_r02_final_c1.json 2> gpurun_out/bench_r02_final_c1.err; tail -2 gpurun_out/bench_r02_final_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/bench_under_ncu_r02_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chat_ -s 4 -c 4 -f -o gpurun_out/chat_r02_final python bench.py --steps 1 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/ncu_full_r02_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chat_small -c 2 -f -o gpurun_out/small_r02_final python tools/lat_probe.py > gpurun_out/ncu_small_r02_final.log 2>&1
ls -la gpurun_out | tail -8
