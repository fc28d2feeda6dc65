# config 3 on JSON requests: tests, bench line, ncu launch list and one full capture of the request scan kernel
timeout 300 python -m pytest tests/test_embeddings_count_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -2 gpurun_out/bench_r02_c3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c3.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['scan_kernel_alone'], d['e2e'], d['cpu_baseline']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02_c3.csv python bench.py --config 3 --requests 1024 --steps 2 --warmup 1 --skip-e2e > gpurun_out/bench_under_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emb_scan -s 1 -c 1 -f -o gpurun_out/emb_scan_r02 python bench.py --config 3 --requests 1024 --steps 1 --warmup 1 --skip-e2e > gpurun_out/ncu_emb_scan.log 2>&1
ls -la gpurun_out/emb_scan_r02.ncu-rep gpurun_out/launches_r02_c3.csv
