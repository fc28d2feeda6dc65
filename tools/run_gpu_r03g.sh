timeout 900 python -m pytest tests/test_bpe_gpu.py -m gpu -x -q 2>&1 | tail -12
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -3 gpurun_out/bench_r02_c3.err; cat gpurun_out/bench_r02_c3.json | head -c 2500
