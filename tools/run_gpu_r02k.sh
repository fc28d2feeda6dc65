set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 3 --bodies 200000 --cpu-sample 20000 > gpurun_out/b2_small.json 2> gpurun_out/b2_small.err; tail -3 gpurun_out/b2_small.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --bodies 200000 --cpu-sample 20000 > gpurun_out/bref_small.json 2> gpurun_out/bref_small.err; tail -3 gpurun_out/bref_small.err
timeout 300 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -3 gpurun_out/b1.err
timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --streams 20000 > gpurun_out/b4_small.json 2> gpurun_out/b4_small.err; tail -3 gpurun_out/b4_small.err
timeout 600 python bench.py --config 5 --steps 2 --warmup 1 --bodies 100000 > gpurun_out/b5_small.json 2> gpurun_out/b5_small.err; tail -3 gpurun_out/b5_small.err
