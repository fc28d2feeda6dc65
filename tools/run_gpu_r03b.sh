timeout 600 python -m pytest tests/test_batcher_gpu.py tests/test_small_batch_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 900 bash tools/batcher_sweep.sh 2>&1 | tail -14
