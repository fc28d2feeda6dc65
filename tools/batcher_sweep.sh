#!/bin/bash
# latency / throughput of the request batcher under concurrency (C++ threads as stand-ins for goroutines behind cgo)
# each line: threads requests_per_thread max_batch window_us
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
g++ -O2 -std=c++17 -I $ROOT/include $ROOT/tools/batcher_load.cpp -L $ROOT/aigw_b200 -laigw_b200 -Wl,-rpath,$ROOT/aigw_b200 -lpthread -o /tmp/batcher_load
python - <<PY
import sys; sys.path.insert(0, "$ROOT/tests"); sys.path.insert(0, "$ROOT")
import numpy as np, _workload as W
arena, offs, lens = W.chat_corpus(2, 0, 20000)
open("/tmp/bl_bodies.bin", "wb").write(b"".join(bytes(arena[int(offs[i]):int(offs[i]) + int(lens[i])]) for i in range(20000)))
po = np.zeros(20001, dtype=np.uint64); np.cumsum(lens[:20000], out=po[1:]); open("/tmp/bl_offs.u64", "wb").write(po.tobytes())
PY
nproc
for cfg in "1 3000 1 0" "8 3000 8 5" "64 2000 64 20" "128 2000 128 30" "256 2000 256 40" "512 1500 256 50" "512 1500 512 50" "1024 1000 512 50" "2048 500 512 50" "2048 500 1024 80"; do
  set -- $cfg
  timeout 120 /tmp/batcher_load /tmp/bl_bodies.bin /tmp/bl_offs.u64 20000 $1 $2 $3 $4 || echo "FAILED $cfg"
done
