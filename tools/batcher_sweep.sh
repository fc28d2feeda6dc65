#!/bin/bash
# latency / throughput of the request batcher under concurrency (C++ threads as stand-ins for goroutines behind cgo)
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
for cfg in "1 2000 1 0" "8 2000 64 20" "64 1000 256 50" "512 400 1024 100" "2048 200 4096 200" "8192 50 8192 300"; do
  set -- $cfg
  /tmp/batcher_load /tmp/bl_bodies.bin /tmp/bl_offs.u64 20000 $1 $2 $3 $4
done
