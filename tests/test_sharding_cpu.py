"""N>1 host logic on CPU (gloo, world_size 2): ranks own disjoint shards of the seeded workload, the union equals the
single-process workload, and the reduction bench.py performs (SUM of bodies, MAX of time) behaves."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import _workload as W
    from aigw_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.bench_shard(rank, 500)
    lens = W.chat_lens(2, first, n, threads=1)
    t = torch.tensor([float(lens.sum()), float(n)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tm = torch.tensor([0.5 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, first, n, int(lens.sum()), t.tolist(), tm.item()))
    dist.destroy_process_group()


def test_two_rank_shards_cover_the_workload():
    import torch.multiprocessing as mp
    import _workload as W
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    whole = W.chat_lens(2, 0, 1000, threads=1)
    assert out[0][1] == 0 and out[1][1] == 500 and out[0][2] == out[1][2] == 500
    assert out[0][3] + out[1][3] == int(whole.sum())
    assert out[0][4] == out[1][4] == [float(whole.sum()), 1000.0]
    assert out[0][5] == out[1][5] == 1.5


def test_request_placement_is_stable_and_balanced():
    from aigw_b200 import shard
    ids = [f"req-{i}-9f1c".encode() for i in range(8000)]
    dev = np.array([shard.device_for_request(r, 8) for r in ids])
    assert all(shard.device_for_request(r, 8) == d for r, d in zip(ids[:50], dev[:50]))
    counts = np.bincount(dev, minlength=8)
    assert counts.min() > 800 and counts.max() < 1200
    # router stream and upstream stream of one request share the internal id ⇒ same device
    assert shard.device_for_request(b"abc-123", 8) == shard.device_for_request(b"abc-123", 8)
