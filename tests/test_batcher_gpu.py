"""The per-GPU request batcher: concurrent single-request calls must return exactly what the batch call returns."""
import json
import os
import subprocess
import threading

import numpy as np
import pytest

import _oracle as O
import _workload as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def test_concurrent_requests_match_oracle(gw):
    from aigw_b200 import capi
    arena, offs, lens = W.chat_corpus(2, 0, 400)
    bodies = [bytes(arena[int(offs[i]):int(offs[i]) + int(lens[i])]) for i in range(400)]
    bodies += [b'{"model":"m","messages":[{"role":"user","content":"hi"}],"stream":true}', b'{"model":5}', b'not json', b'{"model":"m","messages":[{"role":"nobody","content":"x"}]}']
    cfg = capi.Context.cfg("aws-bedrock")
    b = gw.batcher_start(cfg, max_batch=64, window_us=200)
    results = [None] * len(bodies)
    def work(t):
        for i in range(t, len(bodies), 16):
            results[i] = gw.batcher_translate(b, bodies[i])
    th = [threading.Thread(target=work, args=(t,)) for t in range(16)]
    [t.start() for t in th]; [t.join() for t in th]
    st = gw.batcher_stats(b)
    gw.batcher_stop(b)
    assert st["requests"] == len(bodies) and st["batches"] <= len(bodies)
    for body, (rc, status, reason, path, out, kind) in zip(bodies, results):
        t = O.chat_translate("aws-bedrock", body)
        assert rc == 0 and status == t.status, (body[:100], rc, status, t.status)
        if status == 0:
            assert out == t.body and path.decode() == t.path


def test_two_backends_and_large_bodies(gw):
    """one batcher, two backends (each with its own batches); bodies above the fused kernel's class take the host call"""
    from aigw_b200 import capi
    arena, offs, lens = W.chat_corpus(4, 0, 120)
    bodies = [bytes(arena[int(offs[i]):int(offs[i]) + int(lens[i])]) for i in range(120)]
    big_arena, bo, bl = W.chat_corpus(4, 0, 6, target=12000, jitter=64)
    bodies += [bytes(big_arena[int(bo[i]):int(bo[i]) + int(bl[i])]) for i in range(6)]
    b = gw.batcher_start(capi.Context.cfg("aws-bedrock"), max_batch=32, window_us=100)
    k = gw.batcher_add_backend(b, capi.Context.cfg("openai"))
    assert k == 1
    results = {}
    def work(t):
        for i in range(t, len(bodies), 8):
            results[(i, 0)] = gw.batcher_translate(b, bodies[i], backend=0)
            results[(i, 1)] = gw.batcher_translate(b, bodies[i], backend=1)
    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    gw.batcher_stop(b)
    for (i, k), (rc, status, reason, path, out, kind) in results.items():
        t = O.chat_translate("aws-bedrock" if k == 0 else "openai", bodies[i])
        assert rc == 0 and status == t.status, (i, k, rc, status, reason, t.status)
        if status == 0:
            assert out == (t.body if t.body_kind == O.BYTES else b"") and path.decode() == t.path, (i, k)


def test_small_output_buffer_is_reported(gw):
    from aigw_b200 import capi
    b = gw.batcher_start(capi.Context.cfg("aws-bedrock"), max_batch=8, window_us=10)
    rc, status, *_ = gw.batcher_translate(b, b'{"model":"m","messages":[{"role":"user","content":"' + b"x" * 500 + b'"}]}', out_cap=64)
    gw.batcher_stop(b)
    assert rc == -4


def test_load_generator(gw, tmp_path):
    """C++ threads (what goroutines behind cgo would be): every request completes, batches form under concurrency."""
    exe = str(tmp_path / "batcher_load")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "batcher_load.cpp"), "-L", os.path.join(ROOT, "aigw_b200"),
                           "-laigw_b200", "-Wl,-rpath," + os.path.join(ROOT, "aigw_b200"), "-lpthread", "-o", exe])
    arena, offs, lens = W.chat_corpus(2, 0, 2000)
    packed = b"".join(bytes(arena[int(offs[i]):int(offs[i]) + int(lens[i])]) for i in range(2000))
    po = np.zeros(2001, dtype=np.uint64); np.cumsum(lens[:2000], out=po[1:])
    (tmp_path / "bodies.bin").write_bytes(packed); (tmp_path / "offs.u64").write_bytes(po.tobytes())
    out = subprocess.check_output([exe, str(tmp_path / "bodies.bin"), str(tmp_path / "offs.u64"), "2000", "64", "200", "256", "100"]).decode()
    d = json.loads(out.strip().split("\n")[-1])
    assert d["failed"] == 0 and d["declined"] == 0 and d["ok"] == 64 * 200
    assert d["mean_batch"] > 4
