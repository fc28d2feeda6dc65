"""Several contexts in one process (per-device one-time setup, cudaSetDevice at every entry point): two contexts on one GPU driven
from two threads at once — the bench's --e2e-lanes mode and a gateway with several worker threads do exactly this — and, when the box has a
second GPU, one context per device.  Every result is checked against the oracle."""
import threading

import pytest

import _oracle as O
import _workload as W

pytestmark = pytest.mark.gpu


def _bodies(seed, n):
    arena, offs, lens = W.chat_corpus(2, seed, n)
    return [bytes(arena[int(offs[i]):int(offs[i]) + int(lens[i])]) for i in range(n)]


def _check(ctx, bodies, errors, rounds=3):
    import aigw_b200 as A
    try:
        exp = [O.chat_translate("aws-bedrock", b) for b in bodies]
        for _ in range(rounds):
            got = ctx.chat_translate(ctx.cfg("aws-bedrock"), bodies)
            for g, o in zip(got, exp):
                assert g["status"] == A.AIGW_OK and o.status == O.OK and g["body"] == o.body and g["path"].decode() == o.path
            (h,) = ctx.stream_open("openai", b"req-model")
            r = ctx.stream_chunk(h, b'data: {"model":"m","usage":{"prompt_tokens":3,"completion_tokens":4,"total_tokens":7}}\n\n', True)
            assert r["status"] == 0 and r["usage"] == (3, -1, -1, 4, 7, -1) and r["model"] == b"m"
            ctx.stream_close([h])
    except BaseException as e:   # noqa: BLE001 - reported by the main thread
        errors.append(e)


def test_two_contexts_one_device_two_threads():
    import aigw_b200 as A
    c1, c2 = A.Context(0), A.Context(0)
    try:
        errors = []
        ts = [threading.Thread(target=_check, args=(c1, _bodies(1, 3000), errors)), threading.Thread(target=_check, args=(c2, _bodies(2, 700), errors))]
        for t in ts: t.start()
        for t in ts: t.join()
        assert not errors, errors[0]
    finally:
        c1.close(); c2.close()


def test_one_context_per_device():
    import aigw_b200 as A
    try:
        c1 = A.Context(1)
    except Exception:   # noqa: BLE001
        pytest.skip("the box has one GPU")
    c0 = A.Context(0)
    try:
        errors = []
        ts = [threading.Thread(target=_check, args=(c0, _bodies(3, 2000), errors)), threading.Thread(target=_check, args=(c1, _bodies(4, 2000), errors))]
        for t in ts: t.start()
        for t in ts: t.join()
        assert not errors, errors[0]
    finally:
        c0.close(); c1.close()
