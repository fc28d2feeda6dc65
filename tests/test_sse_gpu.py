"""GPU parity tests for the OpenAI SSE usage scan (S1 + C1): CUDA kernel through the C ABI vs the oracle's
chunk-by-chunk replay of extractUsageFromBufferEvent + Override."""
import ctypes as C
import json
import random
import os

import numpy as np
import pytest

import _oracle as O
import _workload as W

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import aigw_b200 as A
    c = A.Context(0)
    yield c
    c.close()


def _oracle_stream(chunks):
    s = O.SSEStream()
    acc = [-1] * 6
    for ch in chunks:
        u = s.feed(ch).as_tuple()
        acc = [n if n >= 0 else a for a, n in zip(acc, u)]
    return tuple(acc), s.model()


def _gpu_streams(ctx, streams):
    """streams: list of list-of-chunks(bytes)."""
    flat = b"".join(b"".join(s) for s in streams)
    buf = np.frombuffer(flat + b"\0" * 64, dtype=np.uint8).copy()
    coff = [0]
    cfirst = [0]
    for s in streams:
        for ch in s:
            coff.append(coff[-1] + len(ch))
        cfirst.append(len(coff) - 1)
    res, _ = ctx.sse_usage_host(buf, np.array(coff, dtype=np.uint64), np.array(cfirst, dtype=np.uint32))
    out = []
    for r in res:
        m = int(r["mask"])
        tup = (int(r["input"]) if m & 1 else -1, int(r["cached"]) if m & 8 else -1, int(r["cache_creation"]) if m & 16 else -1,
               int(r["output"]) if m & 2 else -1, int(r["total"]) if m & 4 else -1, int(r["reasoning"]) if m & 32 else -1)
        model = flat[int(r["model_off"]):int(r["model_off"]) + int(r["model_len"])]
        out.append((int(r["status"]), tup, model))
    return out


def test_reference_usage_vectors(ctx):
    cases = json.load(open(os.path.join(G, "sse_usage_cases.json")))["cases"]
    streams = [[f.encode() for f in c["feeds"]] for c in cases]
    got = _gpu_streams(ctx, streams)
    for c, g in zip(cases, got):
        acc = [-1] * 6
        for e in c["exp"]:
            acc = [n if n >= 0 else a for a, n in zip(acc, e)]
        assert g[0] == 0 and list(g[1]) == acc, c["name"]


def test_reference_streaming_golden(ctx):
    c = next(c for c in json.load(open(os.path.join(G, "testupstream_cases.json"), encoding="utf-8"))["cases"] if c["name"] == "openai - /v1/chat/completions - streaming")
    body = c["expResponseBody"].encode()
    streams = [[body], [body[i:i + 1] for i in range(len(body))], [body[i:i + 7] for i in range(0, len(body), 7)]]
    for st, tup, model in _gpu_streams(ctx, streams):
        assert st == 0 and tup == (13, 0, 0, 12, 25, 0) and model == b"gpt-4o-mini-2024-07-18"


def test_workload_parity(ctx):
    n = 400
    buf, coff, cfirst = W.sse_corpus(4, 0, n)
    res, _ = ctx.sse_usage_host(buf, coff, cfirst)
    raw = bytes(buf)
    for s in range(n):
        chunks = [raw[int(coff[c]):int(coff[c + 1])] for c in range(int(cfirst[s]), int(cfirst[s + 1]))]
        exp, model = _oracle_stream(chunks)
        r = res[s]; m = int(r["mask"])
        got = (int(r["input"]) if m & 1 else -1, int(r["cached"]) if m & 8 else -1, int(r["cache_creation"]) if m & 16 else -1,
               int(r["output"]) if m & 2 else -1, int(r["total"]) if m & 4 else -1, int(r["reasoning"]) if m & 32 else -1)
        assert int(r["status"]) == 0 and got == exp, s
        assert raw[int(r["model_off"]):int(r["model_off"]) + int(r["model_len"])] == model


def test_adversarial_lines(ctx):
    lines = [b'data: {"usage":{"prompt_tokens":1,"completion_tokens":2,"total_tokens":3}}\n', b"data: [DONE]\n", b": comment\n", b"event: x\n", b"data:{\"usage\":{\"total_tokens\":9}}\n",
             b'data: {"usage":{"total_tokens":"7"}}\n', b'data: {"usage":{"total_tokens":7.0}}\n', b'data: {"usage":{"total_tokens":4294967297}}\n', b'data: {"usage":{"total_tokens":99999999999999999999}}\n',
             b'data: {"model":"m1","usage":null}\n', b'data: {"model":"","choices":[{"index":0,"delta":{"content":null}}]}\n', b'data: {"model":5}\n', b'data: {"created":1.5,"model":"m2"}\n',
             b'data: {"created":"1","model":"bad"}\n', b'data: {"created":null,"model":"bad2"}\n', b'data: {"created":1e3,"model":"bad3"}\n', b'data: {"choices":[{"index":"0"}],"model":"bad4"}\n',
             b'data: {"choices":[{"index":0,"delta":{"tool_calls":[{"index":0,"id":null,"function":{"arguments":"{}","name":"f"},"type":"function"}]}}],"model":"m3"}\r\n',
             b'data: {"usage":{"prompt_tokens":5,"prompt_tokens_details":{"cached_tokens":2}},"model":"m4"} \n', b'data: {"usage":{"prompt_tokens":6,"completion_tokens_details":null}}\n',
             b'data:  {"usage":{"total_tokens":11}}\n', b'data: {"usage":{"total_tokens":12}} garbage\n', b'data: {"usage":{"total_tokens":13},}\n', b'data: {"a":{"b":[1,2,{"c":"\\u00e9\\n"}]},"usage":{"total_tokens":14}}\n',
             b'data: {"choices":[{"logprobs":{"content":[{"token":"a","bytes":[1,2],"logprob":-0.5,"top_logprobs":[]}]}}],"model":"m5"}\n',
             b'data: {"choices":[{"logprobs":{"content":[{"token":"a","bytes":[1.5]}]}}],"model":"bad5"}\n', b'data: null\n', b'data: []\n', b'data: {"usage":{"total_tokens":15}}']
    import random
    r = random.Random(5)
    streams = []
    for k in range(300):
        pick = [r.choice(lines[:-1]) for _ in range(r.randint(0, 12))]
        if r.random() < 0.3:
            pick.append(lines[-1])  # unterminated tail
        body = b"".join(pick)
        cuts = sorted(r.sample(range(len(body) + 1), min(len(body) + 1, r.randint(0, 6)))) if body else []
        chunks = [body[a:b] for a, b in zip([0] + cuts, cuts + [len(body)])] or [b""]
        streams.append(chunks)
    got = _gpu_streams(ctx, streams)
    for chunks, (st, tup, model) in zip(streams, got):
        exp, emodel = _oracle_stream(chunks)
        if st == 0:
            assert tup == exp and model == emodel, (b"".join(chunks), tup, exp, model, emodel)


def test_line_filter_is_exact(ctx):
    """the kernel walks only lines that can matter (a `"usage"` member, a \\u escape, the last two data lines of a tile):
    streams built to defeat that shortcut must still give the reference's usage and response model"""
    rng = random.Random(7)
    def chunk(model, content="x", usage=None, extra=""):
        d = {"id": "c", "object": "chat.completion.chunk", "created": 1, "model": model, "choices": [{"index": 0, "delta": {"content": content}}]}
        if usage is not None: d["usage"] = usage
        return ("data: " + json.dumps(d, separators=(",", ":")) + extra + "\n\n").encode()
    U = lambda a, b: {"prompt_tokens": a, "completion_tokens": b, "total_tokens": a + b}
    streams = []
    # usage line early with model A, later valid lines with model B, then an invalid tail
    streams.append([chunk("A", usage=U(1, 2)), chunk("B"), chunk("B"), b"data: {not json}\n\n", b"data: [DONE]\n\n"])
    # last data lines valid but with an empty model; an earlier line carries it
    streams.append([chunk("M1"), chunk(""), chunk(""), b"data: [DONE]\n\n"])
    # the word usage inside content (escaped quotes), a key spelled with \u escapes, a usage member that is null
    streams.append([chunk("m", content='he said "usage" twice'), b'data: {"model":"m","us\\u0061ge":{"prompt_tokens":9,"completion_tokens":1,"total_tokens":10}}\n\n', chunk("m", usage=None), b"data: [DONE]\n\n"])
    # two usage lines, the later one wins field by field; model switches after the last usage line
    streams.append([chunk("a", usage=U(5, 6)), chunk("a", usage={"prompt_tokens": 7, "completion_tokens": 8, "total_tokens": 15, "prompt_tokens_details": {"cached_tokens": 3}}), chunk("zz"), b"data: [DONE]\n\n"])
    # many tiles: long streams whose only valid model-bearing line sits far from the tile ends
    long = [chunk("first")] + [b"data: {broken\n\n"] * 400 + [chunk("mid", usage=U(2, 3))] + [b": comment\n\n"] * 300 + [b"data: {broken\n\n"] * 3
    streams.append(long)
    # random mixtures
    for _ in range(200):
        s = []
        for _ in range(rng.randint(1, 60)):
            r = rng.random()
            if r < 0.6: s.append(chunk(rng.choice(["a", "b", "", "gpt-4o"]), content="t" * rng.randint(0, 200)))
            elif r < 0.7: s.append(chunk(rng.choice(["a", "b"]), usage=U(rng.randint(0, 99), rng.randint(0, 99))))
            elif r < 0.8: s.append(b"data: {oops\n\n")
            elif r < 0.9: s.append(b"event: ping\n\n")
            else: s.append(chunk("esc", content="caf\\u00e9"))
        streams.append(s)
    got = _gpu_streams(ctx, streams)
    for s, g in zip(streams, got):
        exp_u, exp_m = _oracle_stream(s)
        if g[0] == 4:
            continue
        assert g[0] == 0 and g[1] == exp_u and g[2] == exp_m, (b"".join(s)[:300], g, exp_u, exp_m)
    assert sum(1 for g in got if g[0] == 0) >= len(streams) - 60
