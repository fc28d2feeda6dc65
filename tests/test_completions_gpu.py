"""GPU parity for the response side of the legacy /v1/completions endpoint (internal/translator/openai_completions.go:80-203):
the stream usage scan as stream kind 'openai-completions' (one ResponseBody call per chunk) and the buffered branch
(aigw_completions_response_usage_*), against the reference's own answers (openai_completions_test.go:107-337, lifted into
tests/golden/completions_cases.json) and the oracle on corpora under random chunkings."""
import json
import os
import random

import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "completions_cases.json"), encoding="utf-8"))
OKS, INTERNAL, DECLINED = 0, 3, 4
UNCHANGED = 0


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def test_stream_reference_vectors_feed_by_feed(gw):
    for s in GOLD["streams"]:
        (h,) = gw.stream_open("openai-completions", b"req-model")
        for feed, exp, em in zip(s["feeds"], s["exp"], s["exp_model"]):
            r = gw.stream_chunk(h, feed.encode(), False)
            assert r["status"] == OKS and r["body_kind"] == UNCHANGED and r["body"] == b"", s["name"]
            assert list(r["usage"]) == exp, s["name"]
            assert r["model"].decode() == em, s["name"]
        assert r["carry_len"] == 0
        gw.stream_close([h])


def test_stream_one_byte_per_call(gw):
    """the three-call stream of the reference test, one byte per ResponseBody call: every call equals the oracle's, the usage shows up
    exactly once, the model is empty until the first chunk's line is complete (no fallback to the request model)"""
    data = "".join(GOLD["streams"][0]["feeds"]).encode()
    (h,) = gw.stream_open("openai-completions", b"req-model")
    orc = O.CompletionsSSEStream()
    seen = []
    for i in range(len(data)):
        r = gw.stream_chunk(h, data[i:i + 1], False)
        u = orc.feed(data[i:i + 1])
        assert r["status"] == OKS and r["usage"] == u.as_tuple(), i
        assert r["model"] == orc.model(), i
        assert r["carry_len"] == orc.buffered(), i
        if u.mask:
            seen.append(list(r["usage"]))
    assert seen == [GOLD["streams"][0]["exp"][-1]]
    gw.stream_close([h])


def _line(rng, i):
    roll = rng.random()
    if roll < 0.08:
        return rng.choice(["data: [DONE]", "", ": keep-alive", "event: completion", "data: invalid", "data:{\"model\":\"nospace\"}", "data: ", "data: null", "data: []",
                           "data: {\"model\":7}", "data: {\"choices\":{}}", "data: {\"created\":null}", "data: {\"created\":12.5,\"model\":\"frac\"}", "data: {\"created\":1e3}"])
    d = {"id": "cmpl-%d" % i, "object": "text_completion", "created": 1677649420, "model": rng.choice(["gpt-3.5-turbo-instruct", "davinci-002", ""]),
         "choices": [{"text": rng.choice(["Hello", " there!", "", "naïve 中文", "tab\tnew\nline"]), "index": 0, "logprobs": None, "finish_reason": rng.choice([None, "stop", "length"])}]}
    if rng.random() < 0.15:
        del d["model"]
    if rng.random() < 0.2:
        lp = {"tokens": ["He", "llo", None], "token_logprobs": [-0.25, None, -1e-3], "text_offset": [0, 2, 5],
              "top_logprobs": [{"He": -0.25, "he": -2.5}, None, {}]}
        bad = rng.random()
        if bad < 0.1: lp["top_logprobs"][0]["He"] = "x"          # map value of the wrong type: the chunk does not decode
        elif bad < 0.2: lp["text_offset"][1] = 2.5               # []int with a fraction
        elif bad < 0.3: lp["tokens"][0] = 7
        d["choices"][0]["logprobs"] = lp
    if rng.random() < 0.35:
        u = {"prompt_tokens": rng.randint(0, 5000), "completion_tokens": rng.randint(0, 3000)}
        u["total_tokens"] = u["prompt_tokens"] + u["completion_tokens"]
        if rng.random() < 0.4:
            u["prompt_tokens_details"] = rng.choice([{"cached_tokens": rng.randint(0, 99), "cache_creation_input_tokens": rng.randint(0, 9)}, {"cached_tokens": 3}, {}, None])
        if rng.random() < 0.4:
            u["completion_tokens_details"] = rng.choice([{"reasoning_tokens": rng.randint(0, 500)}, {"audio_tokens": 1}, None])
        if rng.random() < 0.08:
            u["total_tokens"] = rng.choice(["12", 1.5, 4294967299])   # wrong type / fraction fail the chunk; above 2^32 wraps like uint32(int)
        if rng.random() < 0.1:
            u = rng.choice([None, {}])
        d["usage"] = u
    sep = rng.choice([(",", ":"), (", ", ": ")])
    return "data: " + json.dumps(d, separators=sep, ensure_ascii=rng.random() < 0.5)


def _stream(rng, n_lines):
    out = []
    for i in range(n_lines):
        out.append(_line(rng, i))
        out.append("")
    return ("\n".join(out) + "\n").encode()


def test_stream_corpus_random_chunking_vs_oracle(gw):
    rng = random.Random(21)
    n = 192
    streams = []
    for s in range(n):
        data = _stream(rng, rng.randint(1, 14))
        cuts = sorted(rng.sample(range(1, len(data)), min(rng.randint(0, 9), len(data) - 1)))
        streams.append([data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])])
    hs = gw.stream_open("openai-completions", b"req-model", n=n)
    orc = [O.CompletionsSSEStream() for _ in range(n)]
    dead = set(); with_usage = 0
    for rnd in range(max(len(s) for s in streams)):
        idx = [i for i in range(n) if rnd < len(streams[i]) and i not in dead]
        got = gw.stream_chunks([hs[i] for i in idx], [streams[i][rnd] for i in idx], [False] * len(idx))
        for i, g in zip(idx, got):
            u = orc[i].feed(streams[i][rnd])
            if g["status"] == DECLINED:   # escaped model strings / member names are outside the GPU path; sticky
                dead.add(i); continue
            assert g["status"] == OKS
            assert g["usage"] == u.as_tuple(), (i, rnd, streams[i][rnd])
            assert g["model"] == orc[i].model(), (i, rnd)
            assert g["carry_len"] == orc[i].buffered()
            with_usage += 1 if u.mask else 0
    gw.stream_close(hs)
    assert len(dead) < n // 8 and with_usage > 100, (len(dead), with_usage)


def _unpack(r):
    m = int(r["mask"])
    return (int(r["input"]) if m & 1 else -1, int(r["cached"]) if m & 8 else -1, int(r["cache_creation"]) if m & 16 else -1,
            int(r["output"]) if m & 2 else -1, int(r["total"]) if m & 4 else -1, int(r["reasoning"]) if m & 32 else -1)


def test_buffered_reference_vectors(gw):
    import aigw_b200 as A
    bodies = [c["body"].encode() for c in GOLD["buffered"]]
    arena, offs, lens = A.pack_bodies(bodies)
    res, costs = gw.response_usage_host(arena, offs, lens, cost_types=(0, 3, 4), completions=True)
    raw = bytes(arena)
    for c, r, k in zip(GOLD["buffered"], res, costs):
        if c["exp_error"]:
            assert int(r["status"]) == INTERNAL, c["name"]
            continue
        assert int(r["status"]) == 0, c["name"]
        assert list(_unpack(r)) == c["exp"], c["name"]
        assert raw[int(r["model_off"]):int(r["model_off"]) + int(r["model_len"])].decode() == c["exp_model"], c["name"]
        assert [int(x) for x in k] == [max(c["exp"][0], 0), max(c["exp"][3], 0), max(c["exp"][4], 0)], c["name"]


def test_buffered_corpus_vs_oracle(gw):
    import aigw_b200 as A
    rng = random.Random(5)
    bodies = []
    for i in range(2500):
        b = _line(rng, i)
        b = b[6:].encode() if b.startswith("data: ") else b.encode()
        roll = rng.random()
        if roll < 0.04 and len(b) > 2: b = b[: rng.randint(1, len(b) - 1)]
        elif roll < 0.08: b = b + rng.choice([b" ", b"\n\t", b" x", b"{}"])        # json.Unmarshal: trailing white space only
        elif roll < 0.10: b = b.replace(b'"prompt_tokens":', b'"prompt_tokens":-', 1).replace(b'"prompt_tokens": ', b'"prompt_tokens": -', 1)
        bodies.append(b)
    bodies += [b"null", b"[]", b"{}", b'{"usage":null,"model":"m"}', b'{"usage":{}}', b"", b"  ", b'{"model":"a\\u0062"}', b'{"mod\\u0065l":"x"}',
               b'{"usage":{"prompt_tokens":-0}}', b'{"usage":{"prompt_tokens":-5,"completion_tokens":1}}']
    arena, offs, lens = A.pack_bodies(bodies)
    res, _ = gw.response_usage_host(arena, offs, lens, completions=True)
    raw = bytes(arena)
    n_ok = n_decl = n_err = 0
    for i, b in enumerate(bodies):
        ok, u, model = O.response_completions(b)
        r = res[i]
        if int(r["status"]) == DECLINED:
            n_decl += 1; continue
        if not ok:
            assert int(r["status"]) == INTERNAL, (i, b[:200]); n_err += 1
            continue
        assert int(r["status"]) == 0, (i, b[:200])
        assert _unpack(r) == u.as_tuple(), (i, b[:300], _unpack(r), u.as_tuple())
        assert raw[int(r["model_off"]):int(r["model_off"]) + int(r["model_len"])] == model, (i, b[:200])
        n_ok += 1
    print("completions buffered: ok", n_ok, "errors", n_err, "declined", n_decl)
    assert n_ok > 1700 and n_err > 100 and n_decl < 250
