"""GPU parity for the RESPONSE direction of /v1/messages served by an OpenAI-schema backend (T5; internal/translator/anthropic_openai.go:102-185,
openai_helper.go:263-766): stream kind 'messages-openai' (OpenAI SSE chunks -> Anthropic SSE events, one ResponseBody call per chunk) and
'messages-openai-buffered' (ChatCompletionResponse -> anthropic.MessagesResponse), against the five "anthropic-openai" data-plane goldens
(tests/data-plane/testupstream_test.go, byte for byte) and the oracle on corpora under random chunkings."""
import json
import os
import random

import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "testupstream_cases.json"), encoding="utf-8"))["cases"]
GOLD = [c for c in CASES if c["name"].startswith("anthropic-openai") and "responseBody" in c and "error" not in c["name"]]
OKS, INTERNAL, DECLINED = 0, 3, 4
UNCHANGED, BYTES, EMPTY = 0, 1, 2


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def sse_lines(body):
    """the fake upstream's line-by-line mode (tests/internal/testupstreamlib/server.go:303-320)"""
    return [b"data: " + l.encode() + b"\n\n" for l in body.split("\n") if l]


def chunkings(whole, blocks, seed=3):
    rng = random.Random(seed)
    cuts = sorted(rng.sample(range(1, len(whole)), min(17, len(whole) - 1)))
    return {"blocks": blocks, "whole": [whole], "bytes": [whole[i:i + 1] for i in range(len(whole))],
            "random": [whole[a:b] for a, b in zip([0] + cuts, cuts + [len(whole)])]}


@pytest.mark.parametrize("chunking", ["blocks", "whole", "bytes", "random"])
def test_stream_goldens(gw, chunking):
    streams = [c for c in GOLD if c.get("responseType") == "sse"]
    assert len(streams) == 2
    for c in streams:
        req_model = json.loads(c["requestBody"])["model"].encode()
        blocks = sse_lines(c["responseBody"])
        chunks = chunkings(b"".join(blocks), blocks)[chunking]
        (h,) = gw.stream_open("messages-openai", req_model)
        orc = O.MessagesOpenAIStream(req_model)
        out = b""
        for ch in chunks + [None]:
            eos = ch is None
            r = gw.stream_chunk(h, ch or b"", eos)
            s, o, u = orc.feed(ch or b"", eos)
            assert r["status"] == s == OKS
            assert r["body"] == o and r["usage"] == u.as_tuple() and r["model"] == orc.model(), c["name"]
            assert r["body_kind"] == (BYTES if o else EMPTY)
            assert r["carry_len"] == orc.buffered()
            out += r["body"]
        assert out.decode().strip() == c["expResponseBody"].strip(), c["name"]
        gw.stream_close([h])


def gen_chunk(rng, i, state):
    """one OpenAI chat-completion chunk line (without the 'data: ' prefix), or a non-chunk line"""
    roll = rng.random()
    if roll < 0.06:
        return rng.choice(["[DONE]", "not json", "null", "[]", "{}", '{"choices":{}}', '{"choices":[null]}', '{"choices":[{"delta":null}]}', '{"id":7}',
                           '{"choices":[{"index":"0"}]}', '{"created":null}', '{"choices":[{"delta":{"content":null},"finish_reason":null}]}'])
    d = {"id": rng.choice(["chatcmpl-%d" % state["sid"], ""]), "object": "chat.completion.chunk", "created": 1731618222, "model": rng.choice(["gpt-4o", "gpt-4o-mini", ""]), "choices": [], "usage": None}
    if rng.random() < 0.12:
        del d["model"]
    if roll > 0.93:
        d["usage"] = {"prompt_tokens": rng.randint(0, 9000), "completion_tokens": rng.randint(0, 4000), "total_tokens": 1}
        if rng.random() < 0.3:
            d["usage"]["prompt_tokens_details"] = {"cached_tokens": 5}
        if rng.random() < 0.15:
            d["choices"] = [{"index": 0, "delta": {"content": "late"}, "finish_reason": None}]   # not a usage-only chunk
        return json.dumps(d, separators=rng.choice([(",", ":"), (", ", ": ")]))
    delta = {}
    kind = rng.random()
    if kind < 0.45:
        delta["content"] = rng.choice(["Hi", " there!", "", "naïve 中文", "quote \" and \\ back", "line\nbreak\ttab", "<tag> & 'q'", "x" * rng.randint(1, 200)])
    elif kind < 0.8:
        tcs = []
        for _ in range(rng.randint(1, 2)):
            idx = rng.choice([0, 0, 1, 2, state["next_tool"]])
            tc = {"index": idx}
            if idx not in state["tools"] or rng.random() < 0.1:
                tc.update({"id": "call_%d" % idx, "type": "function", "function": {"name": rng.choice(["get_weather", "f", ""]), "arguments": ""}})
                state["tools"].add(idx); state["next_tool"] = max(state["next_tool"], idx + 1)
            else:
                tc["function"] = {"arguments": rng.choice(['{"location":', '"Paris"}', "", "{\n", ' "a": 1}', "\\u00e9"])}
            if rng.random() < 0.05:
                tc.pop("function", None)
            tcs.append(tc)
        delta["tool_calls"] = tcs
        if rng.random() < 0.2:
            delta["content"] = "text with tool"
    if rng.random() < 0.2:
        delta["role"] = "assistant"
    ch = {"index": 0, "delta": delta, "finish_reason": rng.choice([None, None, None, "stop", "length", "tool_calls", "content_filter", "function_call", ""])}
    if rng.random() < 0.05:
        del ch["delta"]
    d["choices"] = [ch] + ([{"index": 1, "delta": {"content": "second choice, ignored"}}] if rng.random() < 0.05 else [])
    return json.dumps(d, separators=rng.choice([(",", ":"), (", ", ": ")]), ensure_ascii=rng.random() < 0.3)


def gen_stream(rng, sid):
    state = {"sid": sid, "tools": set(), "next_tool": 0}
    lines = [gen_chunk(rng, i, state) for i in range(rng.randint(1, 16))]
    out = b""
    for l in lines:
        r = rng.random()
        if r < 0.85: out += b"data: " + l.encode() + b"\n\n"
        elif r < 0.9: out += b"event: x\ndata: " + l.encode() + b"\n: comment\n\n"
        elif r < 0.95: out += b"data: " + l.encode() + b"\ndata:   \n\n"          # the last NON-EMPTY data line wins
        else: out += b"data: {\"id\":\"shadowed\"}\ndata: " + l.encode() + b"  \n\n"
    if rng.random() < 0.3:
        out += b"data: [DONE]\n\n"
    if rng.random() < 0.15:
        out += b"data: " + gen_chunk(rng, 99, state).encode()                   # a last block without its blank line: processed at end of stream
    return out


def test_stream_corpus_random_chunking_vs_oracle(gw):
    rng = random.Random(31)
    n = 256
    streams = []
    for s in range(n):
        data = gen_stream(rng, s)
        cuts = sorted(rng.sample(range(1, len(data)), min(rng.randint(0, 12), len(data) - 1)))
        streams.append([data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])] + [b""])
    hs = gw.stream_open("messages-openai", b"req-model", n=n)
    orc = [O.MessagesOpenAIStream(b"req-model") for _ in range(n)]
    dead = set(); produced = 0
    for rnd in range(max(len(s) for s in streams)):
        idx = [i for i in range(n) if rnd < len(streams[i]) and i not in dead]
        eos = [rnd == len(streams[i]) - 1 for i in idx]
        got = gw.stream_chunks([hs[i] for i in idx], [streams[i][rnd] for i in idx], eos)
        for i, g, e in zip(idx, got, eos):
            s, o, u = orc[i].feed(streams[i][rnd], e)
            if g["status"] == DECLINED:   # escapes that would be re-spelled (\\u00e9, \\/), null tool calls: outside the GPU path; sticky
                dead.add(i); continue
            assert g["status"] == s == OKS
            assert g["body"] == o, (i, rnd, streams[i][rnd], g["body"], o)
            assert g["usage"] == u.as_tuple() and g["model"] == orc[i].model(), (i, rnd)
            assert g["body_kind"] == (BYTES if o else EMPTY) and g["carry_len"] == orc[i].buffered()
            produced += len(o)
    gw.stream_close(hs)
    print("messages-openai stream corpus: declined", len(dead), "of", n, "bytes", produced)
    assert len(dead) < n // 2 and produced > 50_000


def gen_response(rng, i):
    usage = {"prompt_tokens": rng.randint(0, 5000), "completion_tokens": rng.randint(0, 3000)}
    usage["total_tokens"] = usage["prompt_tokens"] + usage["completion_tokens"]
    msg = {"role": "assistant", "content": rng.choice(["Hello! How can I help you today?", "", None, "naïve 中文 \" \\ \n", "x" * rng.randint(1, 900)])}
    if rng.random() < 0.45:
        tcs = []
        for k in range(rng.randint(1, 3)):
            args = rng.choice(['{"location":"Paris"}', "{}", "", '{"a":1,"b":[true,null,{"c":"d"}],"c":-0.5}', '{"b":1,"a":2}', '{"location": "Paris"}', "oops", '{"a":1.50}', '{"a":"q\\"uote"}', '{"n":{"x":1,"y":{"z":[]}}}', "null", '{"a":1e3}'])
            tcs.append({"id": "call_%d" % k, "type": "function", "function": {"name": rng.choice(["get_weather", "f"]), "arguments": args}})
        msg["tool_calls"] = tcs
    d = {"id": rng.choice(["chatcmpl-%d" % i, ""]), "object": "chat.completion", "created": 1731618222, "model": rng.choice(["gpt-4o", ""]),
         "choices": [{"index": 0, "message": msg, "logprobs": None, "finish_reason": rng.choice(["stop", "length", "tool_calls", "content_filter", "", None])}], "usage": usage}
    r = rng.random()
    if r < 0.08: del d["usage"]
    elif r < 0.14: d["choices"] = []
    elif r < 0.18: d["choices"] = [None]
    elif r < 0.22: del d["choices"][0]["message"]
    b = json.dumps(d, separators=rng.choice([(",", ":"), (", ", ": ")]), ensure_ascii=rng.random() < 0.2).encode()
    r = rng.random()
    if r < 0.04: b = b[: rng.randint(1, len(b) - 1)]
    elif r < 0.07: b = b.replace(b'"index":0', b'"index":"0"', 1).replace(b'"index": 0', b'"index": "0"', 1)
    elif r < 0.10: b = b + b"  trailing bytes"        # json.Decoder reads one value
    return b


def test_buffered_goldens_and_corpus(gw):
    bodies = [(c["responseBody"].encode(), json.loads(c["requestBody"])["model"].encode(), c["expResponseBody"].encode()) for c in GOLD if c.get("responseType") != "sse"]
    assert len(bodies) == 3
    rng = random.Random(9)
    bodies += [(gen_response(rng, i), b"req-model", None) for i in range(1500)]
    bodies += [(b, b"req-model", None) for b in (b"null", b"{}", b"[]", b"", b'{"choices":[{"message":{"tool_calls":[null]}}]}', b'{"id":"a\\u0062"}', b'{"usage":{"prompt_tokens":-1}}')]
    n_ok = n_err = n_decl = 0
    for lo in range(0, len(bodies), 256):
        part = bodies[lo:lo + 256]
        hs = []
        for (_, m, _) in part:
            hs += gw.stream_open("messages-openai-buffered", m)
        # a body in two calls (the second carries eos) exercises the accumulation in the slot
        gw.stream_chunks(hs, [b[: len(b) // 2] for (b, _, _) in part], [False] * len(part))
        got = gw.stream_chunks(hs, [b[len(b) // 2:] for (b, _, _) in part], [True] * len(part))
        for (b, m, exp), g in zip(part, got):
            ok, o, u, model = O.messages_openai_response(b, m)
            if exp is not None:
                assert g["status"] == OKS and g["body"] == exp, (b, g["status"], g["body"])
            if g["status"] == DECLINED:
                n_decl += 1; continue
            if not ok:
                assert g["status"] == INTERNAL, (b, g["status"]); n_err += 1
                continue
            assert g["status"] == OKS and g["body_kind"] == BYTES, (b, g["status"])
            assert g["body"] == o, (b, g["body"], o)
            assert g["usage"] == u.as_tuple() and g["model"] == model, b
            n_ok += 1
        gw.stream_close(hs)
    print("messages-openai buffered: ok", n_ok, "errors", n_err, "declined", n_decl)
    assert n_ok > 800 and n_err > 60
