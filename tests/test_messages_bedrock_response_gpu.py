"""T5: the buffered AWS Bedrock Converse response of a /v1/messages call -> anthropic.MessagesResponse on the GPU
(anthropicToAWSBedrockTranslator.ResponseBody, internal/translator/anthropic_awsbedrock.go:429-510), against the data-plane golden
"aws-bedrock - /anthropic/v1/messages" and the oracle (oracle/bedrock_response.hpp, to_anthropic)."""
import json
import os

import numpy as np
import pytest

import _oracle as O
import _workload as W

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "testupstream_cases.json"), encoding="utf-8"))["cases"]
GOLD = [c for c in CASES if c["name"] == "aws-bedrock - /anthropic/v1/messages"][0]


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def run(gw, bodies, model="m", rid="r"):
    from aigw_b200 import capi
    cfg = capi.Context.cfg("resp-messages-aws-bedrock", model_override=model)
    cfg.response_id = rid.encode() if rid else None
    arena, offs, lens = capi.pack_bodies(bodies)
    res, out, _ = gw.chat_translate_host(cfg, arena, offs, lens)
    r = []
    for i in range(len(bodies)):
        x = res[i]
        o = int(x["out_off"]); pl = int(x["path_len"]); bl = int(x["body_len"])
        if x["status"] == 0:
            assert pl == 32
            u = tuple(int(v) for v in np.frombuffer(bytes(out[o:o + 32]), dtype="<u4")[:7])
            r.append((0, bytes(out[o + pl:o + pl + bl]), u))
        else:
            r.append((int(x["status"]), b"", None))
    return r


def check(gw, bodies, model="m", rid="r"):
    got = run(gw, bodies, model, rid)
    ok = decl = err = 0
    for b, (st, body, u) in zip(bodies, got):
        est, eout, eu = O.bedrock_response_anthropic(b, model.encode(), rid.encode())
        if st == 4:
            decl += 1
            continue
        assert st == est, (b[:300], st, est)
        if st == 0:
            assert body == eout, (b[:300], body, eout)
            assert u == (eu.input, eu.output, eu.total, eu.cached, eu.cache_creation, eu.reasoning, eu.mask)
            ok += 1
        else:
            err += 1
    return ok, decl, err


def test_golden(gw):
    (st, body, u), = run(gw, [GOLD["responseBody"].encode()], model="anthropic.claude-3-sonnet-20240229-v1:0", rid="bedrock-msg-123")
    assert st == 0
    assert body.decode() == GOLD["expResponseBody"]     # the reference compares with JSONEq; the text is the same
    assert u[:3] == (10, 20, 30)


def test_parity_corpus(gw):
    rng = np.random.default_rng(19)
    bodies = [W.bedrock_response_body(rng, ["plain", "tools", "reasoning", "cache"][i % 4]) for i in range(2000)]
    ok, decl, err = check(gw, bodies, model="the-request-model", rid="req-1")
    print("ok", ok, "declined", decl, "errors", err)
    assert ok > 1900


def test_odd_bodies(gw):
    odd = list(W.BEDROCK_RESPONSE_ODD) + [
        b'{"output":{"message":{"content":[{"text":"a"},{"toolUse":{"name":"n","input":{"b":1,"a":"x"},"toolUseId":"t"}},{"reasoningContent":{"reasoningText":{"text":"th","signature":"s"}}}],"role":"assistant"}},"stopReason":"tool_use","usage":{"inputTokens":1,"outputTokens":2,"totalTokens":3,"cacheReadInputTokens":4}}',
        b'{"output":{"message":{"content":[],"role":"assistant"}},"stopReason":"stop_sequence"}', b'{"output":{"message":{"content":[{"text":"x"}]}},"stopReason":"content_filtered"}', b'{"output":{}}', b'{}', b'null', b'[]',
        b'{"output":{"message":{"content":[{"toolUse":{"name":"n","toolUseId":"t"}}],"role":"assistant"}}}', b'{"output":{"message":{"content":[{"reasoningContent":{"redactedContent":"aGk="}}]}}}',
        b'{"output":{"message":{"content":[{"text":"both","toolUse":{"name":"n","input":{},"toolUseId":"t"}}]}}}', b'{"output":{"message":{"content":[{"reasoningContent":{}}]}},"usage":{"inputTokens":5,"outputTokens":6,"totalTokens":11,"cacheWriteInputTokens":2}}']
    ok, decl, err = check(gw, odd, model="", rid="")
    print("ok", ok, "declined", decl, "errors", err, "of", len(odd))
    assert ok >= 6
