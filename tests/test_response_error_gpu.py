"""T0: Translator.ResponseError for the chat-completion translators of AWS Bedrock, GCP Vertex AI and GCP Anthropic on the GPU
(AIGW_SCHEMA_RESP_ERROR_*): the data-plane error goldens byte for byte, and a corpus of error bodies against the oracle."""
import json
import os
import random

import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "testupstream_cases.json"), encoding="utf-8"))["cases"]
CODE = {"http.StatusTooManyRequests": "429", "http.StatusBadRequest": "400"}
KINDS = ["aws-bedrock", "gcp-vertexai", "gcp-anthropicai"]


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def run(gw, kind, bodies, status="400", aws_type="", json_ct=True):
    from aigw_b200 import capi
    cfg = capi.Context.cfg("resp-error-" + kind, model_override=status, force=not json_ct)
    cfg.response_id = aws_type.encode() if aws_type else None
    return gw.chat_translate(cfg, bodies)


@pytest.mark.parametrize("kind", KINDS)
def test_goldens(gw, kind):
    c = next(c for c in CASES if c.get("backend") == kind and c["name"].endswith("/v1/chat/completions - error response"))
    g, = run(gw, kind, [c["responseBody"].encode()], status=CODE[str(c["expStatus"])], aws_type="ThrottledException" if kind == "aws-bedrock" else "")
    assert g["status"] == 0 and g["path"] == b"" and g["body"].decode() == c["expResponseBody"]


def err_body(r, kind):
    msg = r.choice(["rate limit exceeded", "Invalid request: missing required field", 'quo"ted \\ text', "line\nbreak\ttab", "", "ünïcode ✓", "x" * 300])
    k = r.random()
    if kind == "aws-bedrock":
        d = {"message": msg}
        if r.random() < 0.3: d["code"] = "429"
        if r.random() < 0.3: d["type"] = "ThrottlingException"
        if k < 0.06: d["message"] = 5
        elif k < 0.10: d = None
    elif kind == "gcp-anthropicai":
        d = {"type": "error", "error": {"type": r.choice(["invalid_request_error", "overloaded_error", ""]), "message": msg}}
        if r.random() < 0.3: d["request_id"] = "req_1"
        if k < 0.05: d["error"]["type"] = 7
        elif k < 0.08: d["error"] = "oops"
        elif k < 0.10: del d["error"]
    else:
        d = {"error": {"code": r.choice([400, 429, 503]), "message": msg, "status": r.choice(["INVALID_ARGUMENT", "RESOURCE_EXHAUSTED", ""])}}
        if k < 0.05: d["error"]["details"] = [{"@type": "x"}]
        elif k < 0.10: d["error"]["code"] = "400"            # not the expected type: the raw body becomes the message
        elif k < 0.13: d["error"]["code"] = 4.5
        elif k < 0.16: d = [1, 2]
    body = json.dumps(d, separators=(",", ":") if r.random() < 0.6 else None, ensure_ascii=r.random() < 0.4)
    if 0.16 <= k < 0.22:
        body = r.choice(["upstream connect error or disconnect/reset before headers", "<html><body>503 Service Unavailable</body></html>", "plain \"text\"\nsecond line", body[: len(body) // 2]])
    return body.encode()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("json_ct", [True, False])
def test_corpus_vs_oracle(gw, kind, json_ct):
    r = random.Random(31 + len(kind))
    bodies = [err_body(r, kind) for _ in range(700)] + [b"null", b"{}", b"backend timeout", b' {"message":"spaced"} ', b'{"message":"a"} trailing']
    got = run(gw, kind, bodies, status="503", aws_type="ServiceUnavailableException", json_ct=json_ct)
    ok = decl = 0
    for b, g in zip(bodies, got):
        st, out = O.response_error(kind, b, "503", "ServiceUnavailableException", json_content_type=json_ct)
        if g["status"] == 4:
            decl += 1
            continue
        assert g["status"] == st, (b[:200], g["status"], g["reason"], st)
        if st == 0:
            assert g["body"] == out, (b[:200], g["body"], out)
            ok += 1
    print(kind, json_ct, "ok", ok, "declined", decl)
    assert ok > 450
    # the fused small-batch kernel gives the same answers as the pipeline
    try:
        gw.chat_set_small_batch(0)
        a = run(gw, kind, bodies[:300], status="503", aws_type="ServiceUnavailableException", json_ct=json_ct)
        assert a == got[:300]
    finally:
        gw.chat_set_small_batch(-1)
