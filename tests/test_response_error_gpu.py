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


# ---------------------------------------------------------------- /v1/messages served by an OpenAI backend: openai.Error -> anthropic.ErrorResponse
def test_messages_openai_error_golden_and_corpus(gw):
    """anthropicToOpenAIV1ChatCompletionTranslator.ResponseError (internal/translator/anthropic_openai.go:187-253): the data-plane golden
    "OpenAI JSON error translated to Anthropic error" byte for byte; JSON and raw-body branches against the oracle through both launch paths."""
    c = next(c for c in CASES if c["name"].endswith("OpenAI JSON error translated to Anthropic error"))
    g, = run(gw, "messages-openai", [c["responseBody"].encode()], status="400")
    assert g["status"] == 0 and g["body"].decode() == c["expResponseBody"]
    r = random.Random(17)
    bodies = []
    for i in range(1500):
        msg = r.choice(["Model not found", "Rate limit reached for requests", 'quo"ted \\ text', "line\nbreak\ttab", "", "ünïcode ✓", "x" * 300])
        e = {"type": r.choice(["invalid_request_error", "rate_limit_error", ""]), "message": msg, "code": r.choice(["model_not_found", None]), "param": r.choice([None, "model"])}
        d = {"error": e}
        k = r.random()
        if k < 0.05: e["message"] = 5
        elif k < 0.09: e["param"] = 7
        elif k < 0.12: d["event_id"] = r.choice(["ev_1", 3])
        elif k < 0.15: d["error"] = r.choice([None, "oops", []])
        elif k < 0.18: d = r.choice([None, [1], "str", 5])
        elif k < 0.21: e["event_id"] = r.choice([None, "e", True]); del e["code"]
        elif k < 0.24: e["code"] = 404
        elif k < 0.27: d["type"] = r.choice(["error", 1])
        b = json.dumps(d, separators=r.choice([(",", ":"), (", ", ": ")]), ensure_ascii=r.random() < 0.3).encode()
        if r.random() < 0.04: b = b[: r.randint(1, max(1, len(b) - 1))]
        bodies.append(b)
    raw = [b"upstream connect error or disconnect/reset before headers", b"<html>502 Bad Gateway</html>", b'plain "quoted" \\ text\n', b"tab\there", b"{not json"]
    for small in (0, 1 << 20):
        gw.chat_set_small_batch(small)
        try:
            got = run(gw, "messages-openai", bodies, status="404")
            n_ok = n_err = 0
            for b, g in zip(bodies, got):
                st, o = O.response_error("messages-openai", b, "404", "", True)
                if g["status"] == 4: continue
                assert g["status"] == st, (b, g["status"], g["reason"], st)
                if st == 0:
                    assert g["body"] == o, (b, g["body"], o); n_ok += 1
                else:
                    n_err += 1
            assert n_ok > 900 and n_err > 100, (n_ok, n_err)
            for status in ("400", "401", "403", "404", "413", "429", "500", "503", "529", "418", "50"):
                for b, g in zip(raw, run(gw, "messages-openai", raw, status=status, json_ct=False)):
                    st, o = O.response_error("messages-openai", b, status, "", False)
                    assert g["status"] == st == 0 and g["body"] == o, (status, b, g["body"], o)
        finally:
            gw.chat_set_small_batch(-1)


def test_messages_bedrock_error(gw):
    """anthropicToAWSBedrockTranslator.ResponseError (internal/translator/anthropic_awsbedrock.go:738-794): the reference test's cases
    (anthropic_awsbedrock_test.go:563-638: JSON body at 404, text body at 503, the status table) and a corpus against the oracle."""
    g, = run(gw, "messages-aws-bedrock", [b'{"message":"Model not found"}'], status="404")
    assert g["status"] == 0 and json.loads(g["body"]) == {"type": "error", "error": {"type": "not_found_error", "message": "Model not found"}, "request_id": ""}
    g, = run(gw, "messages-aws-bedrock", [b"Service Unavailable"], status="503", json_ct=False)
    assert g["status"] == 0 and json.loads(g["body"])["error"] == {"type": "service_unavailable_error", "message": "Service Unavailable"}
    for status, typ in (("400", "invalid_request_error"), ("401", "authentication_error"), ("403", "permission_error"), ("404", "not_found_error"), ("429", "rate_limit_error"),
                        ("500", "internal_server_error"), ("503", "service_unavailable_error"), ("502", "internal_server_error"), ("413", "internal_server_error"), ("529", "internal_server_error")):
        g, = run(gw, "messages-aws-bedrock", [b"error"], status=status, json_ct=False)
        assert g["status"] == 0 and json.loads(g["body"])["error"]["type"] == typ, status
    r = random.Random(23)
    bodies = [err_body(r, "aws-bedrock") for _ in range(1200)] + [b"null", b"[]", b"{}", b'{"message":null}', b"{oops"]
    for small in (0, 1 << 20):
        gw.chat_set_small_batch(small)
        try:
            for status in ("429", "404"):
                n_ok = n_err = 0
                for b, g in zip(bodies, run(gw, "messages-aws-bedrock", bodies, status=status)):
                    st, o = O.response_error("messages-aws-bedrock", b, status, "", True)
                    if g["status"] == 4: continue
                    assert g["status"] == st, (b, g["status"], g["reason"], st)
                    if st == 0:
                        assert g["body"] == o, (b, g["body"], o); n_ok += 1
                    else:
                        n_err += 1
                assert n_ok > 700 and n_err > 30, (n_ok, n_err)
        finally:
            gw.chat_set_small_batch(-1)


def test_openai_passthrough_error(gw):
    """convertErrorOpenAIToOpenAIError (internal/translator/openai_openai.go:94-120; the OpenAI / Azure chat, embeddings and completions
    translators): a non-JSON upstream error becomes an OpenAIBackendError — the data-plane golden "non json upstream error mapped to OpenAI"
    (JSON-equal, as the reference compares it) and raw bodies against the oracle; a JSON body is forwarded untouched (nothing to compute)."""
    c = next(c for c in CASES if c["name"].endswith("non json upstream error mapped to OpenAI"))
    g, = run(gw, "openai", [c["responseBody"].encode()], status="503", json_ct=False)
    assert g["status"] == 0 and json.loads(g["body"]) == json.loads(c["expResponseBody"])
    raw = [b"backend timeout", b"upstream connect error or disconnect/reset before headers", b"<html>502 Bad Gateway</html>", b'plain "quoted" \\ text\n', b"tab\there", b"{not json", b'{"looks":"like json"}']
    for small in (0, 1 << 20):
        gw.chat_set_small_batch(small)
        try:
            for status in ("400", "503", "504"):
                for b, g in zip(raw, run(gw, "openai", raw, status=status, json_ct=False)):
                    st, o = O.response_error("openai", b, status, "", False)
                    assert g["status"] == st == 0 and g["body"] == o, (status, b, g["body"], o)
            assert all(g["status"] == 4 for g in run(gw, "openai", raw, status="400", json_ct=True))
        finally:
            gw.chat_set_small_batch(-1)
