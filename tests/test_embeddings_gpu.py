"""P3 + T1 + T6 parity: /v1/embeddings parse, OpenAI / Azure passthrough and Vertex predict translate on the GPU vs oracle and goldens."""
import json
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "testupstream_cases.json"), encoding="utf-8"))["cases"]
EMB_CASES = [c for c in CASES if c.get("path") == "/v1/embeddings" and "requestBody" in c]
GPU_SCHEMA = {"openai": "emb-openai", "azure-openai": "emb-azure-openai", "gcp-vertexai": "emb-gcp-vertexai", "aws-bedrock": "emb-aws-bedrock"}


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def run(gw, schema, bodies, model_override=None, prefix=None, api_version=None, force=False):
    from aigw_b200 import capi
    cfg = capi.Context.cfg(GPU_SCHEMA[schema], model_override=model_override, prefix=prefix, api_version=api_version, force=force)
    return gw.chat_translate(cfg, bodies)


def check(gw, schema, bodies, model_override="", prefix="v1", api_version="2025-01-01-preview", force=False, allow_decline=False):
    got = run(gw, schema, bodies, model_override or None, prefix if schema == "openai" else None, api_version if schema == "azure-openai" else None, force)
    n_ok = 0
    for b, g in zip(bodies, got):
        t = O.embeddings_translate(schema, b, model_override=model_override, prefix=api_version if schema == "azure-openai" else prefix, force=force)
        if g["status"] == 4 and t.status != 4:
            assert allow_decline, (b[:200], g["reason"], t.status)
            continue
        assert g["status"] == t.status, (b[:300], g, t.status, t.err)
        if t.status == 0:
            assert g["path"].decode() == t.path, (b[:200], g["path"], t.path)
            assert g["body_kind"] == t.body_kind, (b[:200], g["body_kind"], t.body_kind)
            if t.body_kind == 1:
                assert g["body"] == t.body, (b[:300], g["body"], t.body)
            assert g["model"] == t.model
        n_ok += 1
    return n_ok


@pytest.mark.parametrize("c", [c for c in EMB_CASES if "expRequestBody" in c], ids=lambda c: c["name"])
def test_vertex_goldens(gw, c):
    """tests/data-plane/testupstream_test.go:782-852 (bytes.Equal at the upstream)."""
    g, = run(gw, "gcp-vertexai", [c["requestBody"].encode()])
    assert g["status"] == 0 and g["body"].decode() == c["expRequestBody"]
    assert c["expPath"].endswith("/" + g["path"].decode())


@pytest.mark.parametrize("c", [c for c in EMB_CASES if "expRequestBody" not in c and c["backend"] in ("openai", "azure-openai")], ids=lambda c: c["name"])
def test_passthrough_goldens(gw, c):
    g, = run(gw, c["backend"], [c["requestBody"].encode()], prefix="v1", api_version="2025-01-01-preview")
    assert g["status"] == 0 and g["body_kind"] == 0
    assert g["path"].decode().split("?")[0] == c["expPath"]


def emb_body(rng):
    words = ["reset", "password", "café", "日本語", "line\nbreak", "quo\"te", "tab\there", "back\\slash", "vector", "search", "embedding", "x" * int(rng.integers(1, 60))]
    txt = lambda: " ".join(words[int(i)] for i in rng.integers(0, len(words), int(rng.integers(1, 12))))
    tasks = ["RETRIEVAL_QUERY", "RETRIEVAL_DOCUMENT", "SEMANTIC_SIMILARITY", "CLUSTERING", ""]
    item = lambda: {k: v for k, v in {"content": txt() if rng.integers(0, 3) else [txt() for _ in range(int(rng.integers(1, 4)))],
                                      "task_type": tasks[int(rng.integers(0, 5))] if rng.integers(0, 2) else None, "title": txt() if rng.integers(0, 2) else None}.items() if v is not None}
    k = int(rng.integers(0, 7))
    inp = [txt(), [txt() for _ in range(int(rng.integers(0, 40)))], item(), [item() for _ in range(int(rng.integers(1, 12)))],
           [int(x) for x in rng.integers(0, 100000, int(rng.integers(1, 50)))], [[int(x) for x in rng.integers(0, 100000, 5)] for _ in range(3)], ""][k]
    d = {"model": ["text-embedding-3-small", "text-embedding-004", "m"][int(rng.integers(0, 3))], "input": inp}
    if rng.integers(0, 3) == 0: d["dimensions"] = int(rng.integers(0, 1024))
    if rng.integers(0, 4) == 0: d["encoding_format"] = "float"
    if rng.integers(0, 4) == 0: d["user"] = "u-1"
    if rng.integers(0, 5) == 0: d["auto_truncate"] = bool(rng.integers(0, 2))
    if rng.integers(0, 5) == 0: d["task_type"] = tasks[int(rng.integers(0, 5))]
    if rng.integers(0, 6) == 0: d = dict(reversed(list(d.items())))
    return json.dumps(d, ensure_ascii=False, separators=(",", ":") if rng.integers(0, 4) else (", ", ": ")).encode()


ODD = [
    b'{"model":"m","input":null}', b'{"model":"m","input":5}', b'{"model":"m","input":[null]}', b'{"model":"m","input":["a",5]}', b'{"model":"m","input":{"content":""}}',
    b'{"model":"m","input":{"task_type":"X"}}', b'{"model":"m","input":[{"content":"a"},{"content":[]}]}', b'{"model":5,"input":"a"}', b'{"model":"m","input":"a","dimensions":1.5}',
    b'{"model":"m","input":[1,2,3]}', b'{"model":"m","input":[[1,2],[3]]}', b'{"model":"m","input":[1,"a"]}', b'{"model":"m"}', b'{"model":"m","input":[]}', b'{"model":"m","input":["a",null]}',
    b'{"model":"m","input":{"content":["a","b"],"task_type":"RETRIEVAL_DOCUMENT","title":"T"}}', b'{"model":"m","input":{"content":"a","task_type":"RETRIEVAL_QUERY","title":"dropped"}}',
    b'{"model":"m","input":[{"content":"a","task_type":"RETRIEVAL_DOCUMENT","title":"T"}],"task_type":"CLUSTERING","auto_truncate":true,"dimensions":8}',
    b'{"model":"m","input":[{"content":null}]}', b'{"model":"m","input":[{"content":"a","title":5}]}', b'{"model":null,"input":"a"}', b'{"input":"a"}', b'{}', b'null', b'[]', b'{"model":"m","input":"a"',
    b'{"model":"m","input":"a","auto_truncate":"yes"}', b'{"model":"m","input":"a","dimensions":0}', b'{"model":"m","input":"a","dimensions":-4}', b'{"model":"m","input":[[1,null],null]}',
    b'{"model":"m","input":{"content":["a",null]},"user":null,"encoding_format":null,"task_type":null,"auto_truncate":null}', b'{"model":"m","input":"a","user":5}',
    b' { "model" : "m" , "input" : [ "a" , "b" ] } ',
]


@pytest.mark.parametrize("schema", ["openai", "azure-openai", "gcp-vertexai", "aws-bedrock"])
def test_parity_odd_bodies(gw, schema):
    assert check(gw, schema, ODD) >= 10
    assert check(gw, schema, ODD, model_override="override-model") >= 10
    if schema not in ("gcp-vertexai", "aws-bedrock"):
        assert check(gw, schema, ODD, force=True) >= 10


@pytest.mark.parametrize("schema", ["openai", "azure-openai", "gcp-vertexai", "aws-bedrock"])
def test_parity_corpus(gw, schema):
    rng = np.random.default_rng(31)
    bodies = [emb_body(rng) for _ in range(3000)]
    assert check(gw, schema, bodies) == len(bodies)
    assert check(gw, schema, bodies[:500], model_override="text-embedding-005") == 500


def test_large_inputs(gw):
    """hundreds of inputs per request (tens of KiB); bodies above 64 KiB are declined, never truncated"""
    rng = np.random.default_rng(2)
    bodies = [json.dumps({"model": "text-embedding-3-small", "input": ["w%d " % j + "x" * 56 for j in range(n)]}, separators=(",", ":")).encode() for n in (10, 100, 400, 900)]
    assert check(gw, "gcp-vertexai", bodies, allow_decline=True) >= 3
    assert check(gw, "openai", bodies, model_override="m2", allow_decline=True) >= 3
    big = json.dumps({"model": "text-embedding-3-small", "input": ["x" * 64 for _ in range(1024)]}, separators=(",", ":")).encode()
    assert len(big) > 65536
    g, = run(gw, "gcp-vertexai", [big])
    assert g["status"] == 4


# the reference's own table for the Titan translator (internal/translator/openai_awsbedrock_embeddings_test.go:25-130): request as the
# test marshals it, expected :path, substrings the body must / must not contain, or the 422
TITAN = [
    (b'{"input":"hello world","model":"amazon.titan-embed-text-v1:2"}', "", "/model/amazon.titan-embed-text-v1:2/invoke", [b'"inputText":"hello world"'], []),
    (b'{"input":"hello world","model":"amazon.titan-embed-text-v2:0"}', "", "/model/amazon.titan-embed-text-v2:0/invoke", [b'"inputText":"hello world"'], []),
    (b'{"input":["hello world"],"model":"amazon.titan-embed-text-v2:0"}', "", "/model/amazon.titan-embed-text-v2:0/invoke", [b'"inputText":"hello world"'], []),
    (b'{"input":["first","second"],"model":"amazon.titan-embed-text-v2:0"}', "", None, [], []),
    (b'{"input":[],"model":"amazon.titan-embed-text-v2:0"}', "", None, [], []),
    (b'{"input":"test","model":"amazon.titan-embed-text-v1:2"}', "", "/model/amazon.titan-embed-text-v1:2/invoke", [b'"inputText":"test"'], [b'"dimensions"', b'"normalize"', b'"embeddingTypes"']),
    (b'{"input":"test","model":"amazon.titan-embed-text-v2:0","dimensions":256}', "", "/model/amazon.titan-embed-text-v2:0/invoke", [b'"inputText":"test"', b'"dimensions":256'], []),
    (b'{"input":"test","model":"amazon.titan-embed-text-v2:0"}', "amazon.titan-embed-text-v1:2", "/model/amazon.titan-embed-text-v1:2/invoke", [b'"inputText":"test"'], []),
    (b'{"input":"escape test","model":"my model v1"}', "", "/model/my%20model%20v1/invoke", [b'"inputText":"escape test"'], []),
]


def test_titan_reference_table(gw):
    for body, override, path, must, must_not in TITAN:
        g, = run(gw, "aws-bedrock", [body], model_override=override or None)
        t = O.embeddings_translate("aws-bedrock", body, model_override=override)
        if path is None:
            assert g["status"] == 2 and t.status == 2, (body, g, t.status)   # ErrInvalidRequestBody
            continue
        assert g["status"] == 0 and g["path"].decode() == path == t.path, (body, g, t.path)
        assert g["body"] == t.body
        for m in must:
            assert m in g["body"]
        for m in must_not:
            assert m not in g["body"]
