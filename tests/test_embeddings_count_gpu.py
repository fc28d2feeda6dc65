"""BASELINE config 3: /v1/embeddings requests above the 64 KiB ceiling of the translate pass — EmbeddingsEndpointSpec.ParseBody over the
string forms of the input union + the BPE count of every input, one call (aigw_embeddings_count_*).  Verdict, model and n_inputs against
the embeddings oracle (oracle/embeddings.hpp, pinned by the reference's goldens); token counts against the BPE self-oracle."""
import json
import os
import random

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB = json.load(open(os.path.join(HERE, "golden", "bpe_vocab.json")))


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    g.bpe = g.bpe_load(VOCAB["byte_to_id"], VOCAB["merges"])
    yield g
    g.bpe_free(g.bpe)
    g.close()


def words(r, n=3000):
    return ["".join(r.choice("etaoinshrdlcumwfgypbvkjxqz"[: r.randint(5, 26)]) for _ in range(r.randint(1, 12))) for _ in range(n)]


def text(r, ws, n):
    t = ""
    while len(t) < n:
        t += (" " if t else "") + ws[min(int(r.paretovariate(1.1)) - 1, len(ws) - 1)]
    return t[:n]


def check(gw, bodies, bpe):
    res = gw.embeddings_count(gw.bpe, bodies)
    n_ok = n_decl = n_err = 0
    for b, x in zip(bodies, res):
        t = O.embeddings_translate("openai", b)
        if x["status"] == 4:
            n_decl += 1
            continue
        assert int(x["status"]) == t.status, (b[:200], x, t.status, t.err)
        if t.status != 0:
            n_err += 1
            continue
        d = json.loads(b)
        inp = d.get("input")     # a present null is a 400 (checked above through the status)
        texts = [] if inp is None else [inp] if isinstance(inp, str) else inp
        assert int(x["n_inputs"]) == len(texts), (b[:200], x)
        mo, ml = int(x["model_off"]), int(x["model_len"])
        assert b[mo:mo + ml] == t.model, (b[:100], b[mo:mo + ml], t.model)
        exp = bpe.count(texts) if texts else np.zeros(0, np.uint32)
        declined = sum(1 for s in texts if max(len(p) for p in s.split(" ")) >= 511) if texts else 0
        if int(x["declined_inputs"]) == 0:
            assert int(x["tokens"]) == int(np.asarray(exp, dtype=np.uint64).sum()), (b[:120], x, int(exp.sum()))
        else:
            assert int(x["declined_inputs"]) <= declined
        n_ok += 1
    return n_ok, n_decl, n_err


def test_config3_shape(gw):
    """1024 inputs of 64 characters: 68.7 KB of JSON per request, above the translate pass's 64 KiB ceiling"""
    r = random.Random(3)
    ws = words(r)
    bpe = O.Bpe(VOCAB)
    bodies = []
    for i in range(48):
        d = {"model": "text-embedding-3-small", "input": [text(r, ws, 64) for _ in range(1024)]}
        bodies.append(json.dumps(d, separators=(",", ":")).encode())
    assert min(len(b) for b in bodies) > 65536
    ok, decl, err = check(gw, bodies, bpe)
    assert ok == len(bodies) and decl == 0 and err == 0


def emb_body(r, ws):
    k = r.random()
    n = r.choice([0, 1, 2, 3, 7, 64, 300, 1024, 2500]) if k < 0.8 else 1
    inputs = [text(r, ws, r.choice([0, 1, 5, 64, 64, 64, 200, 700])) for _ in range(n)]
    d = {"model": r.choice(["text-embedding-3-small", "m", "text-embedding-ada-002"]), "input": inputs if k < 0.8 else text(r, ws, 64)}
    if r.random() < 0.2:
        d["encoding_format"] = r.choice(["float", "base64", None])
    if r.random() < 0.2:
        d["dimensions"] = r.choice([256, 1536, None])
    if r.random() < 0.1:
        d["user"] = "u-1"
    items = list(d.items())
    r.shuffle(items)
    kind = r.random()
    if kind < 0.04:
        items = [(a, (5 if a == "model" else v)) for a, v in items]                  # 400
    elif kind < 0.07:
        items = [(a, (1.5 if a == "dimensions" else v)) for a, v in items]           # 400 when present
    elif kind < 0.10:
        items = [(a, ([1, 2, 3] if a == "input" else v)) for a, v in items]          # token ids: outside the decided shape
    elif kind < 0.13:
        items = [(a, ({"content": "x"} if a == "input" else v)) for a, v in items]   # object input
    elif kind < 0.16:
        items.append(("task_type", "RETRIEVAL_QUERY"))                              # vendor member
    elif kind < 0.19:
        items = [(a, (v + ['quo"te'] if a == "input" and isinstance(v, list) else v)) for a, v in items]   # an escape
    elif kind < 0.21:
        items = [(a, (v + [None] if a == "input" and isinstance(v, list) else v)) for a, v in items]
    elif kind < 0.23:
        items = [(a, (True if a == "input" else v)) for a, v in items]               # 400
    body = json.dumps(dict(items), separators=(",", ":") if r.random() < 0.7 else (", ", ": "), ensure_ascii=False)
    if 0.30 <= kind < 0.34:
        body = json.dumps(dict(items), indent=2)
    elif 0.34 <= kind < 0.37:
        body = body[: max(1, len(body) * r.randint(1, 9) // 10)]                      # truncated
    elif 0.37 <= kind < 0.39:
        body = body[:-1] + ',"model":"again"}'
    elif 0.39 <= kind < 0.41:
        body = body + r.choice([" ", "\n", "x", "}", " {}"])
    elif 0.41 <= kind < 0.43:
        body = body.replace('","', '" "', 1)
    return body.encode()


ODD = [b'', b' ', b'null', b'5', b'[]', b'{}', b' { } ', b'{"model":"m"}', b'{"input":[]}', b'{"input":null,"model":null}', b'{"input":"one text"}', b'{"input":[""]}', b'{"input":["a",]}', b'{"input":[,"a"]}',
       b'{"input":["a""b"]}', b'{"input":["a"],}', b'{"input":["a"]', b'{"input":["a"]}}', b'{"input":["a"]]}', b'{"input":[["a"]]}', b'{"input":["a"],"input":["b"]}', b'{"model":"m","input":["a"] ,"dimensions": 12 }',
       b'{"model":"m","dimensions":-1}', b'{"model":"m","dimensions":01}', b'{"model":"m","dimensions":1e2}', b'{"model":"m","dimensions":tru}', b'{"model":"m","user":false}', b'{"model":"m" "input":[]}',
       b'{"model":"m","input":["a\tb"]}', b'{"model":"m","input":["\xc3\xa9t\xc3\xa9 \xe6\x97\xa5\xe6\x9c\xac"]}', b'{"Model":"m"}', b'{"model":"m","input":[true]}', b'{"model":"m","input":["a",1]}', b'{"model":"m","input":"a" "b"}',
       b'{"model":"m",}', b'{,"model":"m"}', b'{"model"}', b'{"model":}', b'{"model":"m","input":["a"]}\n\n', b'\n {"model":"m","input":[ "a" , "b" ]\n}', b'{"input":["' + b"x" * 600 + b'"]}', b'{"input":["' + b"y " * 5000 + b'"]}']


def test_parity_corpus(gw):
    r = random.Random(41)
    ws = words(r)
    bpe = O.Bpe(VOCAB)
    bodies = [emb_body(r, ws) for _ in range(1500)] + ODD
    ok, decl, err = check(gw, bodies, bpe)
    print("ok", ok, "declined", decl, "errors", err, "of", len(bodies))
    assert ok > 600 and err > 60
    # what the oracle accepts and holds no escape / nesting / vendor member, the GPU accepts
    plain = [b for b in bodies[:1500] if O.embeddings_translate("openai", b).status == 0 and b"\\" not in b and b"null]" not in b and b"task_type" not in b and b"{\"content" not in b and b"[1,2,3]" not in b and b"[1, 2, 3]" not in b and b"again" not in b]
    res = gw.embeddings_count(gw.bpe, plain)
    declined = [b[:120] for b, x in zip(plain, res) if x["status"] == 4]
    assert len(plain) > 500 and len(declined) <= 3, declined[:5]


def test_sixteen_byte_alignment_is_required(gw):
    from aigw_b200 import capi
    arena = np.frombuffer(b" " * 8 + b'{"input":["a"]}' + b" " * 32, dtype=np.uint8).copy()
    with pytest.raises(Exception):
        gw.embeddings_count_host(gw.bpe, arena, np.asarray([8, 0], dtype=np.uint64), np.asarray([15], dtype=np.uint32))
