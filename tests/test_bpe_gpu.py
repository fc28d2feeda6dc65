"""K4: byte-level BPE token count on the GPU (aigw_bpe_*) vs the self-oracle (oracle/bpe.hpp, pinned by `tokenizers`' own answers)."""
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


def test_tokenizers_known_answers(gw):
    cases = json.load(open(os.path.join(HERE, "golden", "bpe_cases.json"), encoding="utf-8"))["cases"]
    got = gw.bpe_count(gw.bpe, [c["text"] for c in cases])
    for c, g in zip(cases, got):
        if max(len(p) for p in c["text"].split(" ")) >= 511:
            assert int(g) == 0xFFFFFFFF          # a space-free run longer than the staging buffer is declined, not approximated
            continue
        assert int(g) == c["count"], (c["text"][:80], int(g), c["count"])


def words(r, n=4000):
    return ["".join(r.choice("etaoinshrdlcumwfgypbvkjxqz"[: r.randint(5, 26)]) for _ in range(r.randint(1, 12))) for _ in range(n)]


def test_corpus_vs_oracle(gw):
    r = random.Random(5)
    ws = words(r)
    texts = []
    for _ in range(30000):
        k = r.random()
        n = 64 if k < 0.7 else r.choice([0, 1, 5, 200, 511, 512, 513, 1023, 1024, 1025, 1500, 2047, 2048, 2049, 5000, 9000])
        t = ""
        while len(t) < n:
            t += (" " if t and r.random() < 0.97 else "") + ws[min(int(r.paretovariate(1.1)) - 1, len(ws) - 1)] + ("  " if r.random() < 0.01 else "")
        texts.append(t[:n])
    texts += ["", " ", "   ", "日本語 テキスト", "a" * 2048, "b" * 2049 + " c", " " * 3000, ("word " * 1000)]
    got = gw.bpe_count(gw.bpe, texts)
    exp = O.Bpe(VOCAB).count(texts)
    n_decl = 0
    for t, g, e in zip(texts, got, exp):
        if int(g) == 0xFFFFFFFF:
            assert max(len(p) for p in t.split(" ")) >= 511, t[:60]
            n_decl += 1
            continue
        assert int(g) == int(e), (len(t), t[:80], int(g), int(e))
    assert n_decl < 50


def test_embeddings_request_shape(gw):
    """BASELINE config 3's unit: 1024 inputs of 64 characters; the count for the request is the sum over its inputs"""
    r = random.Random(8)
    ws = words(r)
    inputs = []
    for _ in range(1024):
        t = ""
        while len(t) < 64:
            t += (" " if t else "") + ws[min(int(r.paretovariate(1.1)) - 1, len(ws) - 1)]
        inputs.append(t[:64])
    got = gw.bpe_count(gw.bpe, inputs)
    exp = O.Bpe(VOCAB).count(inputs)
    assert np.array_equal(got, exp) and int(got.sum()) > 10000
