"""K4 self-oracle: oracle/bpe.hpp against the answers HuggingFace `tokenizers` gave for the committed vocabulary (tests/golden/bpe_cases.json,
made by tests/golden/make_bpe_vocab.py), and — where the library is importable — live on random texts."""
import json
import os
import random

import pytest

import _oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_tokenizers_fixture():
    cases = json.load(open(os.path.join(HERE, "golden", "bpe_cases.json"), encoding="utf-8"))["cases"]
    got = O.Bpe().count([c["text"] for c in cases])
    assert len(cases) >= 600
    for c, g in zip(cases, got):
        assert int(g) == c["count"], c["text"][:80]


def test_oracle_matches_tokenizers_live():
    tk = pytest.importorskip("tokenizers")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_bpe_vocab as M
    tok, words = M.build()
    vocab = json.load(open(os.path.join(HERE, "golden", "bpe_vocab.json")))
    assert tok.get_vocab_size() == vocab["vocab_size"]      # the training is deterministic: the committed vocabulary is reproducible
    r = random.Random(99)
    texts = [M.text(r, words, r.choice([3, 64, 64, 300])) for _ in range(800)] + ["  double  spaces  ", "ünïcödé wörds", "\t\n"]
    got = O.Bpe(vocab).count(texts)
    for t, g in zip(texts, got):
        assert int(g) == len(tok.encode(t).ids), t[:80]
