#!/usr/bin/env python
"""K4 fixtures: a byte-level BPE vocabulary trained HERE with `tokenizers` (HuggingFace, 0.22) on a seeded synthetic corpus, and known
answers produced by that library.  The reference has no tokenizer (token counts come from provider responses), so this is a
SELF-ORACLE, not reference parity: the GPU kernel and oracle/bpe.hpp must reproduce what `tokenizers` computes with this vocabulary.

    python tests/golden/make_bpe_vocab.py        # rewrites tests/golden/bpe_vocab.json and tests/golden/bpe_cases.json

Tokenizer definition (chosen so that the pre-tokenisation is a byte rule, not a Unicode-category regex):
  pre-tokeniser  Split(" ", merged_with_next) then ByteLevel(add_prefix_space=False, use_regex=False): every space starts a new piece,
                 the piece is the space plus the bytes up to the next space; bytes map to the 256 byte-level symbols
  model          BPE without dropout / unk / byte fallback: inside a piece, repeatedly merge the adjacent pair of lowest rank
                 (leftmost first among equal ranks) until no pair is in the merge table
"""
import json
import os
import random

from tokenizers import Tokenizer, models, pre_tokenizers, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB_SIZE = 4096


def zipf_words(r, n=6000):
    alpha = "etaoinshrdlcumwfgypbvkjxqz"
    return ["".join(r.choice(alpha[: r.randint(6, 26)]) for _ in range(r.randint(1, 11))) for _ in range(n)]


def text(r, words, nchar=64):
    out = ""
    while len(out) < nchar:
        w = words[min(int(r.paretovariate(1.05)) - 1, len(words) - 1)]
        if r.random() < 0.03: w = w.capitalize()
        if r.random() < 0.02: w += r.choice([",", ".", "!", "?", ";"])
        if r.random() < 0.01: w = r.choice(["café", "日本", "naïve", "Ünï"])
        out += (" " if out else "") + w
    return out[:nchar]


def build():
    r = random.Random(3)
    words = zipf_words(r)
    corpus = [text(r, words, 200) for _ in range(40000)]
    tok = Tokenizer(models.BPE(unk_token=None))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(" ", behavior="merged_with_next"), pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=VOCAB_SIZE, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False))
    return tok, words


def byte_symbols():
    """GPT-2 byte ↔ unicode map used by ByteLevel"""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def main():
    tok, words = build()
    model = json.loads(tok.to_str())["model"]
    vocab = model["vocab"]
    b2s = byte_symbols()
    byte_to_id = [vocab[b2s[b]] for b in range(256)]
    merges = [[vocab[a], vocab[b], vocab[a + b]] for a, b in model["merges"]]
    json.dump({"about": "byte-level BPE trained by tests/golden/make_bpe_vocab.py with tokenizers " + __import__("tokenizers").__version__ + " (seed 3); self-oracle for K4",
               "vocab_size": tok.get_vocab_size(), "byte_to_id": byte_to_id, "merges": merges}, open(os.path.join(HERE, "bpe_vocab.json"), "w"))
    r = random.Random(11)
    texts = [text(r, words, r.choice([64, 64, 64, 7, 200, 1000])) for _ in range(600)]
    texts += ["", " ", "  ", "a", " a", "a ", "a  b", "   lead and trail   ", "aaaaaaaaaaaaaaaaaaaaaaaa", "x" * 300, "tab\tand\nnewline here", "日本語のテキスト と café", "MiXeD CaSe WoRdS", "e" * 5000 + " tail"]
    cases = [{"text": t, "count": len(tok.encode(t).ids)} for t in texts]
    json.dump({"cases": cases}, open(os.path.join(HERE, "bpe_cases.json"), "w"), ensure_ascii=False)
    print("vocab", tok.get_vocab_size(), "merges", len(merges), "cases", len(cases), "mean tokens per 64-char text", sum(c["count"] for c in cases[:600]) / 600)


if __name__ == "__main__":
    main()
