"""Escaped strings through the CUDA chat translate path: u-escapes (incl. surrogate pairs), backslash-slash,
backslash-b and backslash-f are decoded by the reference and re-spelled by its encoder (oracle/ojson.hpp enc_str); the GPU path
must produce the same bytes, and escaped spellings of keys / enumerated values must keep their meaning."""
import collections
import json
import random

import pytest

import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import aigw_b200 as A
    c = A.Context(0)
    yield c
    c.close()


WORDS = ["héllo", "wörld", "中文字符", "emoji \U0001F600 ok", "tab\tsep", "line\nbreak", 'quo"te', "back\\slash", "plain", "<b>bold</b> & more", "a/b/c",
         "bell" + chr(8), "form" + chr(12), "  sep"]


def ascii_json_body(r, tricky):
    """a chat body serialised the way ASCII-only JSON encoders do it (every non-ASCII character as a u-escape, surrogate pairs for
    astral ones), plus escaped slashes and HTML-safe escapes of < > &"""
    def enc(s):
        out = json.dumps(s, ensure_ascii=True)
        if r.random() < 0.3:
            out = out.replace("/", "\\/")
        if r.random() < 0.3:
            out = out.replace("<", "\\u003c").replace(">", "\\u003E").replace("&", "\\u0026")
        return out
    text = lambda: enc(" ".join(r.choice(WORDS) for _ in range(r.randint(1, 25))))
    msgs = []
    for _ in range(r.randint(1, 6)):
        role = r.choice(["system", "user", "assistant", "developer"])
        role_js = json.dumps(role)
        if tricky and r.random() < 0.3:   # escaped spellings of an enumerated value keep their meaning
            role_js = {"user": '"us\\u0065r"', "assistant": '"\\u0061ssistant"', "system": '"syst\\u0065m"', "developer": '"developer"'}[role]
        if role == "user" and r.random() < 0.3:
            tkey = '"ty\\u0070e"' if tricky else '"type"'
            msgs.append('{"role":%s,"content":[{"type":"text","text":%s},{%s:"text","text":%s}]}' % (role_js, text(), tkey, text()))
        else:
            msgs.append('{"role":%s,"content":%s}' % (role_js, text()))
    extra = r.choice(["", ',"max_tokens":64', ',"stream":true', ',"temperature":0.5', ',"stop":[%s]' % text()])
    mkey = '"mod\\u0065l"' if tricky and r.random() < 0.2 else '"model"'
    return ('{%s:"gpt-4o","messages":[%s]%s}' % (mkey, ",".join(msgs), extra)).encode()


def check(ctx, schema, bodies, must_accept):
    import aigw_b200 as A
    got = ctx.chat_translate(ctx.cfg(schema), bodies)
    declined = collections.Counter()
    for i, (b, g) in enumerate(zip(bodies, got)):
        o = O.chat_translate(schema, b, prefix="" if "anthropic" in schema else "v1")
        if g["status"] == A.AIGW_OK:
            assert o.status == O.OK, (i, b[:200], o.err)
            exp = o.body if o.body_kind == O.BYTES else b""
            assert g["body"] == exp, (i, b[:400], g["body"][:400], exp[:400])
            assert g["path"].decode() == o.path and g["model"] == o.model and g["stream"] == o.stream
        elif g["status"] == A.AIGW_DECLINED:
            declined[g["reason"]] += 1
            assert not must_accept, (i, g["reason"], b[:300])
        else:
            assert g["status"] == o.status, (i, b[:300], g["status"], g["reason"], o.status, o.err)
            declined["status%d" % g["status"]] += 1
    return declined


@pytest.mark.parametrize("schema", ["aws-bedrock", "gcp-vertexai", "gcp-anthropicai", "openai"])
def test_escaped_strings(ctx, schema):
    r = random.Random(31)
    plain = [ascii_json_body(r, False) for _ in range(600)]
    d = check(ctx, schema, plain, must_accept=False)
    if schema != "gcp-anthropicai":   # the only admissible declines are capacity ones (ops / scratch) on escape-dense bodies
        assert set(d) <= {6, 7} and sum(d.values()) <= len(plain) // 10, dict(d)
    tricky = [ascii_json_body(r, True) for _ in range(600)]
    d2 = check(ctx, schema, tricky, must_accept=False)
    print(schema, "plain declined", dict(d), "tricky declined", dict(d2))
    assert sum(d2.values()) < 300


def test_malformed_escapes_never_translate(ctx):
    import aigw_b200 as A
    bad = [b'{"model":"m","messages":[{"role":"user","content":"lone \\ud83d surrogate"}]}', b'{"model":"m","messages":[{"role":"user","content":"bad \\u12G4 hex"}]}',
           b'{"model":"m","messages":[{"role":"user","content":"bad \\x escape"}]}', b'{"model":"m","messages":[{"role":"user","content":"trunc \\u12"}]}']
    for g in ctx.chat_translate(ctx.cfg("aws-bedrock"), bad):
        assert g["status"] != A.AIGW_OK
