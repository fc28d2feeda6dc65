"""Workload helpers for tests and bench: ctypes view of tools/libworkload.so plus a seeded Python
generator of structurally diverse ChatCompletion bodies (edge-case corpus)."""
import ctypes as C
import json
import os
import random
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tools", "libworkload.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "tools", "workload.cpp")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SO, src])
        _lib = C.CDLL(_SO)
        _lib.wl_sse_fill.restype = C.c_uint64
    return _lib


def chat_lens(seed, first, n, target=4096, jitter=32, threads=8):
    lens = np.zeros(n, dtype=np.uint32)
    lib().wl_chat_lens(C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), C.c_uint32(target), C.c_uint32(jitter), C.c_void_p(lens.ctypes.data), C.c_int(threads))
    return lens


def chat_offsets(lens):
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum((lens.astype(np.uint64) + 15) // 16 * 16, out=offs[1:])
    return offs


def chat_fill(seed, first, n, out_arr, offs, lens, target=4096, jitter=32, threads=8):
    lib().wl_chat_fill(C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), C.c_uint32(target), C.c_uint32(jitter), C.c_void_p(out_arr.ctypes.data),
                       C.c_void_p(offs.ctypes.data), C.c_void_p(lens.ctypes.data), C.c_int(threads))


def chat_corpus(seed, first, n, target=4096, jitter=32, threads=8):
    """(arena uint8[], offsets uint64[n+1], lens uint32[n]) for the C2-shaped workload."""
    lens = chat_lens(seed, first, n, target, jitter, threads)
    offs = chat_offsets(lens)
    arena = np.full(int(offs[-1]) + 16, 0x20, dtype=np.uint8)
    chat_fill(seed, first, n, arena, offs, lens, target, jitter, threads)
    return arena, offs, lens


def sse_corpus(seed, first, n, chunks=256, chunk_bytes=80):
    total = lib().wl_sse_fill(C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), C.c_int(chunks), C.c_int(chunk_bytes), None, C.c_uint64(0), None)
    buf = np.zeros(int(total) + 64, dtype=np.uint8)
    coff = np.zeros(n * chunks + 1, dtype=np.uint64)
    lib().wl_sse_fill(C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), C.c_int(chunks), C.c_int(chunk_bytes), C.c_void_p(buf.ctypes.data), C.c_uint64(total), C.c_void_p(coff.ctypes.data))
    cfirst = (np.arange(n + 1, dtype=np.uint32) * chunks).astype(np.uint32)
    return buf, coff, cfirst


# ---------------------------------------------------------------- diverse chat bodies (edge-case corpus)
_WORDS = "alpha beta gamma delta émigré 中文 naïve tab\\tsep quote\\\"d back\\\\slash new\\nline plain text with spaces <tag> & amp / slash".split(" ")


def _text(r, lo=0, hi=60):
    n = r.randint(lo, hi)
    return " ".join(r.choice(_WORDS) for _ in range(n))


def _jstr(s):
    # our own quoting so that the escapes stay in the canonical set the encoder itself would produce
    return '"' + s + '"'


def _num(r):
    return r.choice(["0", "1", "0.7", "0.25", "1.0", "2", "0.95", "0.10", "1.5", "0.000001", "12", "-0.5", "1e0", "0.7000000000000001", "0.1234567890123456789"])


def _schema(r, depth=0):
    if depth > 2 or r.random() < 0.3:
        return r.choice(['{"type":"string"}', '{"type": "integer", "minimum": 0}', '{"description":"d","type":"string","enum":["a","b"]}', '{"type":"number","maximum":1.50}'])
    props = ",".join(f'"{r.choice(["path","depth","query","zeta","alpha","mid"])}{i}":{_schema(r, depth + 1)}' for i in range(r.randint(0, 3)))
    sp = r.choice(["", " ", "\n  "])
    return "{" + sp + '"type":"object",' + sp + '"properties":{' + props + '},' + sp + '"required":["path0"]' + r.choice(["", ',"additionalProperties":false']) + "}"


def diverse_body(seed):
    r = random.Random(seed)
    sp = r.choice(["", "", "", " ", "\n  "])
    msgs = []
    n = r.randint(0, 7)
    for i in range(n):
        role = r.choice(["system", "user", "user", "assistant", "assistant", "tool", "developer"])
        if role in ("system", "developer"):
            c = r.choice(["str", "str", "parts", "parts_cache"])
            if c == "str":
                content = _jstr(_text(r))
            else:
                parts = []
                for _ in range(r.randint(0, 3)):
                    cc = ',"cache_control":{"type":"ephemeral"}' if c == "parts_cache" and r.random() < 0.5 else ""
                    parts.append('{"type":"text","text":' + _jstr(_text(r, 0, 10)) + cc + "}")
                content = "[" + ",".join(parts) + "]"
            msgs.append('{"role":"%s",%s"content":%s}' % (role, sp, content))
        elif role == "user":
            c = r.choice(["str", "str", "str", "parts", "null", "named"])
            if c in ("str", "named"):
                nm = ',"name":"bob"' if c == "named" else ""
                msgs.append('{"role":"user"%s,"content":%s}' % (nm, _jstr(_text(r))))
            elif c == "null":
                msgs.append('{"role":"user","content":null}')
            else:
                parts = ['{"type":"text","text":' + _jstr(_text(r, 0, 10)) + r.choice(["", ',"cache_control":{"type":"ephemeral"}', ',"cache_control":{"type":"other"}']) + "}"
                         for _ in range(r.randint(0, 3))]
                msgs.append('{"content":[' + ",".join(parts) + '],"role":"user"}')
        elif role == "assistant":
            c = r.choice(["str", "empty", "null", "toolcall", "toolcall2", "parts", "single", "absent"])
            if c == "str":
                msgs.append('{"role":"assistant","content":%s}' % _jstr(_text(r, 1, 40)))
            elif c == "empty":
                msgs.append('{"role":"assistant","content":""}')
            elif c == "null":
                msgs.append('{"role":"assistant","content":null}')
            elif c == "absent":
                msgs.append('{"role":"assistant"}')
            elif c in ("toolcall", "toolcall2"):
                calls = []
                for k in range(1 if c == "toolcall" else 2):
                    args = r.choice(['{ \\"path\\": \\"/tmp\\" }', '{\\"b\\":1,\\"a\\":[1,2,{\\"z\\":null,\\"y\\":true}]}', '{}', '{\\"q\\":\\"say \\\\\\"hi\\\\\\"\\",\\"n\\":1.50}', 'null',
                                     '{\\"x\\": 1e3}', '{\\"a\\":1,\\"a\\":2}'])
                    calls.append('{"id":"call_%d","type":"function","function":{"name":"fn%d","arguments":"%s"}}' % (r.randint(1, 999), k, args))
                pre = r.choice(['"content":null,', '"content":"thinking out loud",', ""])
                msgs.append('{"role":"assistant",%s"tool_calls":[%s]}' % (pre, ",".join(calls)))
            elif c == "parts":
                parts = []
                for _ in range(r.randint(0, 3)):
                    t = r.choice(["text", "refusal", "thinking", "thinking_sig", "unknown"])
                    if t == "text":
                        parts.append('{"type":"text","text":' + _jstr(_text(r, 0, 8)) + "}")
                    elif t == "refusal":
                        parts.append('{"type":"refusal","refusal":"no"}')
                    elif t == "thinking":
                        parts.append('{"type":"thinking","text":"hmm"}')
                    elif t == "thinking_sig":
                        parts.append('{"type":"thinking","text":"hmm","signature":"c2ln"}')
                    else:
                        parts.append('{"type":"other","text":"x"}')
                msgs.append('{"role":"assistant","content":[' + ",".join(parts) + "]}")
            else:
                msgs.append('{"role":"assistant","content":{"type":"text","text":"single"}}')
        else:
            c = r.choice(["str", "parts"])
            content = _jstr(_text(r, 0, 12)) if c == "str" else '[{"type":"text","text":"r1"},{"type":"text","text":"r2"}]'
            msgs.append('{"role":"tool","tool_call_id":"call_%d","content":%s}' % (r.randint(1, 999), content))
    top = ['"model":%s' % _jstr(r.choice(["gpt-4o-mini", "anthropic.claude-3-sonnet", "arn:aws:bedrock:us-east-1:123:model/x y", "m"])),
           '"messages":[' + ("," + sp).join(msgs) + "]"]
    if r.random() < 0.5:
        top.append('"temperature":' + _num(r))
    if r.random() < 0.3:
        top.append('"top_p":' + _num(r))
    if r.random() < 0.5:
        top.append(r.choice(['"max_tokens":256', '"max_completion_tokens":1024', '"max_tokens":10,"max_completion_tokens":20', '"max_tokens":null', '"max_tokens":1.5']))
    if r.random() < 0.3:
        top.append(r.choice(['"stop":"END"', '"stop":["a","b"]', '"stop":[]', '"stop":null', '"stop":5']))
    if r.random() < 0.4:
        top.append(r.choice(['"stream":true', '"stream":false', '"stream":true,"stream_options":{"include_usage":true}', '"stream":"yes"']))
    if r.random() < 0.3:
        tools = []
        for k in range(r.randint(0, 2)):
            desc = r.choice(['"description":"does things",', '"description":"",', ""])
            tools.append('{"type":"function","function":{"name":"fn%d",%s"parameters":%s}}' % (k, desc, _schema(r)))
        top.append('"tools":[' + ",".join(tools) + "]")
        if r.random() < 0.6:
            top.append('"tool_choice":' + r.choice(['"auto"', '"required"', '"none"', '"fn0"', '{"type":"function","function":{"name":"fn0"}}', "null", "7"]))
    if r.random() < 0.15:
        top.append(r.choice(['"thinking":{"type":"enabled","budget_tokens":4096}', '"thinking":{"type":"disabled"}', '"thinking":{"type":"adaptive"}', '"thinking":{"type":"bogus"}']))
    if r.random() < 0.2:
        top.append(r.choice(['"service_tier":"flex"', '"user":"u-1"', '"n":1', '"seed":42', '"logprobs":true', '"unknown_field":{"a":[1,2,{"b":null}]}', '"response_format":{"type":"json_object"}',
                             '"presence_penalty":0.5', '"n":"one"', '"parallel_tool_calls":false']))
    r.shuffle(top)
    body = "{" + sp + ("," + sp).join(top) + sp + "}"
    if r.random() < 0.03:
        body = body[: r.randint(1, len(body) - 1)]  # truncated ⇒ malformed
    return body.encode("utf-8")


# ---------------------------------------------------------------- Bedrock ConverseStream (AWS eventstream) corpora
def es_frame(headers, payload: bytes, corrupt=None) -> bytes:
    """One AWS eventstream message: [total u32][headers_len u32][prelude crc][headers][payload][message crc] (big endian, IEEE CRC-32).
    headers: list of (name bytes, type int, value bytes); type 7 values get their u16 length prefix here."""
    import struct, zlib
    hb = b""
    for k, t, v in headers:
        hb += bytes([len(k)]) + k + bytes([t]) + (struct.pack(">H", len(v)) + v if t in (6, 7) else v)
    total = 16 + len(hb) + len(payload)
    pre = struct.pack(">II", total, len(hb))
    pcrc = zlib.crc32(pre) ^ (1 if corrupt == "prelude" else 0)
    msg = pre + struct.pack(">I", pcrc) + hb + payload
    mcrc = zlib.crc32(msg) ^ (1 if corrupt == "message" else 0)
    return msg + struct.pack(">I", mcrc)


def es_event(event_type: str, payload: bytes, extra_headers=True, corrupt=None) -> bytes:
    h = [(b":event-type", 7, event_type.encode())]
    if extra_headers:
        h += [(b":content-type", 7, b"application/json"), (b":message-type", 7, b"event")]
    return es_frame(h, payload, corrupt)


_WORDS = ["Hello", " world", "!", " The", " quick", " brown", " fox", "\\n", " \\\"quoted\\\"", " café", " 日本", " tab\\t", " back\\\\slash", " a", " of", " and", " 42", ".", ","]


def bedrock_stream_bytes(rng, kind="plain") -> bytes:
    """One ConverseStream response as eventstream bytes.  kind: plain | tools | reasoning | cache | odd (decoder / schema corner cases)."""
    import json as _json
    pad = lambda: '"p":"' + "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"[: int(rng.integers(1, 40))] + '"'
    fr = []
    fr.append(es_event("messageStart", ('{' + pad() + ',"role":"assistant"}').encode()))
    blk = 0
    if kind == "reasoning":
        for _ in range(int(rng.integers(1, 6))):
            t = "".join(_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), int(rng.integers(1, 8))))
            fr.append(es_event("contentBlockDelta", ('{"contentBlockIndex":%d,"delta":{"reasoningContent":{"text":"%s"}},%s}' % (blk, t, pad())).encode()))
        fr.append(es_event("contentBlockDelta", ('{"contentBlockIndex":%d,"delta":{"reasoningContent":{"signature":"c2lnbmF0dXJl%d"}},%s}' % (blk, int(rng.integers(0, 1000)), pad())).encode()))
        fr.append(es_event("contentBlockStop", ('{"contentBlockIndex":%d,%s}' % (blk, pad())).encode())); blk += 1
    for _ in range(int(rng.integers(1, 40))):
        t = "".join(_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), int(rng.integers(1, 10))))
        fr.append(es_event("contentBlockDelta", ('{"contentBlockIndex":%d,"delta":{"text":"%s"},%s}' % (blk, t, pad())).encode()))
    fr.append(es_event("contentBlockStop", ('{"contentBlockIndex":%d,%s}' % (blk, pad())).encode())); blk += 1
    stop = "end_turn"
    if kind == "tools":
        for k in range(int(rng.integers(1, 4))):
            fr.append(es_event("contentBlockStart", ('{"contentBlockIndex":%d,%s,"start":{"toolUse":{"name":"get_weather_%d","toolUseId":"tooluse_%08x"}}}' % (blk, pad(), k, int(rng.integers(0, 2**31)))).encode()))
            args = _json.dumps({"location": "San Francisco, CA", "unit": "celsius", "n": k})
            for i in range(0, len(args), 9):
                fr.append(es_event("contentBlockDelta", ('{"contentBlockIndex":%d,"delta":{"toolUse":{"input":%s}},%s}' % (blk, _json.dumps(args[i:i + 9]), pad())).encode()))
            fr.append(es_event("contentBlockStop", ('{"contentBlockIndex":%d,%s}' % (blk, pad())).encode())); blk += 1
        stop = "tool_use"
    else:
        stop = ["end_turn", "max_tokens", "stop_sequence", "content_filtered", "guardrail_intervened"][int(rng.integers(0, 5))]
    fr.append(es_event("messageStop", ('{%s,"stopReason":"%s"}' % (pad(), stop)).encode()))
    usage = '"inputTokens":%d,"outputTokens":%d,"totalTokens":%d' % (int(rng.integers(0, 5000)), int(rng.integers(0, 3000)), int(rng.integers(0, 8000)))
    if kind == "cache":
        usage += ',"cacheReadInputTokens":%d,"cacheWriteInputTokens":%d' % (int(rng.integers(0, 3)) * 128, int(rng.integers(0, 3)) * 64)
    tier = ',"serviceTier":{"type":"priority"}' if kind == "cache" and rng.integers(0, 2) else ""
    fr.append(es_event("metadata", ('{"metrics":{"latencyMs":%d},%s,"usage":{%s}%s}' % (int(rng.integers(100, 9000)), pad(), usage, tier)).encode()))
    if kind == "odd":
        extras = [
            es_event("metadata", b'{"usage":null}'), es_event("metadata", b'{"metrics":{"latencyMs":1}}'),
            es_event("messageStart", b'{"role":null}'), es_event("messageStart", b'{"role":""}'), es_event("contentBlockDelta", b'{"delta":{}}'),
            es_event("contentBlockDelta", b'{"delta":null}'), es_event("contentBlockStart", b'{"start":{}}'), es_event("contentBlockStart", b'{"start":{"toolUse":{}}}'),
            es_event("contentBlockDelta", b'{"delta":{"text":""}}'), es_event("contentBlockDelta", b'{"delta":{"toolUse":{"input":null}}}'),
            es_event("contentBlockDelta", b'{"delta":{"reasoningContent":{}}}'), es_event("contentBlockDelta", b'{"delta":{"reasoningContent":{"redactedContent":""}}}'),
            es_event("somethingElse", b'{"a":1}'), es_event("contentBlockDelta", b'not json'), es_event("contentBlockDelta", b'{"delta":{"text":5}}'),
            es_event("contentBlockDelta", b'{"contentBlockIndex":1.5,"delta":{"text":"x"}}'), es_event("messageStop", b'{"stopReason":null}'), es_event("messageStop", b'{}'),
            es_event("contentBlockStop", b'null'), es_event("contentBlockStop", b'{}'), es_event("messageStop", b' {"stopReason" : "tool_use" } '),
            es_frame([(b"x", 0, b""), (b"y", 1, b""), (b"b", 2, b"\x01"), (b"s", 3, b"\x00\x01"), (b"i", 4, b"\x00\x00\x00\x01"), (b"l", 5, b"\x00" * 8), (b"bin", 6, b"abc"),
                      (b"ts", 8, b"\x00" * 8), (b"uuid", 9, b"\x00" * 16), (b":event-type", 7, b"contentBlockDelta")], b'{"delta":{"text":"typed headers"}}'),
            es_frame([], b'{"eventType":"contentBlockDelta","delta":{"text":"type from payload"}}'),
            es_frame([(b":event-type", 6, b"contentBlockDelta")], b'{"delta":{"text":"bytes-typed header is ignored"}}'),
            es_event("metadata", b'{"usage":{"inputTokens":7,"outputTokens":0,"totalTokens":7,"cacheReadInputTokens":null,"cacheWriteInputTokens":0}}'),
            es_event("metadata", b'{"usage":{},"serviceTier":{"type":""}}'),
            es_event("contentBlockStart", b'{"start":{"toolUse":{"name":"f","toolUseId":"id1"}}}'), es_event("contentBlockStop", b'{}'), es_event("contentBlockStop", b'{}'),
            es_event("contentBlockStart", b'{"start":{"toolUse":{"name":"g","toolUseId":"id2"}}}'), es_event("contentBlockDelta", b'{"delta":{"toolUse":{"input":"{}"}}}'),
        ]
        order = rng.permutation(len(extras))
        k = int(rng.integers(3, len(extras)))
        pos = sorted(int(x) for x in rng.integers(1, len(fr), k))
        for j, p in zip(order[:k], reversed(pos)):
            fr.insert(p, extras[int(j)])
    return b"".join(fr)


def bedrock_stream_corpus(n, seed=0, kinds=("plain", "tools", "reasoning", "cache")):
    """n streams → (bytes array, stream_off u64[n+1], list of per-stream bytes)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    streams = [bedrock_stream_bytes(rng, kinds[i % len(kinds)]) for i in range(n)]
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in streams], out=off[1:])
    return np.frombuffer(b"".join(streams), dtype=np.uint8).copy(), off, streams


# ---------------------------------------------------------------- buffered Bedrock Converse responses (R1, Bedrock)
def bedrock_response_body(rng, kind="plain") -> bytes:
    import json as _json
    words = ["Hello", "world", "the", "quick", "brown", "fox", "café", "日本", "line\nbreak", "quote\"d", "tab\tbed", "back\\slash", "42", "ok."]
    txt = lambda k: " ".join(words[int(i)] for i in rng.integers(0, len(words), k))
    blocks = []
    if kind in ("reasoning",):
        blocks.append({"reasoningContent": {"reasoningText": {"text": txt(int(rng.integers(3, 30))), "signature": "c2ln" + str(int(rng.integers(0, 999)))}}})
    blocks.append({"text": txt(int(rng.integers(1, 120)))})
    if rng.integers(0, 4) == 0:
        blocks.append({"text": "second text block is ignored"})
    stop = ["end_turn", "max_tokens", "stop_sequence", "content_filtered", "guardrail_intervened", None][int(rng.integers(0, 6))]
    if kind == "tools":
        for k in range(int(rng.integers(1, 4))):
            blocks.append({"toolUse": {"toolUseId": "tooluse_%06x" % int(rng.integers(0, 2**24)), "name": "get_weather_%d" % k,
                                       "input": {"location": txt(2), "unit": "celsius", "n": k, "ratio": 0.25 * k, "nested": {"z": [1, 2, {"y": None}], "a": True}}}})
        stop = "tool_use"
    usage = {"inputTokens": int(rng.integers(0, 9000)), "outputTokens": int(rng.integers(0, 4000)), "totalTokens": int(rng.integers(0, 13000))}
    doc = {"metrics": {"latencyMs": int(rng.integers(10, 9000))}, "output": {"message": {"content": blocks, "role": "assistant"}}, "stopReason": stop, "usage": usage}
    if kind == "cache":
        usage["cacheReadInputTokens"] = int(rng.integers(0, 3)) * 100
        if rng.integers(0, 2): usage["cacheWriteInputTokens"] = int(rng.integers(0, 3)) * 10
        if rng.integers(0, 2): doc["serviceTier"] = {"type": ["priority", "default", ""][int(rng.integers(0, 3))]}
    if rng.integers(0, 8) == 0:
        del doc["metrics"]
    return _json.dumps(doc, ensure_ascii=False, separators=(",", ":")).encode()


BEDROCK_RESPONSE_ODD = [
    b'{"output":{"message":{"content":[{"text":"a"}],"role":"assistant"}},"stopReason":"end_turn","usage":null}',
    b'{"output":{"message":{"content":null,"role":null}},"stopReason":null}',
    b'{"output":{"message":null}}', b'{"output":{}}', b'{"output":{"message":{}},"usage":{"inputTokens":0,"outputTokens":0,"totalTokens":0}}',
    b'{"output":{"message":{"content":[],"role":""}},"usage":{"inputTokens":0,"outputTokens":0,"totalTokens":0,"cacheReadInputTokens":0}}',
    b'{"output":{"message":{"content":[{"toolUse":{"name":null,"input":null,"toolUseId":null}}]}}}',
    b'{"output":{"message":{"content":[{"toolUse":{}},{"toolUse":{"input":{}}}]}}}',
    b'{"output":{"message":{"content":[{"toolUse":{"name":"f","input":{"q":"he said \\"hi\\"\\n","b":{"c":[1.50,2e0,-0]}},"toolUseId":"x"},"text":"both"}]}}}',
    b'{"output":{"message":{"content":[{"reasoningContent":{}},{"reasoningContent":{"reasoningText":{}}},{"reasoningContent":{"reasoningText":{"text":"t","signature":""}}}]}}}',
    b'{"output":{"message":{"content":[{"reasoningContent":{"reasoningText":{"text":"first"}}},{"text":"x"},{"reasoningContent":{"reasoningText":{"text":"last","signature":"s"}}}]}}}',
    b'{"output":{"message":{"content":[{"reasoningContent":{"redactedContent":""}}]}}}',
    b'{"output":{"message":{"content":[{"text":""}],"role":"assistant"}},"unknown":{"a":[1,2,{"b":null}]},"stopReason":"tool_use"}',
    b' {"output" : {"message" : {"content" : [ {"text" : "spaced"} ] , "role" : "assistant"} } , "usage" : {"inputTokens" : 1 , "outputTokens" : 2 , "totalTokens" : 3} } ',
    b'{"output":{"message":{"content":[{"text":"x"}]}},"serviceTier":{"type":"priority"},"usage":{"inputTokens":3,"outputTokens":4,"totalTokens":7,"cacheWriteInputTokens":5}}',
    # decode errors (AIGW_INTERNAL)
    b'{"output":{"message":{"content":[{"text":5}]}}}', b'{"output":{"message":{"content":{"text":"x"}}}}', b'{"output":[]}', b'{"output":{"message":{"role":7}}}',
    b'{"output":{"message":{}},"usage":{"inputTokens":"1"}}', b'{"output":{"message":{}},"usage":{"inputTokens":1.5}}', b'{"output":{"message":{}},"stopReason":1}',
    b'{"output":{"message":{"content":[{"toolUse":{"input":[1]}}]}}}', b'{"output":{"message":{"content":[{"toolUse":"x"}]}}}', b'{"output":{"message":{}},"metrics":{"latencyMs":"1"}}',
    b'{"output":{"message":{"content":[{"reasoningContent":{"reasoningText":{"text":1}}}]}}}', b'{"output":{"message":{}},"serviceTier":{"type":1}}', b'[1]', b'"x"',
    # left to the stock path (AIGW_DECLINED)
    b'{"stopReason":"end_turn"}', b'{"output":null}', b'null', b'{"output":{"message":{"content":[null]}}}', b'{"output":{"message":{"content":[{"image":{"format":"png","source":{"bytes":"AAAA"}}}]}}}',
    b'{"output":{"message":{"content":[{"reasoningContent":{"redactedContent":"AAAA"}}]}}}', b'{"output":{"message":{"content":[{"text":"caf\\u00e9"}]}}}',
    b'{"output":{"message":{"content":[{"cachePoint":{"type":"default"}}]}}}', b'{"output":{"message":{}},"usage":{"inputTokens":4294967296}}',
]
