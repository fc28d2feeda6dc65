#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into JSON fixtures.

Run in the authoring container (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/extract_goldens.py

Sources (all read-only, nothing is executed):
  * tests/data-plane/testupstream_test.go  — TestWithTestUpstream / TestStreamingUsageInclusionWithCosts
    table entries: (name, backend, path, requestBody, expRequestBody, expPath, responseBody,
    responseType, expResponseBody, expStatus).  The fake upstream compares expRequestBody with
    bytes.Equal (tests/internal/testupstreamlib/server.go:186-196), so these are byte-exact pins.
  * internal/translator/openai_openai_test.go:424-480 — TestExtractUsageFromBufferEvent vectors
    (transcribed below as data; the Go table is not regular enough to scrape).
  * internal/llmcostcel/cel_test.go:28-77 — cost vectors (transcribed).

Only string literals are lifted; no reference source code is copied.
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def go_string_literals(src: str):
    """Yield (start, end, value) for every Go string literal in src (raw `..` and "..")."""
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c == "/" and src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if c == "`":
            j = src.index("`", i + 1)
            yield (i, j + 1, src[i + 1 : j].replace("\r", ""))
            i = j + 1
            continue
        if c == '"':
            j = i + 1
            buf = []
            while src[j] != '"':
                if src[j] == "\\":
                    e = src[j + 1]
                    m = {"n": "\n", "t": "\t", "r": "\r", '"': '"', "\\": "\\", "'": "'", "a": "\a", "b": "\b", "f": "\f", "v": "\v"}
                    if e in m:
                        buf.append(m[e]); j += 2
                    elif e == "x":
                        buf.append(chr(int(src[j + 2 : j + 4], 16))); j += 4
                    elif e == "u":
                        buf.append(chr(int(src[j + 2 : j + 6], 16))); j += 6
                    else:
                        raise ValueError("escape " + e)
                else:
                    buf.append(src[j]); j += 1
            yield (i, j + 1, "".join(buf))
            i = j + 1
            continue
        if c == "'":
            j = i + 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            i = j + 1
            continue
        i += 1


FIELDS = ["name", "backend", "path", "method", "requestBody", "responseBody", "responseType", "responseStatus",
          "responseHeaders", "expRawQuery", "expPath", "expHost", "expRequestBody", "expStatus", "expResponseBody"]


def extract_cases(src: str, consts: dict):
    lits = list(go_string_literals(src))
    starts = {s: (e, v) for s, e, v in lits}
    lit_spans = [(s, e) for s, e, _ in lits]

    def in_literal(pos):
        for s, e in lit_spans:
            if s < pos < e:
                return True
        return False

    field_re = re.compile(r"\b(" + "|".join(FIELDS) + r"):\s*")
    cases, cur = [], None
    for m in field_re.finditer(src):
        if in_literal(m.start()):
            continue
        f, pos = m.group(1), m.end()
        val = None
        if pos in starts:
            val = starts[pos][1]
        else:
            ident = re.match(r"[A-Za-z_][A-Za-z0-9_.]*", src[pos:])
            if ident and ident.group(0) in consts:
                val = consts[ident.group(0)]
            else:
                mm = re.match(r"strconv\.Itoa\(http\.(\w+)\)|http\.(\w+)", src[pos:])
                if mm:
                    val = "http." + (mm.group(1) or mm.group(2))
        if f == "name":
            cur = {"name": val}
            cases.append(cur)
        elif cur is not None and val is not None and f not in cur:
            cur[f] = val
    return cases


def main():
    path = os.path.join(REF, "tests/data-plane/testupstream_test.go")
    src = open(path, encoding="utf-8").read()
    consts = {}
    m = re.search(r"toolCallResultsRequestBody = `", src)
    s = m.end()
    consts["toolCallResultsRequestBody"] = src[s : src.index("`", s)]
    cases = [c for c in extract_cases(src, consts) if c.get("name") and ("requestBody" in c)]
    out = {"source": "tests/data-plane/testupstream_test.go @ reference commit 7a303f93", "cases": cases}
    with open(os.path.join(OUT, "testupstream_cases.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    n_req = sum(1 for c in cases if "expRequestBody" in c)
    print(f"{len(cases)} cases, {n_req} with expRequestBody", file=sys.stderr)

    # internal/translator/openai_openai_test.go:424-480 (values = tokenUsageFrom(in,cached,cacheCreation,out,total,reasoning); -1 ⇒ unset)
    usage = {
        "source": "internal/translator/openai_openai_test.go:424-480 TestExtractUsageFromBufferEvent",
        "cases": [
            {"name": "valid usage data", "feeds": ["data: {\"usage\": {\"total_tokens\": 42}}\n"], "exp": [[0, -1, -1, 0, 42, -1]], "buffered_empty": True},
            {"name": "valid usage data after invalid", "feeds": ["data: invalid\ndata: {\"usage\": {\"total_tokens\": 42}}\n"], "exp": [[0, -1, -1, 0, 42, -1]], "buffered_empty": True},
            {"name": "no usage data and then become valid", "feeds": ["data: {}\n\ndata: ", "{\"usage\": {\"total_tokens\": 42}}\n"],
             "exp": [[-1, -1, -1, -1, -1, -1], [0, -1, -1, 0, 42, -1]], "buffered_empty": True},
            {"name": "valid usage data with cached tokens",
             "feeds": ["data: {\"usage\": {\"prompt_tokens\": 5, \"completion_tokens\": 3, \"total_tokens\": 8, \"prompt_tokens_details\": {\"cached_tokens\": 2, \"cache_creation_input_tokens\": 1}}}\n"],
             "exp": [[5, 2, 1, 3, 8, -1]], "buffered_empty": True},
            {"name": "valid usage data with reasoning tokens",
             "feeds": ["data: {\"usage\": {\"prompt_tokens\": 10, \"completion_tokens\": 20, \"total_tokens\": 30, \"completion_tokens_details\": {\"reasoning_tokens\": 8}}}\n"],
             "exp": [[10, -1, -1, 20, 30, 8]], "buffered_empty": True},
            {"name": "invalid JSON", "feeds": ["data: invalid\n"], "exp": [[-1, -1, -1, -1, -1, -1]], "buffered_empty": True},
        ],
    }
    with open(os.path.join(OUT, "sse_usage_cases.json"), "w", encoding="utf-8") as f:
        json.dump(usage, f, indent=1)


def bedrock_stream():
    """internal/translator/openai_awsbedrock_test.go:1308-1391,1859: a real captured ConverseStream response (AWS eventstream
    frames, base64) and the exact OpenAI SSE text the translator must produce for it (id "123", created 1731679200,
    model "claude-sonnet-4"), fed one byte per call in the reference test."""
    src = open(os.path.join(REF, "internal/translator/openai_awsbedrock_test.go"), encoding="utf-8").read()
    m = re.search(r'const base64RealStreamingEvents = "([^"]+)"', src)
    i = src.index("normalizedResults = append(normalizedResults, []byte(\"data: [DONE]")
    j = src.index("`", i)
    k = src.index("`", j + 1)
    out = {"source": "internal/translator/openai_awsbedrock_test.go:1308-1391,1859", "eventstream_base64": m.group(1), "expected_sse": src[j + 1:k],
           "response_id": "123", "created": 1731679200, "request_model": "claude-sonnet-4", "expect_output_tokens": 75, "expect_input_tokens": 386}
    with open(os.path.join(OUT, "bedrock_stream_real.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1)


def completions_cases():
    """internal/translator/openai_completions_test.go: the five buffered bodies of TestOpenAIToOpenAITranslatorV1CompletionResponseBody
    (:107-247) with their expected TokenUsage / model, and the two streaming tests (:249-337).  The raw-string bodies are lifted as they
    stand; the expectations are transcribed as data in tokenUsageFrom order (input, cached, cacheCreation, output, total, reasoning; -1 = unset)."""
    src = open(os.path.join(REF, "internal/translator/openai_completions_test.go"), encoding="utf-8").read()
    a = src.index("func TestOpenAIToOpenAITranslatorV1CompletionResponseBody(")
    b = src.index("func TestOpenAIToOpenAITranslatorV1CompletionResponseBodyStreaming(")
    c = src.index("func TestOpenAIToOpenAITranslatorV1CompletionResponseError(")
    raw = [v for (_, _, v) in go_string_literals(src[a:b]) if v.lstrip().startswith("{") or v == "invalid json"]
    names = ["valid_response", "valid_response_with_cached_tokens", "valid_response_with_reasoning_tokens", "invalid_json", "response_without_usage"]
    exp = [[5, -1, -1, 8, 13, -1], [5, 2, 1, 8, 13, -1], [5, -1, -1, 8, 13, 3], None, [-1, -1, -1, -1, -1, -1]]
    assert len(raw) == 5, len(raw)
    buffered = [{"name": n, "body": r, "exp": e, "exp_model": "gpt-3.5-turbo-instruct" if e is not None else None, "exp_error": e is None} for n, r, e in zip(names, raw, exp)]
    chunks = [v for (_, _, v) in go_string_literals(src[b:c]) if v.startswith("data: ")]
    assert len(chunks) == 4, len(chunks)
    streams = [
        {"name": "three ResponseBody calls", "feeds": chunks[:3], "exp": [[-1] * 6, [-1] * 6, [5, -1, -1, 3, 8, -1]], "exp_model": ["gpt-3.5-turbo-instruct"] * 3},
        {"name": "reasoning tokens", "feeds": chunks[3:], "exp": [[5, -1, -1, 3, 8, 2]], "exp_model": ["gpt-3.5-turbo-instruct"]},
    ]
    with open(os.path.join(OUT, "completions_cases.json"), "w", encoding="utf-8") as f:
        json.dump({"source": "internal/translator/openai_completions_test.go:107-337", "buffered": buffered, "streams": streams}, f, indent=1)


def aws_anthropic_messages_real():
    """internal/translator/anthropic_awsanthropic_test.go:237-330: eventstream chunks "extracted from a real streaming response" of Anthropic on
    AWS Bedrock (base64) and the exact Anthropic SSE text the /v1/messages translator must produce for them, fed ONE BYTE PER CALL in the reference test."""
    src = open(os.path.join(REF, "internal/translator/anthropic_awsanthropic_test.go"), encoding="utf-8").read()
    a = src.index("awsBase64Chunks := []string{"); b = src.index("\n\t}\n", a)
    chunks = re.findall(r'"([A-Za-z0-9+/=]+)"', src[a:b])
    i = src.index("require.Equal(t, `event: message_start", b); j = src.index("`", i); k = src.index("`", j + 1)
    assert len(chunks) >= 5
    with open(os.path.join(OUT, "aws_anthropic_messages_real.json"), "w", encoding="utf-8") as f:
        json.dump({"source": "internal/translator/anthropic_awsanthropic_test.go:237-330", "chunks_base64": chunks, "expected_sse": src[j + 1:k], "expect_input_tokens": 10, "expect_output_tokens": 15}, f, indent=1)


if __name__ == "__main__":
    main()
    bedrock_stream()
    completions_cases()
    aws_anthropic_messages_real()
