// Schema table of openai.ChatCompletionResponseChunk for the typed JSON walker (tjson.cuh); shared by the whole-stream usage
// scan (sse_kernel.cu) and the per-chunk stream steps (stream_kernel.cu).  Each translation unit keeps its own device copy.
#pragma once
#include <cstring>

#include "tjson.cuh"

namespace aigw {
using namespace tj;

// ------------------------------------------------------------------ schema: ChatCompletionResponseChunk
// (internal/apischema/openai/openai.go:1497-1565 chunk/choice/delta, :2064-2083 Usage, :1464-1495 details,
//  :1323-1361 logprobs, :1424-1443 annotations, :1875-1879 StreamReasoningContent, :1789-1807 created)
enum Cap : uint8_t { C_PROMPT = 0, C_COMPLETION = 1, C_TOTAL = 2, C_REASONING = 3, C_CACHED = 4, C_CACHE_CREATION = 5,
                     C_OBJ_USAGE = 0, C_OBJ_CTD = 1, C_OBJ_PTD = 2, C_SPAN_MODEL = 0, NOCAP = 0xff };
enum N : uint8_t { N_ANY = 0, N_STR, N_INT, N_FLOAT, N_ROOT, N_CHOICES, N_CHOICE, N_DELTA, N_TOOLCALLS, N_TOOLCALL, N_FUNC, N_ANNOTS, N_ANNOT, N_URLCIT,
                   N_REASON, N_B64, N_LOGPROBS, N_TOKLPS, N_TOKLP, N_INTS, N_TOPLPS, N_TOPLP, N_USAGE, N_CTD, N_PTD, N_CREATED, N_MODEL,
                   N_PROMPT, N_COMPLETION, N_TOTAL, N_REASONING_TOK, N_CACHED, N_CACHE_CREATION,
                   N_CM_CHOICES, N_CM_CHOICE, N_CM_LOGPROBS, N_STRS, N_FLOATS, N_CM_TOPS, N_MAPF, N_COUNT };

struct FieldDef { uint8_t owner; const char* key; uint8_t node; };
static const FieldDef kFields[] = {
  {N_ROOT, "id", N_STR}, {N_ROOT, "choices", N_CHOICES}, {N_ROOT, "created", N_CREATED}, {N_ROOT, "model", N_MODEL}, {N_ROOT, "service_tier", N_STR},
  {N_ROOT, "system_fingerprint", N_STR}, {N_ROOT, "object", N_STR}, {N_ROOT, "usage", N_USAGE}, {N_ROOT, "obfuscation", N_STR},
  {N_CHOICE, "index", N_INT}, {N_CHOICE, "delta", N_DELTA}, {N_CHOICE, "logprobs", N_LOGPROBS}, {N_CHOICE, "finish_reason", N_STR},
  {N_DELTA, "content", N_STR}, {N_DELTA, "role", N_STR}, {N_DELTA, "tool_calls", N_TOOLCALLS}, {N_DELTA, "annotations", N_ANNOTS}, {N_DELTA, "reasoning_content", N_REASON},
  {N_TOOLCALL, "index", N_INT}, {N_TOOLCALL, "id", N_STR}, {N_TOOLCALL, "function", N_FUNC}, {N_TOOLCALL, "type", N_STR},
  {N_FUNC, "arguments", N_STR}, {N_FUNC, "name", N_STR},
  {N_ANNOT, "type", N_STR}, {N_ANNOT, "url_citation", N_URLCIT},
  {N_URLCIT, "end_index", N_INT}, {N_URLCIT, "start_index", N_INT}, {N_URLCIT, "url", N_STR}, {N_URLCIT, "title", N_STR},
  {N_REASON, "text", N_STR}, {N_REASON, "signature", N_STR}, {N_REASON, "redactedContent", N_B64},
  {N_LOGPROBS, "content", N_TOKLPS}, {N_LOGPROBS, "refusal", N_TOKLPS},
  {N_TOKLP, "token", N_STR}, {N_TOKLP, "bytes", N_INTS}, {N_TOKLP, "logprob", N_FLOAT}, {N_TOKLP, "top_logprobs", N_TOPLPS},
  {N_TOPLP, "token", N_STR}, {N_TOPLP, "bytes", N_INTS}, {N_TOPLP, "logprob", N_FLOAT},
  {N_USAGE, "prompt_tokens", N_PROMPT}, {N_USAGE, "completion_tokens", N_COMPLETION}, {N_USAGE, "total_tokens", N_TOTAL},
  {N_USAGE, "completion_tokens_details", N_CTD}, {N_USAGE, "prompt_tokens_details", N_PTD},
  {N_CTD, "text_tokens", N_INT}, {N_CTD, "accepted_prediction_tokens", N_INT}, {N_CTD, "audio_tokens", N_INT}, {N_CTD, "reasoning_tokens", N_REASONING_TOK}, {N_CTD, "rejected_prediction_tokens", N_INT},
  {N_PTD, "text_tokens", N_INT}, {N_PTD, "audio_tokens", N_INT}, {N_PTD, "cached_tokens", N_CACHED}, {N_PTD, "cache_creation_input_tokens", N_CACHE_CREATION},
};
static constexpr int kNumFields = sizeof(kFields) / sizeof(kFields[0]);

struct alignas(16) SchemaBlob {
  Node nodes[N_COUNT];
  Field fields[64];
  char keys[768];
};
static_assert(kNumFields <= 64, "field table too small");

static inline SchemaBlob build_schema() {
  SchemaBlob b; memset(&b, 0, sizeof b);
  auto set = [&](int n, uint8_t kind, uint8_t cap = NOCAP, uint8_t elem = 0) { b.nodes[n].kind = kind; b.nodes[n].cap = cap; b.nodes[n].elem = elem; };
  set(N_ANY, K_ANY); set(N_STR, K_STR); set(N_INT, K_INT); set(N_FLOAT, K_FLOAT);
  set(N_ROOT, K_OBJ); set(N_CHOICES, K_ARR, NOCAP, N_CHOICE); set(N_CHOICE, K_OBJ); set(N_DELTA, K_OBJ);
  set(N_TOOLCALLS, K_ARR, NOCAP, N_TOOLCALL); set(N_TOOLCALL, K_OBJ); set(N_FUNC, K_OBJ);
  set(N_ANNOTS, K_ARR, NOCAP, N_ANNOT); set(N_ANNOT, K_OBJ); set(N_URLCIT, K_OBJ); set(N_REASON, K_OBJ); set(N_B64, K_B64);
  set(N_LOGPROBS, K_OBJ); set(N_TOKLPS, K_ARR, NOCAP, N_TOKLP); set(N_TOKLP, K_OBJ); set(N_INTS, K_ARR, NOCAP, N_INT);
  set(N_TOPLPS, K_ARR, NOCAP, N_TOPLP); set(N_TOPLP, K_OBJ);
  set(N_USAGE, K_OBJ, C_OBJ_USAGE); set(N_CTD, K_OBJ, C_OBJ_CTD); set(N_PTD, K_OBJ, C_OBJ_PTD);
  set(N_CREATED, K_CREATED); set(N_MODEL, K_STR, C_SPAN_MODEL);
  set(N_PROMPT, K_INT, C_PROMPT); set(N_COMPLETION, K_INT, C_COMPLETION); set(N_TOTAL, K_INT, C_TOTAL);
  set(N_REASONING_TOK, K_INT, C_REASONING); set(N_CACHED, K_INT, C_CACHED); set(N_CACHE_CREATION, K_INT, C_CACHE_CREATION);
  int ko = 0;
  for (int f = 0; f < kNumFields; f++) {
    const FieldDef& d = kFields[f];
    Node& o = b.nodes[d.owner];
    if (o.nf == 0) o.f0 = (uint8_t)f;   // fields of one owner are contiguous in kFields
    o.nf++;
    int kl = (int)strlen(d.key);
    b.fields[f].koff = (uint16_t)ko; b.fields[f].klen = (uint8_t)kl; b.fields[f].node = d.node;
    memcpy(b.keys + ko, d.key, kl); ko += kl;
  }
  return b;
}

// ------------------------------------------------------------------ schema: CompletionResponse (legacy /v1/completions; the stream
// chunks and the buffered body share the shape: internal/apischema/openai/openai.go:2000-2055; Usage :2064-2081)
static const FieldDef kCmplFields[] = {
  {N_ROOT, "id", N_STR}, {N_ROOT, "object", N_STR}, {N_ROOT, "created", N_CREATED}, {N_ROOT, "model", N_MODEL}, {N_ROOT, "system_fingerprint", N_STR},
  {N_ROOT, "choices", N_CM_CHOICES}, {N_ROOT, "usage", N_USAGE},
  {N_CM_CHOICE, "text", N_STR}, {N_CM_CHOICE, "index", N_INT}, {N_CM_CHOICE, "logprobs", N_CM_LOGPROBS}, {N_CM_CHOICE, "finish_reason", N_STR},
  {N_CM_LOGPROBS, "tokens", N_STRS}, {N_CM_LOGPROBS, "token_logprobs", N_FLOATS}, {N_CM_LOGPROBS, "top_logprobs", N_CM_TOPS}, {N_CM_LOGPROBS, "text_offset", N_INTS},
  {N_USAGE, "prompt_tokens", N_PROMPT}, {N_USAGE, "completion_tokens", N_COMPLETION}, {N_USAGE, "total_tokens", N_TOTAL},
  {N_USAGE, "completion_tokens_details", N_CTD}, {N_USAGE, "prompt_tokens_details", N_PTD},
  {N_CTD, "text_tokens", N_INT}, {N_CTD, "accepted_prediction_tokens", N_INT}, {N_CTD, "audio_tokens", N_INT}, {N_CTD, "reasoning_tokens", N_REASONING_TOK}, {N_CTD, "rejected_prediction_tokens", N_INT},
  {N_PTD, "text_tokens", N_INT}, {N_PTD, "audio_tokens", N_INT}, {N_PTD, "cached_tokens", N_CACHED}, {N_PTD, "cache_creation_input_tokens", N_CACHE_CREATION},
};
static constexpr int kNumCmplFields = sizeof(kCmplFields) / sizeof(kCmplFields[0]);
static_assert(kNumCmplFields <= 64, "field table too small");

static inline SchemaBlob build_completion_schema() {
  SchemaBlob b; memset(&b, 0, sizeof b);
  auto set = [&](int n, uint8_t kind, uint8_t cap = NOCAP, uint8_t elem = 0) { b.nodes[n].kind = kind; b.nodes[n].cap = cap; b.nodes[n].elem = elem; };
  set(N_ANY, K_ANY); set(N_STR, K_STR); set(N_INT, K_INT); set(N_FLOAT, K_FLOAT);
  set(N_ROOT, K_OBJ); set(N_CM_CHOICES, K_ARR, NOCAP, N_CM_CHOICE); set(N_CM_CHOICE, K_OBJ); set(N_CM_LOGPROBS, K_OBJ);
  set(N_STRS, K_ARR, NOCAP, N_STR); set(N_FLOATS, K_ARR, NOCAP, N_FLOAT); set(N_INTS, K_ARR, NOCAP, N_INT);
  set(N_CM_TOPS, K_ARR, NOCAP, N_MAPF); set(N_MAPF, K_MAP, NOCAP, N_FLOAT);   // []map[string]float64
  set(N_USAGE, K_OBJ, C_OBJ_USAGE); set(N_CTD, K_OBJ, C_OBJ_CTD); set(N_PTD, K_OBJ, C_OBJ_PTD);
  set(N_CREATED, K_CREATED); set(N_MODEL, K_STR, C_SPAN_MODEL);
  set(N_PROMPT, K_INT, C_PROMPT); set(N_COMPLETION, K_INT, C_COMPLETION); set(N_TOTAL, K_INT, C_TOTAL);
  set(N_REASONING_TOK, K_INT, C_REASONING); set(N_CACHED, K_INT, C_CACHED); set(N_CACHE_CREATION, K_INT, C_CACHE_CREATION);
  int ko = 0;
  for (int f = 0; f < kNumCmplFields; f++) {
    const FieldDef& d = kCmplFields[f];
    Node& o = b.nodes[d.owner];
    if (o.nf == 0) o.f0 = (uint8_t)f;   // fields of one owner are contiguous in kCmplFields
    o.nf++;
    int kl = (int)strlen(d.key);
    b.fields[f].koff = (uint16_t)ko; b.fields[f].klen = (uint8_t)kl; b.fields[f].node = d.node;
    memcpy(b.keys + ko, d.key, kl); ko += kl;
  }
  return b;
}

// ------------------------------------------------------------------ R1: non-stream ChatCompletionResponse usage
// json.NewDecoder(body).Decode(&openai.ChatCompletionResponse{}) then usage / model
// (internal/translator/openai_openai.go:146-174; struct internal/apischema/openai/openai.go:1269-1306,1365-1422).
enum RN : uint8_t { R_ANY = 0, R_STR, R_INT, R_FLOAT, R_ROOT, R_CHOICES, R_CHOICE, R_MSG, R_TOOLCALLS, R_TOOLCALL, R_FUNC, R_CACHE, R_ANNOTS, R_ANNOT, R_URLCIT,
                  R_AUDIO, R_REASON_U, R_REASON_O, R_REASON_BLOCK, R_REASON_TEXT, R_B64, R_THINKS, R_THINK, R_LOGPROBS, R_TOKLPS, R_TOKLP, R_INTS, R_TOPLPS, R_TOPLP,
                  R_USAGE, R_CTD, R_PTD, R_CREATED, R_MODEL, R_PROMPT, R_COMPLETION, R_TOTAL, R_REASONING_TOK, R_CACHED, R_CACHE_CREATION,
                  R_EM_DATA, R_EM_ITEM, R_EM_VEC, R_EM_FLOATS, R_EM_USAGE, R_COUNT };
static const FieldDef kRespFields[] = {
  {R_ROOT, "id", R_STR}, {R_ROOT, "choices", R_CHOICES}, {R_ROOT, "created", R_CREATED}, {R_ROOT, "model", R_MODEL}, {R_ROOT, "service_tier", R_STR},
  {R_ROOT, "system_fingerprint", R_STR}, {R_ROOT, "object", R_STR}, {R_ROOT, "usage", R_USAGE}, {R_ROOT, "obfuscation", R_STR},
  {R_CHOICE, "finish_reason", R_STR}, {R_CHOICE, "index", R_INT}, {R_CHOICE, "logprobs", R_LOGPROBS}, {R_CHOICE, "message", R_MSG},
  {R_MSG, "content", R_STR}, {R_MSG, "role", R_STR}, {R_MSG, "tool_calls", R_TOOLCALLS}, {R_MSG, "annotations", R_ANNOTS}, {R_MSG, "audio", R_AUDIO},
  {R_MSG, "reasoning_content", R_REASON_U}, {R_MSG, "thinking_blocks", R_THINKS}, {R_MSG, "safety_ratings", R_ANY}, {R_MSG, "grounding_metadata", R_ANY},
  {R_TOOLCALL, "id", R_STR}, {R_TOOLCALL, "function", R_FUNC}, {R_TOOLCALL, "type", R_STR}, {R_TOOLCALL, "cache_control", R_CACHE},
  {R_FUNC, "arguments", R_STR}, {R_FUNC, "name", R_STR},
  {R_CACHE, "type", R_STR}, {R_CACHE, "ttl", R_STR},
  {R_ANNOT, "type", R_STR}, {R_ANNOT, "url_citation", R_URLCIT},
  {R_URLCIT, "end_index", R_INT}, {R_URLCIT, "start_index", R_INT}, {R_URLCIT, "url", R_STR}, {R_URLCIT, "title", R_STR},
  {R_AUDIO, "data", R_STR}, {R_AUDIO, "expires_at", R_INT}, {R_AUDIO, "id", R_STR}, {R_AUDIO, "transcript", R_STR},
  {R_REASON_O, "reasoningContent", R_REASON_BLOCK},
  {R_REASON_BLOCK, "reasoningText", R_REASON_TEXT}, {R_REASON_BLOCK, "redactedContent", R_B64},
  {R_REASON_TEXT, "text", R_STR}, {R_REASON_TEXT, "signature", R_STR},
  {R_THINK, "type", R_STR}, {R_THINK, "thinking", R_STR}, {R_THINK, "signature", R_STR}, {R_THINK, "data", R_STR},
  {R_LOGPROBS, "content", R_TOKLPS}, {R_LOGPROBS, "refusal", R_TOKLPS},
  {R_TOKLP, "token", R_STR}, {R_TOKLP, "bytes", R_INTS}, {R_TOKLP, "logprob", R_FLOAT}, {R_TOKLP, "top_logprobs", R_TOPLPS},
  {R_TOPLP, "token", R_STR}, {R_TOPLP, "bytes", R_INTS}, {R_TOPLP, "logprob", R_FLOAT},
  {R_USAGE, "prompt_tokens", R_PROMPT}, {R_USAGE, "completion_tokens", R_COMPLETION}, {R_USAGE, "total_tokens", R_TOTAL},
  {R_USAGE, "completion_tokens_details", R_CTD}, {R_USAGE, "prompt_tokens_details", R_PTD},
  {R_CTD, "text_tokens", R_INT}, {R_CTD, "accepted_prediction_tokens", R_INT}, {R_CTD, "audio_tokens", R_INT}, {R_CTD, "reasoning_tokens", R_REASONING_TOK}, {R_CTD, "rejected_prediction_tokens", R_INT},
  {R_PTD, "text_tokens", R_INT}, {R_PTD, "audio_tokens", R_INT}, {R_PTD, "cached_tokens", R_CACHED}, {R_PTD, "cache_creation_input_tokens", R_CACHE_CREATION},
};
static constexpr int kNumRespFields = sizeof(kRespFields) / sizeof(kRespFields[0]);
struct alignas(16) RespSchemaBlob { Node nodes[R_COUNT]; Field fields[80]; char keys[1024]; };
static_assert(kNumRespFields <= 80, "response field table too small");

static inline RespSchemaBlob build_resp_schema() {
  RespSchemaBlob b; memset(&b, 0, sizeof b);
  auto set = [&](int n, uint8_t kind, uint8_t cap = NOCAP, uint8_t elem = 0) { b.nodes[n].kind = kind; b.nodes[n].cap = cap; b.nodes[n].elem = elem; };
  set(R_ANY, K_ANY); set(R_STR, K_STR); set(R_INT, K_INT); set(R_FLOAT, K_FLOAT);
  set(R_ROOT, K_OBJ); set(R_CHOICES, K_ARR, NOCAP, R_CHOICE); set(R_CHOICE, K_OBJ); set(R_MSG, K_OBJ);
  set(R_TOOLCALLS, K_ARR, NOCAP, R_TOOLCALL); set(R_TOOLCALL, K_OBJ); set(R_FUNC, K_OBJ); set(R_CACHE, K_OBJ);
  set(R_ANNOTS, K_ARR, NOCAP, R_ANNOT); set(R_ANNOT, K_OBJ); set(R_URLCIT, K_OBJ); set(R_AUDIO, K_OBJ);
  set(R_REASON_U, K_STROBJ, NOCAP, R_REASON_O); set(R_REASON_O, K_OBJ); set(R_REASON_BLOCK, K_OBJ); set(R_REASON_TEXT, K_OBJ); set(R_B64, K_B64);
  set(R_THINKS, K_ARR, NOCAP, R_THINK); set(R_THINK, K_OBJ);
  set(R_LOGPROBS, K_OBJ); set(R_TOKLPS, K_ARR, NOCAP, R_TOKLP); set(R_TOKLP, K_OBJ); set(R_INTS, K_ARR, NOCAP, R_INT); set(R_TOPLPS, K_ARR, NOCAP, R_TOPLP); set(R_TOPLP, K_OBJ);
  set(R_USAGE, K_OBJ, C_OBJ_USAGE); set(R_CTD, K_OBJ, C_OBJ_CTD); set(R_PTD, K_OBJ, C_OBJ_PTD);
  set(R_CREATED, K_CREATED); set(R_MODEL, K_STR, C_SPAN_MODEL);
  set(R_PROMPT, K_INT, C_PROMPT); set(R_COMPLETION, K_INT, C_COMPLETION); set(R_TOTAL, K_INT, C_TOTAL);
  set(R_REASONING_TOK, K_INT, C_REASONING); set(R_CACHED, K_INT, C_CACHED); set(R_CACHE_CREATION, K_INT, C_CACHE_CREATION);
  int ko = 0;
  for (int f = 0; f < kNumRespFields; f++) {
    const FieldDef& d = kRespFields[f];
    Node& o = b.nodes[d.owner];
    if (o.nf == 0) o.f0 = (uint8_t)f;
    o.nf++;
    int kl = (int)strlen(d.key);
    b.fields[f].koff = (uint16_t)ko; b.fields[f].klen = (uint8_t)kl; b.fields[f].node = d.node;
    memcpy(b.keys + ko, d.key, kl); ko += kl;
  }
  return b;
}


}  // namespace aigw
