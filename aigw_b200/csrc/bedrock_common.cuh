// Pieces of the Bedrock ConverseStream → OpenAI SSE back-translation shared by the whole-stream kernels
// (bedrock_stream_kernel.cu) and the per-chunk stream steps (stream_kernel.cu): the typed-walk schema of
// awsbedrock.ConverseStreamEvent, the chunk serialiser and small byte helpers.  Everything lives in an anonymous namespace:
// each translation unit has its own copy (no relocatable device code in this build).
#pragma once
#include <cstring>

#include "bedrock_stream_kernel.cuh"
#include "tjson.cuh"

namespace aigw {
using namespace tj;
namespace {

// ------------------------------------------------------------------ schema: awsbedrock.ConverseStreamEvent
// (internal/apischema/awsbedrock/awsbedrock.go:423-503; TokenUsage :352-366, ServiceTier :505-509)
enum Span : uint8_t { S_EVT = 0, S_ROLE, S_STOP, S_TEXT, S_TOOLIN, S_RTEXT, S_RSIG, S_NAME, S_TUID, S_TIER, S_REDACT, S_COUNT };
enum Int : uint8_t { I_IN = 0, I_OUT, I_TOTAL, I_READ, I_WRITE };
enum Obj : uint8_t { O_DELTA = 0, O_DTOOL, O_DREASON, O_USAGE, O_START, O_STOOL, O_TIER };
enum BN : uint8_t { B_ANY = 0, B_STR, B_INT, B_ROOT, B_DELTA, B_DTOOL, B_DREASON, B_USAGE, B_START, B_STOOL, B_TIERO,
                    B_EVT, B_ROLE, B_STOPR, B_TEXT, B_TOOLIN, B_RTEXT, B_RSIG, B_NAME, B_TUID, B_TIER, B_REDACT,
                    B_IN, B_OUT, B_TOT, B_READ, B_WRITE, B_COUNT };
constexpr uint8_t B_NOCAP = 0xff;
using BCap = CaptureT<S_COUNT>;

struct BFieldDef { uint8_t owner; const char* key; uint8_t node; };
const BFieldDef kBedrockFields[] = {
  {B_ROOT, "eventType", B_EVT}, {B_ROOT, "contentBlockIndex", B_INT}, {B_ROOT, "delta", B_DELTA}, {B_ROOT, "role", B_ROLE}, {B_ROOT, "stopReason", B_STOPR},
  {B_ROOT, "usage", B_USAGE}, {B_ROOT, "start", B_START}, {B_ROOT, "serviceTier", B_TIERO},
  {B_DELTA, "text", B_TEXT}, {B_DELTA, "toolUse", B_DTOOL}, {B_DELTA, "reasoningContent", B_DREASON},
  {B_DTOOL, "input", B_TOOLIN},
  {B_DREASON, "text", B_RTEXT}, {B_DREASON, "signature", B_RSIG}, {B_DREASON, "redactedContent", B_REDACT},
  {B_USAGE, "inputTokens", B_IN}, {B_USAGE, "outputTokens", B_OUT}, {B_USAGE, "totalTokens", B_TOT}, {B_USAGE, "cacheReadInputTokens", B_READ}, {B_USAGE, "cacheWriteInputTokens", B_WRITE},
  {B_START, "toolUse", B_STOOL},
  {B_STOOL, "name", B_NAME}, {B_STOOL, "toolUseId", B_TUID},
  {B_TIERO, "type", B_TIER},
};
constexpr int kNumBedrockFields = sizeof(kBedrockFields) / sizeof(kBedrockFields[0]);

struct alignas(16) BedrockSchema { Node nodes[B_COUNT]; Field fields[32]; char keys[320]; };
static_assert(kNumBedrockFields <= 32, "field table too small");

inline BedrockSchema build_schema_bedrock() {
  BedrockSchema b; memset(&b, 0, sizeof b);
  auto set = [&](int n, uint8_t kind, uint8_t cap = B_NOCAP, uint8_t elem = 0) { b.nodes[n].kind = kind; b.nodes[n].cap = cap; b.nodes[n].elem = elem; };
  set(B_ANY, K_ANY); set(B_STR, K_STR); set(B_INT, K_INT);
  set(B_ROOT, K_OBJ); set(B_DELTA, K_OBJ, O_DELTA); set(B_DTOOL, K_OBJ, O_DTOOL); set(B_DREASON, K_OBJ, O_DREASON); set(B_USAGE, K_OBJ, O_USAGE);
  set(B_START, K_OBJ, O_START); set(B_STOOL, K_OBJ, O_STOOL); set(B_TIERO, K_OBJ, O_TIER);
  set(B_EVT, K_STR, S_EVT); set(B_ROLE, K_STR, S_ROLE); set(B_STOPR, K_STR, S_STOP); set(B_TEXT, K_STR, S_TEXT); set(B_TOOLIN, K_STR, S_TOOLIN);
  set(B_RTEXT, K_STR, S_RTEXT); set(B_RSIG, K_STR, S_RSIG); set(B_NAME, K_STR, S_NAME); set(B_TUID, K_STR, S_TUID); set(B_TIER, K_STR, S_TIER); set(B_REDACT, K_B64, S_REDACT);
  set(B_IN, K_INT, I_IN); set(B_OUT, K_INT, I_OUT); set(B_TOT, K_INT, I_TOTAL); set(B_READ, K_INT, I_READ); set(B_WRITE, K_INT, I_WRITE);
  int ko = 0;
  for (int f = 0; f < kNumBedrockFields; f++) {
    const BFieldDef& d = kBedrockFields[f];
    Node& o = b.nodes[d.owner];
    if (o.nf == 0) o.f0 = (uint8_t)f;
    o.nf++;
    int kl = (int)strlen(d.key);
    b.fields[f].koff = (uint16_t)ko; b.fields[f].klen = (uint8_t)kl; b.fields[f].node = d.node;
    memcpy(b.keys + ko, d.key, kl); ko += kl;
  }
  return b;
}

// ------------------------------------------------------------------ output text
enum Kind : uint32_t { KD_NONE = 0, KD_MSGSTART, KD_TEXT, KD_TOOLDELTA, KD_REASON, KD_TOOLSTART, KD_STOP, KD_META, KD_EMPTY, KD_BLOCKSTOP };
enum Flag : uint32_t { F_FINISH_MASK = 3u, F_HAS_READ = 4u, F_HAS_WRITE = 8u };

#define AIGW_LIT(name, text) __device__ const char name[] = text
AIGW_LIT(L_DATA, "data: {");
AIGW_LIT(L_ID, "\"id\":\"");
AIGW_LIT(L_CHOICES, "\"choices\":[");
AIGW_LIT(L_DELTA_OPEN, "{\"index\":0,\"delta\":{");
AIGW_LIT(L_CONTENT_EMPTY, "\"content\":\"\"");
AIGW_LIT(L_CONTENT, "\"content\":\"");
AIGW_LIT(L_ROLE, "\"role\":\"");
AIGW_LIT(L_TOOLCALLS, "\"tool_calls\":[{\"index\":");
AIGW_LIT(L_TC_DELTA_A, ",\"id\":null,\"function\":{\"arguments\":\"");
AIGW_LIT(L_TC_DELTA_B, "\",\"name\":\"\"},\"type\":\"function\"}]}}");
AIGW_LIT(L_TC_START_A, ",\"id\":\"");
AIGW_LIT(L_TC_START_B, "\",\"function\":{\"arguments\":\"\",\"name\":\"");
AIGW_LIT(L_TC_START_C, "\"},\"type\":\"function\"}]}}");
AIGW_LIT(L_REASON, "\"reasoning_content\":{");
AIGW_LIT(L_RTEXT, "\"text\":\"");
AIGW_LIT(L_RSIG, "\"signature\":\"");
AIGW_LIT(L_FINISH, "},\"finish_reason\":\"");
AIGW_LIT(L_CREATED, "],\"created\":");
AIGW_LIT(L_MODEL, ",\"model\":\"");
AIGW_LIT(L_TIER, ",\"service_tier\":\"");
AIGW_LIT(L_OBJECT, ",\"object\":\"chat.completion.chunk\"");
AIGW_LIT(L_USAGE, ",\"usage\":{");
AIGW_LIT(L_PROMPT, "\"prompt_tokens\":");
AIGW_LIT(L_COMPLETION, "\"completion_tokens\":");
AIGW_LIT(L_TOTAL, "\"total_tokens\":");
AIGW_LIT(L_PTD, "\"prompt_tokens_details\":{");
AIGW_LIT(L_CACHED, "\"cached_tokens\":");
AIGW_LIT(L_CACHE_CREATION, "\"cache_creation_input_tokens\":");
AIGW_LIT(L_END, "}\n\n");
AIGW_LIT(L_DONE, "data: [DONE]\n");
AIGW_LIT(L_FR_STOP, "stop");
AIGW_LIT(L_FR_LENGTH, "length");
AIGW_LIT(L_FR_FILTER, "content_filter");
AIGW_LIT(L_FR_TOOLS, "tool_calls");
#define LIT(w, name) (w).lit(name, (uint32_t)sizeof(name) - 1u)

template <bool WRITE>
struct Writer {
  uint8_t* p; uint32_t n;
  __device__ __forceinline__ void ch(char c) { if (WRITE) p[n] = (uint8_t)c; n++; }
  __device__ __forceinline__ void lit(const char* s, uint32_t l) { if (WRITE) for (uint32_t k = 0; k < l; k++) p[n + k] = (uint8_t)s[k]; n += l; }
  __device__ __forceinline__ void raw(const uint8_t* s, uint32_t l) { if (WRITE) for (uint32_t k = 0; k < l; k++) p[n + k] = s[k]; n += l; }
  __device__ inline void dec(unsigned long long v) {
    char b[20]; int k = 0;
    do { b[k++] = (char)('0' + v % 10ull); v /= 10ull; } while (v);
    if (WRITE) for (int t = 0; t < k; t++) p[n + t] = (uint8_t)b[k - 1 - t];
    n += k;
  }
  __device__ inline void sdec(long long v) { if (v < 0) { ch('-'); dec(0ull - (unsigned long long)v); } else dec((unsigned long long)v); }
};

// serializeOpenAIChatCompletionChunk of the chunk convertEvent builds for this record
// (internal/translator/openai_awsbedrock.go:858-1006; field order of openai.ChatCompletionResponseChunk,
// internal/apischema/openai/openai.go:1497-1565).  `src` = first byte of the stream (unused when counting).
template <bool WRITE, class ENV>
__device__ uint32_t emit_chunk(const BedrockRec& r, const ENV& P, const uint8_t* src, uint8_t* dst) {
  Writer<WRITE> w{dst, 0};
  const uint32_t kind = r.kf & 0xffu, flags = r.kf >> 8;
  LIT(w, L_DATA);
  if (P.id_len) { LIT(w, L_ID); w.raw((const uint8_t*)P.id, P.id_len); w.ch('"'); w.ch(','); }
  LIT(w, L_CHOICES);
  auto role_member = [&](bool lead, bool trail) {
    if (!r.role_len) return;
    if (lead) w.ch(',');
    LIT(w, L_ROLE); w.raw(src + r.role_off, r.role_len); w.ch('"');
    if (trail) w.ch(',');
  };
  switch (kind) {
    case KD_MSGSTART:
      LIT(w, L_DELTA_OPEN); LIT(w, L_CONTENT_EMPTY); role_member(true, false); w.ch('}'); w.ch('}'); break;
    case KD_TEXT:
      LIT(w, L_DELTA_OPEN); LIT(w, L_CONTENT); w.raw(src + r.a[0], r.a[1]); w.ch('"'); role_member(true, false); w.ch('}'); w.ch('}'); break;
    case KD_TOOLDELTA:
      LIT(w, L_DELTA_OPEN); role_member(false, true); LIT(w, L_TOOLCALLS); w.dec(r.tool_index); LIT(w, L_TC_DELTA_A); w.raw(src + r.a[0], r.a[1]); LIT(w, L_TC_DELTA_B); break;
    case KD_REASON: {
      LIT(w, L_DELTA_OPEN); role_member(false, true); LIT(w, L_REASON);
      if (r.a[1]) { LIT(w, L_RTEXT); w.raw(src + r.a[0], r.a[1]); w.ch('"'); }
      if (r.a[3]) { if (r.a[1]) w.ch(','); LIT(w, L_RSIG); w.raw(src + r.a[2], r.a[3]); w.ch('"'); }
      w.ch('}'); w.ch('}'); w.ch('}'); break;
    }
    case KD_TOOLSTART:
      LIT(w, L_DELTA_OPEN); role_member(false, true); LIT(w, L_TOOLCALLS); w.dec(r.tool_index); LIT(w, L_TC_START_A); w.raw(src + r.a[2], r.a[3]);
      LIT(w, L_TC_START_B); w.raw(src + r.a[0], r.a[1]); LIT(w, L_TC_START_C); break;
    case KD_STOP: {
      LIT(w, L_DELTA_OPEN); LIT(w, L_CONTENT_EMPTY); role_member(true, false); LIT(w, L_FINISH);
      const uint32_t fr = flags & F_FINISH_MASK;
      if (fr == 0) LIT(w, L_FR_STOP); else if (fr == 1) LIT(w, L_FR_LENGTH); else if (fr == 2) LIT(w, L_FR_FILTER); else LIT(w, L_FR_TOOLS);
      w.ch('"'); w.ch('}'); break;
    }
    default: break;  // KD_META, KD_EMPTY: "choices":[]
  }
  LIT(w, L_CREATED); w.sdec(P.created);
  if (P.model_len) { LIT(w, L_MODEL); w.raw((const uint8_t*)P.model, P.model_len); w.ch('"'); }
  if (kind == KD_META) {
    if (r.a[5]) { LIT(w, L_TIER); w.raw(src + r.a[4], r.a[5]); w.ch('"'); }
    LIT(w, L_OBJECT); LIT(w, L_USAGE);
    bool first = true;
    auto num = [&](const char* k, uint32_t kl, uint32_t v) { if (!v) return; if (!first) w.ch(','); first = false; w.lit(k, kl); w.dec(v); };
    num(L_PROMPT, sizeof(L_PROMPT) - 1, r.a[0]); num(L_COMPLETION, sizeof(L_COMPLETION) - 1, r.a[1]); num(L_TOTAL, sizeof(L_TOTAL) - 1, r.a[0] + r.a[1]);
    if (flags & (F_HAS_READ | F_HAS_WRITE)) {
      if (!first) w.ch(',');
      LIT(w, L_PTD); first = true;
      if (flags & F_HAS_READ) num(L_CACHED, sizeof(L_CACHED) - 1, r.a[2]);
      if (flags & F_HAS_WRITE) num(L_CACHE_CREATION, sizeof(L_CACHE_CREATION) - 1, r.a[3]);
      w.ch('}');
    }
    w.ch('}');
  } else LIT(w, L_OBJECT);
  LIT(w, L_END);
  return w.n;
}

__device__ __forceinline__ uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
__device__ __forceinline__ bool same(const uint8_t* p, uint32_t l, const char* w, uint32_t wl) { if (l != wl) return false; for (uint32_t k = 0; k < l; k++) if (p[k] != (uint8_t)w[k]) return false; return true; }
// a captured JSON string body that encoding/json would re-emit byte for byte (escapes limited to \" \\ \n \r \t, no raw controls)
__device__ inline bool canonical(const uint8_t* p, uint32_t len) {
  for (uint32_t i = 0; i < len; i++) {
    const uint32_t c = p[i];
    if (c < 0x20) return false;
    if (c == '\\') { const uint32_t e = p[i + 1]; if (!(e == '"' || e == '\\' || e == 'n' || e == 'r' || e == 't')) return false; i++; }
  }
  return true;
}

}  // namespace
}  // namespace aigw
