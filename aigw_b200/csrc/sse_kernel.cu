// See sse_kernel.cuh.  sm_100a only.
#include <cstring>
#include <vector>

#include "sse_kernel.cuh"
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
                   N_PROMPT, N_COMPLETION, N_TOTAL, N_REASONING_TOK, N_CACHED, N_CACHE_CREATION, N_COUNT };

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
__device__ SchemaBlob g_schema;

static SchemaBlob build_schema() {
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

// ------------------------------------------------------------------ kernel
static constexpr int kTile = 8192;       // bytes of a stream staged per warp
static constexpr int kMaxLines = 512;    // line-start slots per tile
static constexpr int kWarps = 4;

struct LineOut { uint32_t seq; uint32_t v[6]; uint32_t flags; uint32_t model_off, model_len; };

__global__ void __launch_bounds__(kWarps * 32) sse_usage_kernel(const __grid_constant__ SseParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  SchemaBlob* sch = (SchemaBlob*)smem;
  for (uint32_t i = threadIdx.x; i < sizeof(SchemaBlob) / 4; i += blockDim.x) ((uint32_t*)sch)[i] = ((const uint32_t*)&g_schema)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t FULLM = 0xffffffffu;
  uint8_t* tile = smem + ((sizeof(SchemaBlob) + 15) & ~15u) + (size_t)warp * (kTile + 16 + kMaxLines * 2);
  uint16_t* lstart = (uint16_t*)(tile + kTile + 16);

  for (;;) {
    uint32_t s = 0;
    if (lane == 0) s = atomicAdd(P.next, 1u);
    s = __shfl_sync(FULLM, s, 0);
    if (s >= P.n_streams) break;
    const uint64_t sb = P.chunk_off[P.chunk_first[s]], se = P.chunk_off[P.chunk_first[s + 1]];
    // per-lane "latest" records, by group: 0 usage(in/out/total) 1 ptd(cached, creation) 2 ctd(reasoning) 3 model
    uint32_t seq_u = 0, seq_p = 0, seq_c = 0, seq_m = 0;  // 0 = none; line ordinal + 1
    uint32_t v_in = 0, v_out = 0, v_tot = 0, v_cached = 0, v_cc = 0, v_reason = 0;
    uint64_t m_off = 0; uint32_t m_len = 0;
    uint32_t status = 0;
    uint32_t line_no = 0;
    uint64_t pos = sb;         // next unread stream byte
    uint32_t carry = 0;        // bytes of an unfinished line at tile[0..carry)
    while (pos < se || carry) {
      // ---- stage the next tile after the carry
      uint32_t room = kTile - carry;
      uint32_t take = (uint32_t)((se - pos) < room ? (se - pos) : room);
      if (take == 0 && pos >= se) break;  // only an unterminated tail is left: the reference never parses it
      {
        // global → shared; source alignment is arbitrary, so align the 16-byte loads on the source
        const uint8_t* g = P.bytes + pos;
        uint32_t head = (uint32_t)((16 - ((uintptr_t)g & 15)) & 15); if (head > take) head = take;
        for (uint32_t i = lane; i < head; i += 32) tile[carry + i] = g[i];
        const uint32_t body = (take - head) & ~15u;
        const uint4* g4 = (const uint4*)(g + head);
        for (uint32_t i = lane; i < (body >> 4); i += 32) {
          uint4 v = __ldg(g4 + i);
          uint8_t* d = tile + carry + head + (i << 4);
          if ((((uintptr_t)d) & 15) == 0) *(uint4*)d = v;
          else { const uint8_t* b = (const uint8_t*)&v; for (int k = 0; k < 16; k++) d[k] = b[k]; }
        }
        for (uint32_t i = head + body + lane; i < take; i += 32) tile[carry + i] = g[i];
      }
      __syncwarp();
      const uint32_t filled = carry + take;
      const uint64_t tile_base = pos - carry;  // stream offset of tile[0]
      pos += take;
      // ---- newline positions
      uint32_t nlines = 0; bool overflow = false;
      for (uint32_t b = 0; b < filled; b += 32) {
        const uint32_t i = b + lane;
        const bool nl = i < filled && tile[i] == '\n';
        const uint32_t m = __ballot_sync(FULLM, nl);
        if (nl) { const uint32_t k = nlines + __popc(m & ((1u << lane) - 1u)); if (k < kMaxLines) lstart[k] = (uint16_t)i; else overflow = true; }
        nlines += __popc(m);
      }
      overflow = __any_sync(FULLM, overflow);
      __syncwarp();
      if (nlines == 0) {
        if (filled >= (uint32_t)kTile) { status = AIGW_DECLINED; break; }  // a single line longer than the tile
        if (pos >= se) break;
        carry = filled; continue;
      }
      if (overflow) { status = AIGW_DECLINED; break; }
      // ---- one lane per complete line
      for (uint32_t l0 = 0; l0 < nlines; l0 += 32) {
        const uint32_t li = l0 + lane;
        if (li < nlines) {
          const int b = li == 0 ? 0 : lstart[li - 1] + 1;
          const int e = lstart[li];
          const uint8_t* q = tile + b; const int n = e - b;
          if (n >= 6 && q[0] == 'd' && q[1] == 'a' && q[2] == 't' && q[3] == 'a' && q[4] == ':' && q[5] == ' ') {
            Capture cp; cp.int_set = 0; cp.obj_seen = 0; cp.span_set = 0; cp.weird = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) cp.ints[k] = 0;
            const bool ok = walk(q + 6, n - 6, sch->nodes, sch->fields, sch->keys, N_ROOT, cp);
            if (cp.weird) status = AIGW_DECLINED;
            if (ok) {
              const uint32_t seq = line_no + li + 1;
              if ((cp.span_set & 1u) && cp.span_len[0] > 0) { seq_m = seq; m_off = tile_base + (uint64_t)b + 6 + cp.span_off[0]; m_len = cp.span_len[0]; }
              if (cp.obj_seen & (1u << C_OBJ_USAGE)) {
                seq_u = seq; v_in = cp.ints[C_PROMPT]; v_out = cp.ints[C_COMPLETION]; v_tot = cp.ints[C_TOTAL];
                if (cp.obj_seen & (1u << C_OBJ_PTD)) { seq_p = seq; v_cached = cp.ints[C_CACHED]; v_cc = cp.ints[C_CACHE_CREATION]; }
                if (cp.obj_seen & (1u << C_OBJ_CTD)) { seq_c = seq; v_reason = cp.ints[C_REASONING]; }
              }
            }
          }
        }
      }
      __syncwarp();
      line_no += nlines;
      // ---- carry the unterminated tail to the front of the tile
      const uint32_t last = (uint32_t)lstart[nlines - 1] + 1;
      const uint32_t rem = filled - last;
      for (uint32_t b = 0; b < rem; b += 32) {
        const uint32_t i = b + lane;
        uint8_t c = 0; if (i < rem) c = tile[last + i];
        __syncwarp();
        if (i < rem) tile[i] = c;
      }
      __syncwarp();
      carry = rem;
      if (pos >= se) break;
    }
    // ---- latest-wins merge across lanes
    status = __reduce_or_sync(FULLM, status);
    aigw_sse_result r; memset(&r, 0, sizeof r);
    auto pick = [&](uint32_t seq) -> int { const uint32_t mx = __reduce_max_sync(FULLM, seq); if (mx == 0) return -1; const uint32_t who = __ballot_sync(FULLM, seq == mx); return __ffs(who) - 1; };
    int w;
    w = pick(seq_u); if (w >= 0) { r.usage.input = __shfl_sync(FULLM, v_in, w); r.usage.output = __shfl_sync(FULLM, v_out, w); r.usage.total = __shfl_sync(FULLM, v_tot, w); r.usage.mask |= 1u | 2u | 4u; }
    w = pick(seq_p); if (w >= 0) { r.usage.cached = __shfl_sync(FULLM, v_cached, w); r.usage.cache_creation = __shfl_sync(FULLM, v_cc, w); r.usage.mask |= 8u | 16u; }
    w = pick(seq_c); if (w >= 0) { r.usage.reasoning = __shfl_sync(FULLM, v_reason, w); r.usage.mask |= 32u; }
    w = pick(seq_m); if (w >= 0) { r.model_off = __shfl_sync(FULLM, m_off, w); r.model_len = __shfl_sync(FULLM, m_len, w); }
    r.status = status;
    if (lane == 0) P.results[s] = r;
  }
}

cudaError_t launch_sse_usage(const SseParams& P, int sm_count, cudaStream_t st) {
  static bool ready = false; static int bps = 1;
  const size_t smem = ((sizeof(SchemaBlob) + 15) & ~15u) + (size_t)kWarps * (kTile + 16 + kMaxLines * 2);
  if (!ready) {
    SchemaBlob b = build_schema();
    cudaError_t e = cudaMemcpyToSymbol(g_schema, &b, sizeof b);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sse_usage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sse_usage_kernel, kWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (bps < 1) bps = 1;
    ready = true;
  }
  long long want = ((long long)P.n_streams + kWarps - 1) / kWarps;
  long long grid = (long long)sm_count * bps;
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  sse_usage_kernel<<<(unsigned)grid, kWarps * 32, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace aigw
