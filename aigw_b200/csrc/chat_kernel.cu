// Index (K1), sort (Ks) and emit (K3) stages of the chat translate pass + the launcher.  See chat_kernel.cuh for the design;
// the schema walk (K2) is in chat_walk.cu.  sm_100a only.
//
//   * tokens carry their byte (tty) and a precomputed key / value id (tkid) that 32 lanes compute in
//     parallel, so the schema walk compares one byte instead of strings;
//   * request bodies are staged into shared memory by 1-D bulk asynchronous copies (bulk.cuh);
//   * the output record is assembled as 16-byte chunks (one chunk per lane, funnel-shifted unaligned
//     reads from shared memory) and stored straight to the arena — no shared-memory image of the output.
#include "chat_internal.cuh"

#include <cstdlib>
#include <mutex>

#include "bulk.cuh"
#include "classify.cuh"

namespace aigw {


// 16 bytes starting at an arbitrary shared-memory address (buffers carry ≥ 4 bytes of slack)
__device__ __forceinline__ uint4 lds16_unaligned(const uint8_t* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  const uint32_t sh = (a & 3u) * 8u;
  const uint32_t* w = (const uint32_t*)(p - (a & 3u));
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
  uint4 r;
  r.x = __funnelshift_r(w0, w1, sh); r.y = __funnelshift_r(w1, w2, sh); r.z = __funnelshift_r(w2, w3, sh); r.w = __funnelshift_r(w3, w4, sh);
  return r;
}

// ------------------------------------------------------------------ the kernels
// Launch sequence per sub-batch of documents, intermediates in a reusable global workspace:
//   K1 index   (warp per document)    stage 1+2+2.5 → token words, token count, histogram of token counts
//   Ks sort    (tiny)                 counting sort of the documents by token count → permutation
//   K2 walk    (thread per document)  stage 3+4 → jump table, copy ops, scratch, plan header
//   K3 emit    (warp per document)    stage 5 → output record + result
// K2 is where SIMT pays: 32 bodies with the same token count (≈ the same shape) share every instruction fetch.
// Sub-batches are small enough for the bodies to stay in L2 between K1 and K3 (the body is read from DRAM once) and
// alternate between two streams, so the latency-bound walk of one sub-batch runs under the issue-bound index / emit
// of the other.
__device__ __forceinline__ aigw_doc_result blank_result(uint32_t len) {
  aigw_doc_result res;
  res.out_off = 0; res.body_len = 0; res.path_len = 0; res.status = AIGW_DECLINED; res.reason = AIGW_R_NONE;
  res.model_off = 0; res.model_len = 0; res.body_kind = AIGW_BODY_UNCHANGED; res.flags = 0; res.in_len = len; res._pad = 0;
  return res;
}
__device__ __forceinline__ uint32_t bin_of(uint32_t w) { return (w & 0x80000000u) ? 0u : (w < (uint32_t)kBins ? w : (uint32_t)kBins - 1u); }

// per-byte equality against a 7-bit constant as a most-significant-bit mask (0x80 per matching byte); z = w & 0x7f7f7f7f
__device__ __forceinline__ uint32_t eq_msb(uint32_t w, uint32_t z, uint32_t c4) { return ~(((z ^ c4) + 0x7f7f7f7fu) | w) & 0x80808080u; }
// bytes below 0x20
__device__ __forceinline__ uint32_t ctl_msb(uint32_t w, uint32_t z) { return ~((z + 0x60606060u) | w) & 0x80808080u; }
// msb mask → 4-bit mask (byte k → bit k): the four products land on bits 28..31 without carries
__device__ __forceinline__ uint32_t nib_from_msb(uint32_t t) { return (t * 0x00204081u) >> 28; }

// A round's sparse work (bytes outside strings, token compaction) is done cooperatively when few lanes own any of it:
// the owning lane's mask is broadcast and lane i handles its bit i, so a 28-byte cluster of punctuation between two
// messages costs one pass of ~12 instructions instead of a 28-trip single-lane loop.

// ---- K1: structural index, one warp per document
template <int MAXD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) chat_index_kernel(const __grid_constant__ ChatParams P, uint32_t doc0, uint32_t ndocs, uint8_t* work) {
  using C = Cls<MAXD>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  // CTA-shared tables: key/value id hash tables
  IdTables* s_ids = (IdTables*)smem;
  uint64_t* s_bar = (uint64_t*)(smem + sizeof(IdTables));
  {
    const uint32_t* src = (const uint32_t*)((P.schema & 48) == AIGW_SCHEMA_RESP_AWS_BEDROCK ? &c_ids_resp : &c_ids);
    for (uint32_t i = threadIdx.x; i < sizeof(IdTables) / 4; i += blockDim.x) ((uint32_t*)s_ids)[i] = src[i];
  }
  if (threadIdx.x == 0) { for (int w = 0; w < WARPS; w++) mbar_init(&s_bar[w], 1); mbar_fence_init(); }
  __syncthreads();
  constexpr int kHdr = (int)((sizeof(IdTables) + 8 * WARPS + 15) & ~(size_t)15);
  constexpr int kNcBytes = ((MAXD / 1024 + 1) * 4 + 15) & ~15;   // one word per round: lanes (32-byte blocks) that hold a valid escape the encoder rewrites
  constexpr int kWarpBytes = C::kIn + C::kTok * 6 + kNcBytes;
  uint8_t* wb = smem + kHdr + (size_t)warp * kWarpBytes;
  uint8_t* s_in = wb;
  uint32_t* s_tw = (uint32_t*)(wb + C::kIn);
  uint16_t* s_bs = (uint16_t*)(s_tw + C::kTok);   // running backslash count in front of each token; later overlaid by the lookup list
  uint32_t* s_nc = (uint32_t*)(wb + C::kIn + C::kTok * 6);
  uint64_t* bar = &s_bar[warp];
  const WorkPtrs wp = carve(work, ndocs, layout_of<MAXD>());
  uint32_t phase = 0;

  // work fetch, one document ahead: the atomic and the location loads of document k+1 complete under the scan of document k
  uint32_t li = 0;
  if (lane == 0) li = atomicAdd(wp.c_index, 1u);
  li = __shfl_sync(FULL, li, 0);
  uint32_t doc = 0, len = 0; const uint8_t* g = nullptr;
  if (li < ndocs) { doc = P.doc_map ? P.doc_map[doc0 + li] : doc0 + li; len = P.lens[doc]; g = P.bodies + P.offsets[doc]; }
  while (li < ndocs) {
    uint32_t li_next = 0;
    if (lane == 0) li_next = atomicAdd(wp.c_index, 1u);
    li_next = __shfl_sync(FULL, li_next, 0);
    uint32_t ndoc = 0, nlen = 0; const uint8_t* ng = nullptr;
    if (li_next < ndocs) { ndoc = P.doc_map ? P.doc_map[doc0 + li_next] : doc0 + li_next; nlen = P.lens[ndoc]; ng = P.bodies + P.offsets[ndoc]; }
#define AIGW_NEXT_DOC() do { li = li_next; doc = ndoc; len = nlen; g = ng; } while (0)
    if (len > (uint32_t)MAXD || len == 0) {
      if (lane == 0) { wp.ntok[li] = 0x80000000u | (len ? AIGW_R_TOO_LARGE : AIGW_R_E400_SYNTAX); atomicAdd(&wp.bins[0], 1u); }
      AIGW_NEXT_DOC();
      continue;
    }
    // ---- stage 1: one bulk asynchronous copy stages the body (no load / store issue slots)
    // (a caller that broke the 16-byte alignment contract gets an ordinary cooperative copy instead of a fault)
    const uint32_t rounds = (len + 1023u) >> 10;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      if (lane == 0) { const uint32_t nbytes = (len + 15u) & ~15u; mbar_expect_tx(bar, nbytes); bulk_g2s(s_in, g, nbytes, bar); }
      mbar_wait(bar, phase); phase ^= 1u;
    } else {
      for (uint32_t i = lane; i < len; i += 32) s_in[i] = g[i];
      __syncwarp();
    }
    // the next document starts its way from DRAM to L2 now (its location was loaded before this one's copy was issued)
    if (lane == 0 && nlen && nlen <= (uint32_t)MAXD && (reinterpret_cast<uintptr_t>(ng) & 15u) == 0) bulk_prefetch_l2(ng, (nlen + 15u) & ~15u);
    // ---- stage 2: structural index
    uint32_t carry_esc = 0, carry_str = 0, carry_sc = 0;
    uint32_t ntok = 0, bs_run = 0;
    uint32_t flags = 0;  // bit0 ctrl in string, bit1 bad escape, bit2 token overflow
    uint32_t any_nc = 0; // some round holds a valid escape the encoder re-spells (warp-uniform)
    for (uint32_t r = 0; r < rounds; r++) {
      const uint32_t rbase = r << 10;
      const uint32_t base = rbase + (lane << 5);
      // bytes at or beyond len (stale shared memory) are masked out of every class
      const uint32_t valid = base + 32u <= len ? 0xffffffffu : (base >= len ? 0u : ((1u << (len - base)) - 1u));
      const uint4 a = *(const uint4*)(s_in + base), b = *(const uint4*)(s_in + base + 16);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      // every class mask of the lane's 32 bytes at once (classify.cuh: 8 bit planes, then LOP3s)
      const ByteClasses bc = classify32(w);
      const uint32_t mq = bc.quote & valid, mb = bc.bslash & valid, mctl = bc.ctl & valid;
      // escaped characters: odd-length backslash runs, carry across lanes
      uint32_t esc = 0, bs_lane = 0;
      const uint32_t any_bs = __ballot_sync(FULL, mb != 0);
      if (any_bs | carry_esc) {
        const uint32_t tail = __clz(~mb);  // backslash run length at the top of this lane's 32 bytes
        const uint32_t odd = __ballot_sync(FULL, tail < 32u && (tail & 1u));
        const uint32_t full = __ballot_sync(FULL, tail == 32u);
        uint32_t cin;
        {
          const uint32_t below = ~full & lt;
          if (below == 0) cin = carry_esc; else cin = (odd >> (31 - __clz(below))) & 1u;
        }
        {
          const uint32_t below = ~full;
          carry_esc = below == 0 ? carry_esc : (odd >> (31 - __clz(below))) & 1u;
        }
        const uint32_t bs = mb & ~cin;  // a leading backslash that is itself escaped starts no run
        const uint32_t follows = (bs << 1) | cin;
        const uint32_t even = 0x55555555u;
        const uint32_t odd_starts = bs & ~even & ~follows;
        const uint32_t seq_even = odd_starts + bs;  // carry-out handled through `tail`
        const uint32_t invert = seq_even << 1;
        esc = (even ^ invert) & follows;
        // running backslash count in front of this lane (tokens remember it: a string has an escape iff the count moves)
        const uint32_t pc = __popc(mb);
        uint32_t incl = pc;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, sft); if (lane >= sft) incl += v; }
        bs_lane = bs_run + incl - pc;
        bs_run += __shfl_sync(FULL, incl, 31);
      } else bs_lane = bs_run;
      esc &= valid;
      const uint32_t uq = mq & ~esc;
      uint32_t ps = uq;
      ps ^= ps << 1; ps ^= ps << 2; ps ^= ps << 4; ps ^= ps << 8; ps ^= ps << 16;
      const uint32_t par = __ballot_sync(FULL, __popc(uq) & 1);
      const uint32_t sin = (__popc(par & lt) & 1u) ^ carry_str;
      if (sin) ps = ~ps;
      carry_str ^= __popc(par) & 1u;
      ps &= valid;
      // ps: bit set from an opening quote up to the byte before its closing quote
      if (mctl & ps) flags |= 1u;
      {  // escaped characters: \" \\ \n \r \t are what the encoder writes back; \/ \b \f \uXXXX are valid but re-spelled (marked per
         // 32-byte block, re-escaped by the walk when the string is echoed); anything else is not JSON
        uint32_t e = esc & ps; bool nc = false;
        while (e) {
          const int j = __ffs(e) - 1; e &= e - 1; const uint32_t c = s_in[base + j];
          if (!(c == '"' || c == '\\' || c == 'n' || c == 'r' || c == 't')) { if (c == '/' || c == 'b' || c == 'f' || c == 'u') nc = true; else flags |= 2u; }
        }
        const uint32_t ncm = __ballot_sync(FULL, nc);
        if (lane == 0) s_nc[r] = ncm;
        any_nc |= ncm;
      }
      // bytes outside strings: structural characters are tokens, everything that is neither structural nor whitespace is a
      // scalar character (a control character other than \t \n \r lands there and fails the scalar grammar in the walk)
      const uint32_t outside = ~ps & ~uq & valid;
      uint32_t mtok = uq | (bc.op & outside);
      const uint32_t msc = outside & ~bc.op & ~bc.ws;
      {
        uint32_t prev = __shfl_up_sync(FULL, msc >> 31, 1);
        if (lane == 0) prev = carry_sc;
        carry_sc = __shfl_sync(FULL, msc >> 31, 31);
        mtok |= msc & ~((msc << 1) | prev);
      }
      // compact token words: position | byte | (scalars) length of the run inside this 32-byte block
      const uint32_t cnt = __popc(mtok);
      uint32_t incl = cnt;
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, sft); if (lane >= sft) incl += v; }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      const uint32_t wpos0 = ntok + incl - cnt;
      if (ntok + total > (uint32_t)C::kTok) flags |= 4u;
      else {
        const uint32_t tl = __ballot_sync(FULL, mtok != 0);
        if (17u * (uint32_t)__popc(tl) <= 13u * __reduce_max_sync(FULL, cnt)) {   // ≈ instructions per owning lane vs per trip of the per-lane loop
          uint32_t m = tl;
          while (m) {
            const int L = __ffs(m) - 1; m &= m - 1;
            const uint32_t mk = __shfl_sync(FULL, mtok, L), wq = __shfl_sync(FULL, wpos0, L), sc = __shfl_sync(FULL, msc, L);
            const uint32_t mbL = __shfl_sync(FULL, mb, L), bsL = __shfl_sync(FULL, bs_lane, L);
            if ((mk >> lane) & 1u) {
              const uint32_t pos = rbase + (L << 5) + lane;
              uint32_t tword = pos | ((uint32_t)s_in[pos] << 16);
              if ((sc >> lane) & 1u) { const uint32_t run = __ffs(~(sc >> lane)) - 1u; if (lane + run < 32u) tword |= run << 24; }  // a run that reaches the block end is measured by the walker
              const uint32_t slot = wq + __popc(mk & lt);
              s_tw[slot] = tword;
              s_bs[slot] = (uint16_t)(bsL + __popc(mbL & lt));
            }
          }
        } else {
          uint32_t m = mtok, wpos = wpos0;
          while (m) {
            const int j = __ffs(m) - 1; m &= m - 1;
            uint32_t tword = (base + j) | ((uint32_t)s_in[base + j] << 16);
            if ((msc >> j) & 1u) { const uint32_t run = __ffs(~(msc >> j)) - 1u; if (j + run < 32u) tword |= run << 24; }
            s_tw[wpos] = tword;
            s_bs[wpos] = (uint16_t)(bs_lane + __popc(mb & ((1u << j) - 1u)));
            wpos++;
          }
        }
      }
      ntok += total;
    }
    flags = __reduce_or_sync(FULL, flags);
    __syncwarp();
    int reason = 0;
    if (flags & 4u) reason = AIGW_R_TOKENS;
    else if (flags & 1u) reason = AIGW_R_CTRL_IN_STRING;
    else if (flags & 2u) reason = AIGW_R_ESCAPE;
    else if (carry_str) reason = AIGW_R_E400_SYNTAX;
    if (reason) { if (lane == 0) { wp.ntok[li] = 0x80000000u | (uint32_t)reason; atomicAdd(&wp.bins[0], 1u); } AIGW_NEXT_DOC(); continue; }
    // ---- stage 2.5: strings get their escape flag and (short, unescaped ones) a key / value id.  Quote tokens alternate
    // open/close, so opening quotes are the quote tokens of even rank.  Pass 1 (one token per lane) collects the strings worth
    // a lookup into a dense list, pass 2 runs the hash lookup with every lane busy.
    uint32_t nlk = 0;
    {
      uint16_t* s_lk = s_bs;   // list entry k ≤ i/2 never overtakes the counts still to be read (same-iteration reads are fenced)
      uint32_t qbase = 0;
      for (uint32_t b0 = 0; b0 < ntok; b0 += 32) {
        const uint32_t i = b0 + lane;
        const uint32_t w = i < ntok ? s_tw[i] : 0u;
        const bool isq = i < ntok && ((w >> 16) & 0xffu) == '"';
        const uint32_t qm = __ballot_sync(FULL, isq);
        bool want = false;
        if (isq && ((qbase + __popc(qm & lt)) & 1u) == 0 && i + 1 < ntok) {
          const uint32_t o = (w & 0xffffu) + 1u, n = (s_tw[i + 1] & 0xffffu) - o;
          const bool esc = s_bs[i + 1] != s_bs[i];
          if (esc) {
            uint32_t ncf = 0, id = 0;
            if (any_nc) {   // only documents that hold a re-spelled escape pay for the per-string questions
              // bit 31: some block the string touches holds a re-spelled escape (conservative: the walk re-checks byte by byte)
              const uint32_t b0k = o >> 5, b1k = (o + n) >> 5;
              for (uint32_t wk = b0k >> 5; wk <= (b1k >> 5); wk++) {
                uint32_t m = s_nc[wk];
                if (wk == (b0k >> 5)) m &= ~0u << (b0k & 31u);
                if (wk == (b1k >> 5)) m &= (b1k & 31u) == 31u ? ~0u : ((2u << (b1k & 31u)) - 1u);
                ncf |= m;
              }
              // ids are looked up on the decoded bytes, so an escaped spelling of a key or of an enumerated value keeps its meaning
              // (a string whose escapes are all canonical cannot equal a table entry: none contains such a character)
              if (ncf && n <= 6u * (uint32_t)kMaxIdLen) { const bool is_key = i + 2 < ntok && ((s_tw[i + 2] >> 16) & 0xffu) == ':'; id = lookup_id_escaped(s_ids, s_in + o, n, is_key); }
            }
            s_tw[i] = w | (1u << 30) | (ncf ? (1u << 31) : 0u) | (id << 24);
          }
          want = !esc && n >= 1u && n <= (uint32_t)kMaxIdLen;
        }
        __syncwarp();
        const uint32_t wm = __ballot_sync(FULL, want);
        if (want) s_lk[nlk + __popc(wm & lt)] = (uint16_t)i;
        nlk += __popc(wm);
        qbase += __popc(qm);
      }
      __syncwarp();
      for (uint32_t k = lane; k < nlk; k += 32) {
        const uint32_t i = s_lk[k];
        const uint32_t w = s_tw[i];
        const uint32_t o = (w & 0xffffu) + 1u, n = (s_tw[i + 1] & 0xffffu) - o;
        const bool is_key = i + 2 < ntok && ((s_tw[i + 2] >> 16) & 0xffu) == ':';
        s_tw[i] = w | (lookup_id(s_ids, s_in + o, n, is_key) << 24);
      }
      __syncwarp();
    }
    // ---- write the token words out, coalesced; histogram of token counts for the shape sort
    {
      uint32_t* gt = wp.tw + (size_t)li * C::kTok;
      for (uint32_t i = lane; i < ntok; i += 32) gt[i] = s_tw[i];
      if (lane == 0) { wp.ntok[li] = ntok; atomicAdd(&wp.bins[bin_of(ntok)], 1u); }
    }
    __syncwarp();
    AIGW_NEXT_DOC();
  }
#undef AIGW_NEXT_DOC
}

// ---- Ks: counting sort by token count (documents with the same count almost always have the same shape).  Every block
// scans the (tiny) histogram itself; positions inside a bin come from a global cursor.
__global__ void __launch_bounds__(1024) chat_sort_kernel(const uint32_t* ntok, const uint32_t* bins, uint32_t* cursor, uint32_t* perm, uint32_t ndocs) {
  __shared__ uint32_t s[kBins];
  const int t = threadIdx.x;
  const uint32_t v = bins[t];
  s[t] = v;
  __syncthreads();
  for (int off = 1; off < kBins; off <<= 1) { const uint32_t x = t >= off ? s[t - off] : 0u; __syncthreads(); s[t] += x; __syncthreads(); }
  const uint32_t excl = s[t] - v;
  __syncthreads();
  s[t] = excl;
  __syncthreads();
  for (uint32_t li = blockIdx.x * blockDim.x + t; li < ndocs; li += gridDim.x * blockDim.x) {
    const uint32_t b = bin_of(ntok[li]);
    perm[s[b] + atomicAdd(&cursor[b], 1u)] = li;
  }
}

// ---- K3: emit, one warp per document.
// The body is staged by one bulk asynchronous copy while the warp loads the ops and scans their lengths.  Output is built
// as 16-byte chunks, one chunk per lane: the op a chunk starts in comes from a scatter (last op starting at or before each
// chunk) + running max scan instead of a search; chunks inside one op are a single funnel-shifted shared-memory read; the
// chunks that straddle op boundaries are collected and assembled afterwards with every lane busy.
static constexpr int kEmitTile = 256;   // chunks per tile of the op-start table
template <int MAXD>
struct EmitSmem {
  using C = Cls<MAXD>;
  static constexpr int kPreB = (C::kOps + 4) * 4, kOpsB = C::kOps * 4, kScrB = (C::kScr + 15) & ~15, kFirstB = kEmitTile * 2, kBlB = (C::kOps + 4) * 4;
  static constexpr int kWarpBytes = C::kIn + kPreB + kOpsB + kScrB + kFirstB + kBlB;
};

template <int MAXD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) chat_emit_kernel(const __grid_constant__ ChatParams P, uint32_t doc0, uint32_t ndocs, uint8_t* work) {
  using C = Cls<MAXD>;
  using E = EmitSmem<MAXD>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint8_t* s_lits = smem;
  uint64_t* s_bar = (uint64_t*)(smem + sizeof(c_lits.bytes));
  for (uint32_t i = threadIdx.x; i < sizeof(c_lits.bytes) / 4; i += blockDim.x) ((uint32_t*)s_lits)[i] = ((const uint32_t*)c_lits.bytes)[i];
  if (threadIdx.x == 0) { for (int w = 0; w < WARPS; w++) mbar_init(&s_bar[w], 1); mbar_fence_init(); }
  __syncthreads();
  constexpr int kHdr = (int)((sizeof(c_lits.bytes) + 8 * WARPS + 15) & ~(size_t)15);
  uint8_t* wb = smem + kHdr + (size_t)warp * E::kWarpBytes;
  uint8_t* s_in = wb;
  uint32_t* s_pre = (uint32_t*)(wb + C::kIn);
  uint32_t* s_ops = s_pre + C::kOps + 4;
  uint8_t* s_scr = (uint8_t*)(s_ops + C::kOps);
  uint16_t* s_first = (uint16_t*)(s_scr + E::kScrB);
  uint32_t* s_bl = (uint32_t*)(s_first + kEmitTile);
  uint64_t* bar = &s_bar[warp];
  const WorkPtrs wp = carve(work, ndocs, layout_of<MAXD>());
  uint32_t phase = 0;

  // work fetch, one document ahead: plan header, location and the first 32 ops of document k+1 are loaded under the emit of
  // document k (each of these loads used to stall the warp for a DRAM round trip at the top of the loop)
  struct Meta { uint32_t doc, len; const uint8_t* g; PlanOut po; uint32_t op0; };
  auto load_meta = [&](uint32_t l) {
    Meta m;
    m.doc = P.doc_map ? P.doc_map[doc0 + l] : doc0 + l;
    m.len = P.lens[m.doc]; m.g = P.bodies + P.offsets[m.doc];
    m.po = wp.plan[l];
    m.op0 = (wp.ops + (size_t)l * (C::kOps + kSysCap))[lane];
    return m;
  };
  uint32_t li = 0;
  if (lane == 0) li = atomicAdd(wp.c_emit, 1u);
  li = __shfl_sync(FULL, li, 0);
  Meta cur{};
  if (li < ndocs) cur = load_meta(li);
  while (li < ndocs) {
    uint32_t li_next = 0;
    if (lane == 0) li_next = atomicAdd(wp.c_emit, 1u);
    li_next = __shfl_sync(FULL, li_next, 0);
    Meta nxt{};
    if (li_next < ndocs) nxt = load_meta(li_next);
    const uint32_t doc = cur.doc, len = cur.len;
    const PlanOut po = cur.po;
    const uint32_t op_first = cur.op0;
    const uint8_t* g = cur.g;
    aigw_doc_result res = blank_result(len);
    if (po.reason) {
      res.reason = po.reason;
      res.status = po.reason >= 48 ? AIGW_INTERNAL : po.reason >= 40 ? AIGW_INVALID_422 : po.reason >= 32 ? AIGW_MALFORMED_400 : AIGW_DECLINED;
      if (lane == 0) P.results[doc] = res;
      li = li_next; cur = nxt;
      continue;
    }
    const uint32_t nops = po.nops, olen = po.olen;
    const bool bulk = (reinterpret_cast<uintptr_t>(g) & 15u) == 0;
    if (bulk) { if (lane == 0) { const uint32_t nbytes = (len + 15u) & ~15u; mbar_expect_tx(bar, nbytes); bulk_g2s(s_in, g, nbytes, bar); } }
    else for (uint32_t i = lane; i < len; i += 32) s_in[i] = g[i];
    // while the body is in flight: ops (+ exclusive prefix of their lengths) and the scratch bytes they reference
    {
      const uint32_t* gops = wp.ops + (size_t)li * (C::kOps + kSysCap);
      uint32_t run = 0, scr_hi = 0;
      for (uint32_t b0 = 0; b0 < nops; b0 += 32) {
        const uint32_t k = b0 + lane;
        const uint32_t op = k < nops ? (b0 == 0 ? op_first : gops[k]) : 0u;
        const uint32_t l = (op >> 16) & 0x3fffu;
        if ((op >> 30) == 2u) scr_hi = max(scr_hi, (op & 0xffffu) + l);
        uint32_t incl = l;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, sft); if (lane >= sft) incl += v; }
        if (k < nops) { s_ops[k] = op; s_pre[k] = run + incl - l; }
        run += __shfl_sync(FULL, incl, 31);
      }
      if (lane == 0) s_pre[nops] = olen;
      scr_hi = __reduce_max_sync(FULL, scr_hi);
      const uint8_t* gs = wp.scr + (size_t)li * C::kScr;
      for (uint32_t i = lane; i < ((scr_hi + 3u) >> 2); i += 32) ((uint32_t*)s_scr)[i] = ((const uint32_t*)gs)[i];
    }
    const uint32_t rec = (olen + 15u) & ~15u;
    const uint32_t nchunks = rec >> 4;
    unsigned long long obase = 0;
    if (lane == 0) obase = atomicAdd(P.out_used, (unsigned long long)rec);
    obase = __shfl_sync(FULL, obase, 0);
    __syncwarp();
    if (bulk) { mbar_wait(bar, phase); phase ^= 1u; }
    if (obase + rec > P.out_capacity || nchunks > 0xffffu) {
      res.reason = nchunks > 0xffffu ? AIGW_R_OUT_SPACE : AIGW_R_ARENA_FULL;
      if (lane == 0) P.results[doc] = res;
      __syncwarp();
      li = li_next; cur = nxt;
      continue;
    }
    if (lane == 0 && li_next < ndocs && !nxt.po.reason && (reinterpret_cast<uintptr_t>(nxt.g) & 15u) == 0) bulk_prefetch_l2(nxt.g, (nxt.len + 15u) & ~15u);
    uint4* o4 = (uint4*)(P.out + obase);
    uint32_t carry = 0, nb = 0;
    for (uint32_t t0 = 0; t0 < nchunks; t0 += kEmitTile) {
      const uint32_t tn = min((uint32_t)kEmitTile, nchunks - t0);
      for (uint32_t i = lane; i < (tn + 1u) >> 1; i += 32) ((uint32_t*)s_first)[i] = 0u;
      __syncwarp();
      // op k is the last op that starts in (16(c-1), 16c] ⇒ it owns the table entry of chunk c
      for (uint32_t k = lane; k < nops; k += 32) {
        const uint32_t c = (s_pre[k] + 15u) >> 4, cn = (s_pre[k + 1] + 15u) >> 4;
        if (c != cn && c - t0 < tn) s_first[c - t0] = (uint16_t)k;
      }
      __syncwarp();
      for (uint32_t i0 = 0; i0 < tn; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < tn;
        uint32_t k = valid ? (uint32_t)s_first[i] : 0u;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) k = max(k, __shfl_up_sync(FULL, k, sft));
        k = max(k, carry);
        carry = __shfl_sync(FULL, k, 31);
        const uint32_t c = t0 + i, pos = c << 4;
        const uint32_t p0 = s_pre[k], p1 = s_pre[k + 1];
        const bool interior = valid && p1 >= pos + 16u;
        if (interior) {
          const uint32_t op = s_ops[k];
          const uint32_t kind = op >> 30, off = op & 0xffffu;
          const uint8_t* sp = (kind == 0 ? s_in : kind == 1 ? s_lits : s_scr) + off + (pos - p0);
          o4[c] = lds16_unaligned(sp);
        }
        const bool edge = valid && !interior;
        const uint32_t bm = __ballot_sync(FULL, edge);
        if (edge) s_bl[nb + __popc(bm & lt)] = c | (k << 16);
        nb += __popc(bm);
      }
      __syncwarp();
    }
    // chunks that straddle op boundaries: OR together one shifted, masked 16-byte read per overlapping op
    for (uint32_t bi = lane; bi < nb; bi += 32) {
      const uint32_t ent = s_bl[bi];
      const uint32_t c = ent & 0xffffu, pos = c << 4;
      uint32_t k = ent >> 16;
      uint32_t op = s_ops[k];
      uint32_t kind = op >> 30, l = (op >> 16) & 0x3fffu, off = op & 0xffffu;
      uint32_t within = pos - s_pre[k];
      const uint8_t* sp = (kind == 0 ? s_in : kind == 1 ? s_lits : s_scr) + off;
      uint4 v = make_uint4(0, 0, 0, 0);
      uint32_t filled = 0;
      for (;;) {
        if (within >= l) {
          if (++k >= nops) break;
          op = s_ops[k]; kind = op >> 30; l = (op >> 16) & 0x3fffu; off = op & 0xffffu; within = 0;
          sp = (kind == 0 ? s_in : kind == 1 ? s_lits : s_scr) + off;
          continue;
        }
        const uint32_t take = min(l - within, 16u - filled);
        uint4 x = lds16_unaligned(sp + within);
        {  // keep the first `take` bytes
          const uint32_t t0 = take, t1 = take > 4u ? take - 4u : 0u, t2 = take > 8u ? take - 8u : 0u, t3 = take > 12u ? take - 12u : 0u;
          x.x &= t0 >= 4u ? 0xffffffffu : ((1u << (8u * t0)) - 1u);
          x.y &= t1 >= 4u ? 0xffffffffu : ((1u << (8u * t1)) - 1u);
          x.z &= t2 >= 4u ? 0xffffffffu : ((1u << (8u * t2)) - 1u);
          x.w &= t3 >= 4u ? 0xffffffffu : ((1u << (8u * t3)) - 1u);
        }
        {  // shift left by `filled` bytes (128-bit) and merge
          const uint32_t bsh = (filled & 3u) * 8u, wsh = filled >> 2;
          const uint32_t r0 = x.x << bsh, r1 = __funnelshift_l(x.x, x.y, bsh), r2 = __funnelshift_l(x.y, x.z, bsh), r3 = __funnelshift_l(x.z, x.w, bsh);
          if (wsh == 0) { v.x |= r0; v.y |= r1; v.z |= r2; v.w |= r3; }
          else if (wsh == 1) { v.y |= r0; v.z |= r1; v.w |= r2; }
          else if (wsh == 2) { v.z |= r0; v.w |= r1; }
          else v.w |= r0;
        }
        filled += take; within += take;
        if (filled >= 16u) break;
      }
      o4[c] = v;
    }
    if (lane == 0) {
      res.out_off = obase + P.out_bias; res.body_len = olen - po.path_len; res.path_len = (uint16_t)po.path_len; res.status = AIGW_OK; res.reason = 0;
      res.model_off = po.model_off; res.model_len = po.model_len; res.body_kind = (po.flags & 0x80u) ? AIGW_BODY_UNCHANGED : AIGW_BODY_BYTES; res.flags = po.flags & 0x7fu;
      P.results[doc] = res;
    }
    __syncwarp();
    li = li_next; cur = nxt;
  }
}

// ------------------------------------------------------------------ host-side launcher
template <int MAXD>
size_t work_bytes_cls(size_t ndocs) { return work_per_doc(layout_of<MAXD>()) * ndocs + kWorkHdr + 256; }

struct ClsConfig { bool ready = false; int bps1 = 1, bps3 = 1; };

// documents per sub-batch: the bodies of the sub-batches in flight (two streams) stay in the 126 MB L2 between index and emit
static size_t sub_batch_docs(int maxd) {
  static const long env = getenv("AIGW_CHAT_SUB") ? atol(getenv("AIGW_CHAT_SUB")) : 0;
  if (env > 0) return (size_t)env;
  size_t d = ((size_t)40 << 20) / (size_t)maxd;
  size_t p = 256; while (p * 2 <= d) p *= 2;
  return p;
}

template <int MAXD, int WARPS>
static cudaError_t launch_cls(const ChatParams& P, uint32_t first, int device, int sm_count, cudaStream_t st, uint8_t* work, size_t work_cap, const ChatAux* aux, cudaEvent_t* ev, int ev_cap,
                              int* launches) {
  using C = Cls<MAXD>;
  constexpr size_t hdr1 = (sizeof(IdTables) + 8 * WARPS + 15) & ~(size_t)15, hdr3 = (sizeof(LitTable::bytes) + 8 * WARPS + 15) & ~(size_t)15;
  const size_t smem1 = hdr1 + (size_t)(C::kIn + C::kTok * 6 + (((MAXD / 1024 + 1) * 4 + 15) & ~15)) * WARPS;
  const size_t smem3 = hdr3 + (size_t)EmitSmem<MAXD>::kWarpBytes * WARPS;
  static ClsConfig cfgs[kMaxDevices];
  static std::mutex mu;
  if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
  ClsConfig& cc = cfgs[device];
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!cc.ready) {
      cudaError_t e = cudaFuncSetAttribute(chat_index_kernel<MAXD, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(chat_emit_kernel<MAXD, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
      if (e != cudaSuccess) return e;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cc.bps1, chat_index_kernel<MAXD, WARPS>, WARPS * 32, smem1);
      if (e != cudaSuccess) return e;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cc.bps3, chat_emit_kernel<MAXD, WARPS>, WARPS * 32, smem3);
      if (e != cudaSuccess) return e;
      if (cc.bps1 < 1) cc.bps1 = 1;
      if (cc.bps3 < 1) cc.bps3 = 1;
      cc.ready = true;
    }
  }
  // stage timing (ev) needs the stages back to back on one stream; otherwise sub-batches rotate over the auxiliary streams
  static const int env_streams = getenv("AIGW_CHAT_STREAMS") ? atoi(getenv("AIGW_CHAT_STREAMS")) : 1;   // > 1: sub-batches rotate over auxiliary streams
  static const int env_idx_ctas = getenv("AIGW_IDX_CTAS") ? atoi(getenv("AIGW_IDX_CTAS")) : 0;      // experiments: cap resident CTAs per SM
  static const int env_emit_ctas = getenv("AIGW_EMIT_CTAS") ? atoi(getenv("AIGW_EMIT_CTAS")) : 0;
  static const int env_walk_ctas = getenv("AIGW_WALK_CTAS") ? atoi(getenv("AIGW_WALK_CTAS")) : 0;
  static const int env_pipe = getenv("AIGW_CHAT_PIPE") ? atoi(getenv("AIGW_CHAT_PIPE")) : 0;
  // ---- stage-pipelined mode: one stream per STAGE, sub-batches flow index -> walk -> emit through a ring of workspaces, so in
  // the steady state one index, one walk and one emit kernel (of three consecutive sub-batches) are resident on every SM at
  // once.  The walk is latency-bound (IPC 0.4) while index and emit are bound by the integer issue rate, so sharing an SM costs
  // each little; every kernel's grid is capped so that the three fit together (registers: 3x128x80 + 2x128x54 + 2x128x56).
  if (env_pipe && aux && !ev && P.n > 4096) {
    const int ic = env_idx_ctas > 0 ? env_idx_ctas : 2, wc = env_walk_ctas > 0 ? env_walk_ctas : 3, ec = env_emit_ctas > 0 ? env_emit_ctas : 2;
    const int bps1 = ic < cc.bps1 ? ic : cc.bps1, bps3 = ec < cc.bps3 ? ec : cc.bps3;
    size_t sub = sub_batch_docs(MAXD);
    int ring = kChatRing;
    const size_t per = (chat_work_bytes(MAXD, sub) + 511) & ~(size_t)255;
    while (ring > 1 && per * ring > work_cap) ring--;
    if (per * ring > work_cap) { sub = (work_cap - kWorkHdr - 256) / work_per_doc(layout_of<MAXD>()); ring = 1; }
    if (sub == 0) return cudaErrorMemoryAllocation;
    cudaStream_t sI = aux->s[0], sW = aux->s[1], sE = aux->s[2];
    cudaError_t e;
    if ((e = cudaEventRecord(aux->fork, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(sI, aux->fork, 0)) != cudaSuccess) return e;
    int nl = 0, sb = 0;
    for (uint32_t rel = 0; rel < P.n; rel += (uint32_t)sub, sb++) {
      const uint32_t doc0 = first + rel;
      const uint32_t nd = (uint32_t)(P.n - rel < sub ? P.n - rel : sub);
      const int slot = sb % ring;
      uint8_t* wk = work + (size_t)slot * per;
      const WorkPtrs wp = carve(wk, nd, layout_of<MAXD>());
      long long want = ((long long)nd + WARPS - 1) / WARPS;
      long long g1 = (long long)sm_count * bps1; if (want < g1) g1 = want;
      long long g3 = (long long)sm_count * bps3; if (want < g3) g3 = want;
      if (sb >= ring && (e = cudaStreamWaitEvent(sI, aux->emit_done[slot], 0)) != cudaSuccess) return e;   // the slot's previous tenant has been emitted
      if ((e = cudaMemsetAsync(wk, 0, kWorkHdr, sI)) != cudaSuccess) return e;
      chat_index_kernel<MAXD, WARPS><<<(unsigned)g1, WARPS * 32, smem1, sI>>>(P, doc0, nd, wk);
      if ((e = cudaEventRecord(aux->idx_done[slot], sI)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(sW, aux->idx_done[slot], 0)) != cudaSuccess) return e;
      { const unsigned gs = (nd + 8191) / 8192; chat_sort_kernel<<<gs < 32 ? gs : 32, kBins, 0, sW>>>(wp.ntok, wp.bins, wp.cursor, wp.perm, nd); }
      if ((e = launch_chat_walk(P, doc0, nd, wk, layout_of<MAXD>(), sW, wc)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(aux->walk_done[slot], sW)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(sE, aux->walk_done[slot], 0)) != cudaSuccess) return e;
      chat_emit_kernel<MAXD, WARPS><<<(unsigned)g3, WARPS * 32, smem3, sE>>>(P, doc0, nd, wk);
      if ((e = cudaEventRecord(aux->emit_done[slot], sE)) != cudaSuccess) return e;
      nl += 4;
    }
    // emits complete in order on sE, and every index / walk precedes its emit
    if ((e = cudaEventRecord(aux->join[0], sE)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(st, aux->join[0], 0)) != cudaSuccess) return e;
    if (launches) *launches = nl;
    return cudaGetLastError();
  }
  const bool overlap = aux && !ev && P.n > 1024 && env_streams > 1;
  const int nstreams = overlap ? (env_streams < kChatStreams ? env_streams : kChatStreams) : 1;
  const size_t half = (work_cap / nstreams) & ~(size_t)255;
  size_t sub = (half - kWorkHdr - 256) / work_per_doc(layout_of<MAXD>());
  const size_t want_sub = overlap ? sub_batch_docs(MAXD) : 262144;
  if (sub > want_sub) sub = want_sub;
  if (sub == 0) return cudaErrorMemoryAllocation;
  cudaError_t e;
  if (overlap) {
    if ((e = cudaEventRecord(aux->fork, st)) != cudaSuccess) return e;
    for (int s = 0; s < nstreams; s++) if ((e = cudaStreamWaitEvent(aux->s[s], aux->fork, 0)) != cudaSuccess) return e;
  }
  const int bps1 = overlap && env_idx_ctas > 0 && env_idx_ctas < cc.bps1 ? env_idx_ctas : cc.bps1;
  const int bps3 = overlap && env_emit_ctas > 0 && env_emit_ctas < cc.bps3 ? env_emit_ctas : cc.bps3;
  int nl = 0, sb = 0;
  for (uint32_t rel = 0; rel < P.n; rel += (uint32_t)sub, sb++) {
    const uint32_t doc0 = first + rel;
    const uint32_t nd = (uint32_t)(P.n - rel < sub ? P.n - rel : sub);
    cudaStream_t s = overlap ? aux->s[sb % nstreams] : st;
    uint8_t* wk = work + (size_t)(overlap ? (sb % nstreams) : 0) * half;
    const WorkPtrs wp = carve(wk, nd, layout_of<MAXD>());
    if ((e = cudaMemsetAsync(wk, 0, kWorkHdr, s)) != cudaSuccess) return e;
    long long want = ((long long)nd + WARPS - 1) / WARPS;
    long long g1 = (long long)sm_count * bps1; if (want < g1) g1 = want;
    long long g3 = (long long)sm_count * bps3; if (want < g3) g3 = want;
    const bool timed = ev && 4 * sb + 3 < ev_cap;
    if (timed) cudaEventRecord(ev[4 * sb + 0], s);
    chat_index_kernel<MAXD, WARPS><<<(unsigned)g1, WARPS * 32, smem1, s>>>(P, doc0, nd, wk);
    if (timed) cudaEventRecord(ev[4 * sb + 1], s);
    { const unsigned gs = (nd + 8191) / 8192; chat_sort_kernel<<<gs < 32 ? gs : 32, kBins, 0, s>>>(wp.ntok, wp.bins, wp.cursor, wp.perm, nd); }
    if ((e = launch_chat_walk(P, doc0, nd, wk, layout_of<MAXD>(), s, env_walk_ctas)) != cudaSuccess) return e;
    if (timed) cudaEventRecord(ev[4 * sb + 2], s);
    chat_emit_kernel<MAXD, WARPS><<<(unsigned)g3, WARPS * 32, smem3, s>>>(P, doc0, nd, wk);
    if (timed) cudaEventRecord(ev[4 * sb + 3], s);
    nl += 4;
  }
  if (overlap) {
    for (int s = 0; s < nstreams; s++) {
      if ((e = cudaEventRecord(aux->join[s], aux->s[s])) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(st, aux->join[s], 0)) != cudaSuccess) return e;
    }
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

size_t chat_work_bytes(uint32_t max_len, size_t ndocs) {
  if (max_len <= 2048) return work_bytes_cls<2048>(ndocs);
  if (max_len <= 5120) return work_bytes_cls<5120>(ndocs);
  if (max_len <= 9216) return work_bytes_cls<9216>(ndocs);
  if (max_len <= 17408) return work_bytes_cls<17408>(ndocs);
  if (max_len <= 33792) return work_bytes_cls<33792>(ndocs);
  return work_bytes_cls<65536>(ndocs);
}
// workspace that lets a call of n documents run at full overlap: two sub-batches in flight
size_t chat_work_bytes_for(uint32_t max_len, size_t n) {
  static const int env_streams = getenv("AIGW_CHAT_STREAMS") ? atoi(getenv("AIGW_CHAT_STREAMS")) : 1;
  static const int env_pipe = getenv("AIGW_CHAT_PIPE") ? atoi(getenv("AIGW_CHAT_PIPE")) : 0;
  if (env_pipe && n > 4096) {
    const int maxd = max_len <= 2048 ? 2048 : max_len <= 5120 ? 5120 : max_len <= 9216 ? 9216 : max_len <= 17408 ? 17408 : max_len <= 33792 ? 33792 : 65536;
    return kChatRing * ((chat_work_bytes(maxd, sub_batch_docs(maxd)) + 511) & ~(size_t)255) + 512;
  }
  if (n <= 1024 || env_streams <= 1) return chat_work_bytes(max_len, n < 262144 ? n : 262144) + 512;
  const int maxd = max_len <= 2048 ? 2048 : max_len <= 5120 ? 5120 : max_len <= 9216 ? 9216 : max_len <= 17408 ? 17408 : max_len <= 33792 ? 33792 : 65536;
  size_t sub = sub_batch_docs(maxd);
  if (sub > n) sub = n;
  return kChatStreams * (chat_work_bytes(max_len, sub) + 512);
}

cudaError_t launch_chat_translate_range(const ChatParams& P, uint32_t first, uint32_t count, uint32_t max_len, int device, int sm_count, cudaStream_t st, uint8_t* work, size_t work_cap,
                                        const ChatAux* aux, int* launches) {
  ChatParams Q = P; Q.n = count;
  return launch_chat_translate(Q, max_len, device, sm_count, st, work, work_cap, aux, launches, nullptr, 0, first);
}

cudaError_t launch_chat_translate(const ChatParams& P, uint32_t max_len, int device, int sm_count, cudaStream_t st, uint8_t* work, size_t work_cap, const ChatAux* aux, int* launches,
                                  cudaEvent_t* ev, int ev_cap, uint32_t first) {
  if (max_len <= 2048) return launch_cls<2048, 4>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
  if (max_len <= 5120) return launch_cls<5120, 4>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
  if (max_len <= 9216) return launch_cls<9216, 2>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
  if (max_len <= 17408) return launch_cls<17408, 2>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
  if (max_len <= 33792) return launch_cls<33792, 1>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
  return launch_cls<65536, 1>(P, first, device, sm_count, st, work, work_cap, aux, ev, ev_cap, launches);
}

}  // namespace aigw
