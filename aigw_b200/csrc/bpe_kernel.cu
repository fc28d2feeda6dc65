// K4: byte-level BPE token count on the GPU (north_star: "a GPU BPE tokenizer with the merge table staged in shared memory for
// token-based rate limiting").  The reference has no tokenizer, so the semantics are the ones include/aigw_b200.h
// states and tests/golden/make_bpe_vocab.py fixes with HuggingFace `tokenizers`:
//   pre-tokenisation  every space starts a new piece; a piece is its first byte plus the bytes up to the next space
//   model             inside a piece, merge the adjacent pair of lowest rank (leftmost first among equal ranks) until none is mergeable
//   result            tokens per text = sum of the pieces' final lengths
//
// Mapping.  Persistent CTAs; the merge table (open-addressing hash, key = a << 16 | b, value = rank << 16 | merged) is copied into
// shared memory once per CTA.  A warp pulls a GROUP of consecutive texts whose bytes fit its staging buffer (short inputs — 64
// characters in BASELINE config 3 — would leave most lanes idle one text at a time), stages the bytes, turns
// them into symbol ids, finds the piece starts with ballots, and then every LANE runs the merge loop of one piece at a time
// (pieces are independent).  Per piece the ranks of the adjacent pairs are kept next to the symbols, so a merge costs two
// table lookups and one shift instead of a rescan.  A text longer than the buffer is cut at piece boundaries; a single piece
// longer than the buffer is DECLINED (count 0xFFFFFFFF) rather than approximated.  sm_100a only.
#include "bpe_kernel.cuh"

#include <cstdio>

namespace aigw {

static constexpr int kBpeWarps = 32;            // the merge loops are latency-bound (dependent shared-memory lookups): as many warps as shared memory allows
static constexpr int kBpeBuf = 512;            // staged bytes per warp (a piece longer than this is declined); small, so that 32 warps fit beside the table
static constexpr int kBpeMaxTexts = 32;         // texts per group
static constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ uint32_t bpe_hash(uint32_t key) { key *= 0x9E3779B1u; return key ^ (key >> 15); }

// rank << 16 | merged of the pair (a, b), or kEmpty
__device__ __forceinline__ uint32_t bpe_lookup(const uint2* tab, uint32_t mask, uint32_t a, uint32_t b) {
  const uint32_t key = (a << 16) | b;
  uint32_t s = bpe_hash(key) & mask;
  for (;;) {
    const uint2 e = tab[s];
    if (e.x == key) return e.y;
    if (e.x == kEmpty) return kEmpty;
    s = (s + 1) & mask;
  }
}

struct BpeWarpSmem {
  uint8_t bytes[kBpeBuf + 16];
  uint16_t sym[kBpeBuf];
  uint16_t rk[kBpeBuf];          // rank of the pair (sym[i], sym[i+1]); 0xffff = not mergeable
  uint16_t start[kBpeBuf + 1];   // piece starts, ascending; start[np] = end of data
  uint32_t tcount[kBpeMaxTexts];
  uint32_t tend[kBpeMaxTexts];   // staged end offset of each text of the group
};

__global__ void __launch_bounds__(kBpeWarps * 32, 1) bpe_count_kernel(const BpeParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* tab = (uint2*)smem;
  const uint32_t mask = P.slots - 1u;
  for (uint32_t i = threadIdx.x; i < P.slots; i += blockDim.x) tab[i] = P.table[i];
  __shared__ uint16_t s_b2i[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_b2i[i] = P.byte_to_id[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  BpeWarpSmem& W = *(BpeWarpSmem*)(smem + (size_t)P.slots * 8 + (size_t)warp * sizeof(BpeWarpSmem));

  for (;;) {
    // ---- a group of consecutive texts: the counter hands out kBpeMaxTexts at a time, the warp walks them in buffer-sized runs
    uint32_t g0 = 0;
    if (lane == 0) g0 = atomicAdd(P.next, (unsigned)kBpeMaxTexts);
    g0 = __shfl_sync(0xffffffffu, g0, 0);
    if (g0 >= P.n) break;
    const uint32_t g1 = min(P.n, g0 + (uint32_t)kBpeMaxTexts);
    uint32_t t = g0;        // next text to stage
    uint32_t t_off = 0;     // bytes of text t already counted (a text longer than the buffer is cut at piece boundaries)
    uint32_t t_acc = 0;     // tokens of that counted part
    while (t < g1) {
      // ---- stage texts t, t+1, … while they fit
      const uint32_t first = t; const bool first_cont = t_off != 0;
      uint32_t fill = 0, nt = 0; bool last_partial = false;
      while (t < g1) {
        const uint32_t len = P.lens[t] - t_off;
        const uint8_t* src = P.text + P.offsets[t] + t_off;
        uint32_t take = len;
        if (fill + len > (uint32_t)kBpeBuf) {
          if (nt > 0) break;                                   // what is staged so far is counted first
          uint32_t cut = 0;                                    // the last piece start inside the window
          for (uint32_t b0 = 0; b0 < (uint32_t)kBpeBuf; b0 += 32) { const uint32_t i = b0 + lane; const uint32_t m = __ballot_sync(0xffffffffu, i > 0 && src[i] == ' '); if (m) cut = b0 + 31u - __clz(m); }
          if (cut == 0) { if (lane == 0) P.counts[t] = 0xffffffffu; t++; t_off = 0; t_acc = 0; break; }   // one piece longer than the buffer: declined
          take = cut; last_partial = true;
        }
        for (uint32_t i = lane; i < take; i += 32) W.bytes[fill + i] = src[i];
        fill += take;
        if (lane == 0) { W.tend[nt] = fill; W.tcount[nt] = 0; }
        nt++;
        if (last_partial) { t_off += take; break; }
        t++; t_off = 0;
      }
      if (nt == 0) continue;
      __syncwarp();
      // ---- symbols and piece starts: a piece starts at every space and at the first byte of every staged text
      for (uint32_t i = lane; i < fill; i += 32) W.rk[i] = 0;
      __syncwarp();
      for (uint32_t k = lane; k < nt; k += 32) { const uint32_t s0 = k == 0 ? 0u : W.tend[k - 1]; if (s0 < fill) W.rk[s0] = 1; }
      __syncwarp();
      uint32_t np = 0;
      for (uint32_t b0 = 0; b0 < fill; b0 += 32) {
        const uint32_t i = b0 + lane;
        const uint32_t c = i < fill ? W.bytes[i] : 0u;
        if (i < fill) W.sym[i] = s_b2i[c];
        const bool st = i < fill && (c == ' ' || W.rk[i] != 0);
        const uint32_t m = __ballot_sync(0xffffffffu, st);
        if (st) W.start[np + __popc(m & lt)] = (uint16_t)i;
        np += __popc(m);
      }
      if (lane == 0) W.start[np] = (uint16_t)fill;
      __syncwarp();
      // ---- one piece per lane at a time
      for (uint32_t p0 = 0; p0 < np; p0 += 32) {
        const uint32_t pi = p0 + lane;
        if (pi < np) {
          const uint32_t a = W.start[pi], e = W.start[pi + 1];
          uint32_t L = e - a;
          uint16_t* s = W.sym + a; uint16_t* r = W.rk + a;
          for (uint32_t k = 0; k + 1 < L; k++) { const uint32_t v = bpe_lookup(tab, mask, s[k], s[k + 1]); r[k] = v == kEmpty ? 0xffffu : (uint16_t)(v >> 16); }
          while (L > 1) {
            uint32_t best = 0xffffu, at = 0;
            for (uint32_t k = 0; k + 1 < L; k++) { const uint32_t v = r[k]; if (v < best) { best = v; at = k; } }
            if (best == 0xffffu) break;
            s[at] = (uint16_t)(bpe_lookup(tab, mask, s[at], s[at + 1]) & 0xffffu);
            for (uint32_t k = at + 1; k + 1 < L; k++) { s[k] = s[k + 1]; r[k] = r[k + 1]; }
            L--;
            if (at > 0) { const uint32_t v = bpe_lookup(tab, mask, s[at - 1], s[at]); r[at - 1] = v == kEmpty ? 0xffffu : (uint16_t)(v >> 16); }
            if (at + 1 < L) { const uint32_t v = bpe_lookup(tab, mask, s[at], s[at + 1]); r[at] = v == kEmpty ? 0xffffu : (uint16_t)(v >> 16); }
          }
          uint32_t lo = 0; while (W.tend[lo] <= a) lo++;      // the text this piece belongs to (empty texts own no piece)
          atomicAdd(&W.tcount[lo], L);
        }
      }
      __syncwarp();
      // ---- results of the texts that ended in this run; a cut text carries its partial count into the next run
      uint32_t carry = 0;
      for (uint32_t k0 = 0; k0 < nt; k0 += 32) {
        const uint32_t k = k0 + lane;
        if (k < nt) {
          const uint32_t c = W.tcount[k] + ((k == 0 && first_cont) ? t_acc : 0u);
          if (k == nt - 1 && last_partial) carry = c; else P.counts[first + k] = c;
        }
      }
      carry = __reduce_max_sync(0xffffffffu, carry);
      t_acc = last_partial ? carry : 0u;
      __syncwarp();
    }
  }
}

size_t bpe_smem_bytes(uint32_t slots) { return (size_t)slots * 8 + (size_t)kBpeWarps * sizeof(BpeWarpSmem); }

cudaError_t launch_bpe_count(const BpeParams& P, int sm_count, cudaStream_t st) {
  const size_t smem = bpe_smem_bytes(P.slots);
  static DeviceOnce once;
  int* vals = nullptr;
  cudaError_t e = device_once(once, &vals, [&](int*) { return cudaFuncSetAttribute(bpe_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
  if (e != cudaSuccess) return e;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  const unsigned groups = (P.n + kBpeMaxTexts - 1) / kBpeMaxTexts;
  const unsigned want = (groups + kBpeWarps - 1) / kBpeWarps;
  const unsigned grid = want < (unsigned)sm_count ? (want ? want : 1u) : (unsigned)sm_count;
  bpe_count_kernel<<<grid, kBpeWarps * 32, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace aigw
