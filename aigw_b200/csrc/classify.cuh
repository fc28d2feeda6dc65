// Bit-sliced byte classification for the structural index: a lane's 32 bytes (8 words) are transposed into 8 bit planes
// (plane b, bit i = bit b of byte i), after which every character class is a handful of LOP3s on 32-bit masks that are
// already in byte order — no per-class compare + mask extraction per word, and no sparse loops over bytes outside strings.
//   step 1  two 4x4 byte transposes (16 PRMT): word J, byte K := memory byte 8K + J
//   step 2  8x8 bit transpose across the 8 words, 4 byte lanes at once (12 delta swaps)
//   step 3  class masks as boolean functions of the planes
// __host__ __device__ so that tests/test_classify_cpu exercises the same code on the CPU against a byte loop.
#pragma once
#include <stdint.h>

namespace aigw {

struct ByteClasses {
  uint32_t quote;   // '"'
  uint32_t bslash;  // '\\'
  uint32_t ctl;     // < 0x20
  uint32_t op;      // { } [ ] : ,
  uint32_t ws;      // space \t \n \r
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
static inline uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) r |= (uint32_t)((src >> (8 * ((sel >> (4 * k)) & 7))) & 0xff) << (8 * k);
  return r;
}
#endif

#if defined(__CUDACC__)
#define AIGW_HD __host__ __device__ __forceinline__
#else
#define AIGW_HD inline
#endif

AIGW_HD void transpose4x4_bytes(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t* out) {
  const uint32_t t0 = bperm(a0, a1, 0x5140), t1 = bperm(a0, a1, 0x7362), t2 = bperm(a2, a3, 0x5140), t3 = bperm(a2, a3, 0x7362);
  out[0] = bperm(t0, t2, 0x5410); out[1] = bperm(t0, t2, 0x7632); out[2] = bperm(t1, t3, 0x5410); out[3] = bperm(t1, t3, 0x7632);
}

AIGW_HD void delta_swap(uint32_t& a, uint32_t& b, int d, uint32_t m) {
  const uint32_t t = ((a >> d) ^ b) & m;
  b ^= t;
  a ^= t << d;
}

// planes[b] bit i = bit b of memory byte i (i = 0..31), w = the 32 bytes as little-endian words
AIGW_HD void bit_planes32(const uint32_t* w, uint32_t* p) {
  transpose4x4_bytes(w[0], w[2], w[4], w[6], p);
  transpose4x4_bytes(w[1], w[3], w[5], w[7], p + 4);
  delta_swap(p[0], p[1], 1, 0x55555555u); delta_swap(p[2], p[3], 1, 0x55555555u); delta_swap(p[4], p[5], 1, 0x55555555u); delta_swap(p[6], p[7], 1, 0x55555555u);
  delta_swap(p[0], p[2], 2, 0x33333333u); delta_swap(p[1], p[3], 2, 0x33333333u); delta_swap(p[4], p[6], 2, 0x33333333u); delta_swap(p[5], p[7], 2, 0x33333333u);
  delta_swap(p[0], p[4], 4, 0x0f0f0f0fu); delta_swap(p[1], p[5], 4, 0x0f0f0f0fu); delta_swap(p[2], p[6], 4, 0x0f0f0f0fu); delta_swap(p[3], p[7], 4, 0x0f0f0f0fu);
}

AIGW_HD ByteClasses classify32(const uint32_t* w) {
  uint32_t p[8];
  bit_planes32(w, p);
  const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3], b4 = p[4], b5 = p[5], b6 = p[6], b7 = p[7];
  ByteClasses c;
  const uint32_t hi00 = ~(b7 | b6);                 // 00xx xxxx
  const uint32_t hi01 = ~b7 & b6;                   // 01xx xxxx
  c.ctl = hi00 & ~b5;                               // 000x xxxx
  const uint32_t n432 = ~(b4 | b3 | b2);            // xxx0 00xx
  const uint32_t x001 = hi00 & b5 & ~b4;            // 0010 xxxx
  c.quote = x001 & ~b3 & ~b2 & b1 & ~b0;            // 0010 0010
  c.bslash = hi01 & ~b5 & b4 & b3 & b2 & ~b1 & ~b0; // 0101 1100
  const uint32_t brace = hi01 & b4 & b3 & b0 & (b2 ^ b1);  // 01x1 1011 / 01x1 1101: { } [ ]
  const uint32_t colon = hi00 & b5 & b4 & b3 & ~b2 & b1 & ~b0;  // 0011 1010
  const uint32_t comma = x001 & b3 & b2 & ~b1 & ~b0;            // 0010 1100
  c.op = brace | colon | comma;
  const uint32_t space = x001 & ~b3 & ~b2 & ~b1 & ~b0;          // 0010 0000
  // 0000 1001 (\t) 0000 1010 (\n) 0000 1101 (\r): 0000 1xxx with low bits 001, 010, 101
  const uint32_t low = (~b2 & (b1 ^ b0)) | (b2 & ~b1 & b0);
  c.ws = space | (c.ctl & ~b4 & b3 & low);
  (void)n432;
  return c;
}

}  // namespace aigw
