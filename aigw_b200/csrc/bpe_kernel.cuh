// K4: byte-level BPE token count (bpe_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_once.cuh"

namespace aigw {

struct BpeParams {
  const uint8_t* text;        // texts back to back
  const uint64_t* offsets;    // n
  const uint32_t* lens;       // n
  uint32_t n;
  uint32_t* counts;           // n: tokens per text, 0xFFFFFFFF = declined (a space-free run longer than the staging buffer)
  unsigned int* next;         // work counter (zeroed by the caller)
  const uint2* table;         // open-addressing hash of the merges: {a << 16 | b, rank << 16 | merged}, empty = {0xFFFFFFFF, …}
  uint32_t slots;             // power of two
  const uint16_t* byte_to_id; // 256
};

size_t bpe_smem_bytes(uint32_t slots);
cudaError_t launch_bpe_count(const BpeParams& P, int sm_count, cudaStream_t st);

}  // namespace aigw
