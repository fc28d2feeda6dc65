// Large /v1/embeddings requests: ParseBody over the string forms of the input union + the text table of the BPE count (emb_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aigw_b200.h"

namespace aigw {

struct EmbScanParams {
  const uint8_t* bodies;      // requests back to back, each start 16-byte aligned
  const uint64_t* offsets;    // n
  const uint32_t* lens;       // n
  uint32_t n;
  aigw_emb_count_result* results;
  uint32_t* tok_ws;           // emb_scan_grid() slices of tok_cap token words (position | type << 24)
  uint32_t tok_cap;
  uint64_t* text_off;         // text table: absolute offset / length of every input string, rows allocated per request
  uint32_t* text_len;
  uint32_t text_cap;
  unsigned int* next;         // work counter (zeroed by the caller)
  unsigned int* text_used;    // rows handed out (zeroed by the caller)
};

int emb_scan_grid(int sm_count);
cudaError_t launch_emb_scan(const EmbScanParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_emb_sum(aigw_emb_count_result* results, uint32_t n, const uint32_t* counts, cudaStream_t st);

}  // namespace aigw
