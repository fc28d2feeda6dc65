// Entry points shared between translation units of the library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/aigw_b200.h"

namespace aigw {
int ctx_device(const aigw_ctx* ctx);
// Fused small-batch translate (chat_walk_impl.cuh, one CTA per body of at most kSmallMaxLen bytes), asynchronous on the
// caller's stream.  Every pointer must be device-accessible (mapped pinned or device memory): bodies at in + offsets[i],
// record of body i stored at out + out_slot[i] (slot capacity out_slot[i+1] - out_slot[i]), results[i].out_off relative to out.
// Uses no context state, so it may run while another thread is inside a host call of the same context.
cudaError_t chat_small_launch(const aigw_backend_cfg* cfg, const uint8_t* in, const uint64_t* offsets, const uint32_t* lens, const uint64_t* out_slot, uint32_t n, uint8_t* out,
                              aigw_doc_result* results, cudaStream_t st);
// true when a body of this length goes through the fused kernel under cfg (response schemas plan in a roomier class)
bool chat_small_fits(const aigw_backend_cfg* cfg, uint32_t len);
}  // namespace aigw
