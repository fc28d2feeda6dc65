// See sha256_kernel.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aigw_b200.h"

namespace aigw {
cudaError_t launch_sha256(const uint8_t* bytes, const uint64_t* off, const uint32_t* len, const aigw_doc_result* results, uint32_t n, uint8_t* digests, cudaStream_t st);
}
