// Per-device one-time setup (schema tables in __device__ memory, dynamic shared-memory attributes, occupancy): a process may
// hold one context per GPU, so "configured" is a per-device fact, decided under a lock (process-wide statics left
// the second device with zeroed tables and un-raised shared-memory limits).
#pragma once
#include <cuda_runtime.h>

#include <mutex>

namespace aigw {

static constexpr int kOnceDevices = 16;
struct DeviceOnce {
  std::mutex mu;
  bool ready[kOnceDevices] = {};
  int val[kOnceDevices][2] = {};   // per-device results of the setup (e.g. resident blocks per SM)
};

// runs f(int* val) once per device; returns its error, or cudaSuccess and the per-device values in *out
template <class F>
inline cudaError_t device_once(DeviceOnce& o, int** out, F&& f) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kOnceDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(o.mu);
  if (!o.ready[dev]) {
    e = f(o.val[dev]);
    if (e != cudaSuccess) return e;
    o.ready[dev] = true;
  }
  if (out) *out = o.val[dev];
  return cudaSuccess;
}

}  // namespace aigw
