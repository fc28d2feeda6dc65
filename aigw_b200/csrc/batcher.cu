// aigw_b200 — per-GPU request batcher: the synchronous single-request call a cgo shim makes from thousands of goroutines
// ("a synchronous single-doc wrapper for the Go shim sits on top of the batch queue"; the reference calls
// ParseBody / RequestBody once per request on the request's own goroutine, internal/extproc/processor_impl.go:211-398).
//
// Data path (no lock on it):
//   * every backend ("lane") owns a ring of batch buffers in mapped pinned memory: fixed-stride input slots, fixed-stride
//     output slots, a result table.  The fused small-batch kernel (chat_walk_impl.cuh) reads and writes them in place;
//   * a submitter claims a slot of the lane's OPEN batch with one compare-and-swap on the batch word (generation | state |
//     count), copies its body into the slot ITSELF (the copies of a batch run on the callers' cores, in parallel), bumps the
//     batch's `filled` counter and sleeps on the batch's futex word;
//   * one dispatcher thread closes a batch when it is full or `window_us` after it was opened, waits for `filled`, launches
//     the kernel on the batch's own stream and goes on to the next batch: several batches are in flight at once, and the
//     next one fills while the previous ones run;
//   * the same thread polls the in-flight batches' events; a finished batch is published with one store + one FUTEX_WAKE
//     for all of its waiters; every waiter copies its own record out and the last one returns the buffer to the ring.
// Bodies the fused kernel does not take (above its 5120-byte class) go through aigw_chat_translate_host one at a time under a
// mutex.  A failed launch never surfaces as an error: its requests come back DECLINED (the stock path handles them), as
// SURVEY §8b asks ("a GPU fault must degrade to the CPU path, never to a stream error").  There is no CPU path here either.
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <climits>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "chat_kernel.cuh"
#include "lib_internal.h"

namespace {

inline long futex(std::atomic<uint32_t>* addr, int op, uint32_t val) { return syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), op | FUTEX_PRIVATE_FLAG, val, nullptr, nullptr, 0); }
inline void cpu_relax() { __builtin_ia32_pause(); }
inline int64_t now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

constexpr uint32_t kInStride = ((aigw::kSmallMaxLen + 15u) & ~15u) + 16;                      // one input slot
constexpr uint32_t kOutStride = ((aigw::kSmallMaxLen * 3 + 1024 + 15) & ~15u);   // one output slot (a record that does not fit is DECLINED)
constexpr int kRing = 4;          // batch buffers per lane
constexpr int kMaxLanes = 16;
enum : uint64_t { ST_FREE = 0, ST_OPEN = 1, ST_CLOSED = 2 };
// batch word: generation (high 24 bits) | state (8 bits) | count (32 bits)
inline uint64_t mk(uint64_t gen, uint64_t st, uint64_t cnt) { return (gen << 40) | (st << 32) | cnt; }
inline uint64_t w_gen(uint64_t w) { return w >> 40; }
inline uint64_t w_state(uint64_t w) { return (w >> 32) & 0xff; }
inline uint32_t w_count(uint64_t w) { return (uint32_t)w; }

struct Batch {
  uint8_t* in = nullptr; uint32_t* lens = nullptr; uint8_t* out = nullptr; aigw_doc_result* res = nullptr;   // mapped pinned
  std::atomic<uint64_t> word{mk(0, ST_FREE, 0)};
  std::atomic<uint32_t> filled{0}, done{0}, readers{0};
  std::atomic<int64_t> t_open{0};
  std::atomic<uint32_t> launched{0};   // set by the dispatcher once per generation
  int rc = 0; uint32_t n = 0;
  bool in_flight = false;
  cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
};

struct Lane {
  aigw_backend_cfg cfg{};
  std::string s_override, s_prefix, s_version, s_rid;
  Batch ring[kRing];
  std::atomic<int> open{-1};   // index of the OPEN batch, -1 when none
  uint64_t* offs = nullptr; uint64_t* slots = nullptr;   // static tables (mapped pinned): i * stride
  uint8_t* pinned = nullptr;
};

}  // namespace

struct aigw_batcher {
  aigw_ctx* ctx = nullptr;
  int device = 0;
  uint32_t max_batch = 256, window_us = 50;
  Lane* lanes[kMaxLanes] = {};
  std::atomic<int> n_lanes{0};
  std::mutex lane_mu;              // lane creation
  std::mutex big_mu;               // bodies outside the fused kernel's class: one context host call at a time
  std::atomic<bool> stop{false};
  std::atomic<uint32_t> kick{0};   // futex: the dispatcher sleeps here when nothing is open or in flight
  std::atomic<uint32_t> freed{0};  // futex: submitters that found no FREE buffer sleep here
  std::atomic<uint32_t> active{0}; // batches that are OPEN, CLOSED or in flight
  std::thread th;
  std::atomic<uint64_t> n_batches{0}, n_docs{0}, max_seen{0};

  int add_lane(const aigw_backend_cfg* cfg) {
    std::lock_guard<std::mutex> g(lane_mu);
    const int idx = n_lanes.load();
    if (idx >= kMaxLanes) return -6;
    cudaSetDevice(device);
    Lane* L = new Lane();
    L->cfg = *cfg;
    auto keep = [](std::string& dst, const char*& p) { if (p) { dst = p; p = dst.c_str(); } };   // the configuration strings must outlive the caller's
    keep(L->s_override, L->cfg.model_name_override); keep(L->s_prefix, L->cfg.openai_prefix); keep(L->s_version, L->cfg.api_version); keep(L->s_rid, L->cfg.response_id);
    const size_t tab = (((size_t)max_batch + 1) * 16 + 63) & ~(size_t)63;
    auto r64 = [](size_t x) { return (x + 63) & ~(size_t)63; };
    const size_t sz_in = r64((size_t)max_batch * kInStride + 64), sz_len = r64((size_t)max_batch * 4), sz_out = r64((size_t)max_batch * kOutStride + 64), sz_res = r64((size_t)max_batch * sizeof(aigw_doc_result));
    const size_t per = sz_in + sz_len + sz_out + sz_res;
    if (cudaHostAlloc((void**)&L->pinned, tab + per * kRing, cudaHostAllocMapped) != cudaSuccess) { delete L; return -3; }
    L->offs = (uint64_t*)L->pinned; L->slots = L->offs + max_batch + 1;
    for (uint32_t i = 0; i <= max_batch; i++) { L->offs[i] = (uint64_t)i * kInStride; L->slots[i] = (uint64_t)i * kOutStride; }
    uint8_t* p = L->pinned + tab;
    for (int k = 0; k < kRing; k++) {
      Batch& B = L->ring[k];
      B.in = p; p += sz_in;
      B.lens = (uint32_t*)p; p += sz_len;
      B.out = p; p += sz_out;
      B.res = (aigw_doc_result*)p; p += sz_res;
      if (cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&B.ev, cudaEventDisableTiming) != cudaSuccess) { return -3; }
    }
    lanes[idx] = L;
    n_lanes.store(idx + 1, std::memory_order_release);
    return idx;
  }

  // ---- dispatcher: close, launch, publish
  void publish(Batch& B, int rc) {
    B.rc = rc; B.in_flight = false;
    B.readers.store(B.n, std::memory_order_relaxed);
    B.done.store(1, std::memory_order_release);
    futex(&B.done, FUTEX_WAKE, INT_MAX);
  }
  void launch(Lane& L, Batch& B, uint32_t n) {
    B.n = n; B.launched.store(1, std::memory_order_relaxed);
    cudaError_t e = aigw::chat_small_launch(&L.cfg, B.in, L.offs, B.lens, L.slots, n, B.out, B.res, B.st);
    if (e == cudaSuccess) e = cudaEventRecord(B.ev, B.st);
    n_batches++; n_docs += n; if (n > max_seen.load()) max_seen = n;
    if (e != cudaSuccess) { cudaGetLastError(); publish(B, (int)e); return; }
    B.in_flight = true;
  }
  // no claim succeeds after the compare-and-swap; whoever wins it detaches the batch from the lane
  bool close_batch(Lane& L, int bi, uint64_t w) {
    Batch& B = L.ring[bi];
    for (;;) {
      if (w_state(w) != ST_OPEN) return false;
      if (B.word.compare_exchange_weak(w, mk(w_gen(w), ST_CLOSED, w_count(w)), std::memory_order_acq_rel)) break;
    }
    int e = bi; L.open.compare_exchange_strong(e, -1, std::memory_order_acq_rel);
    return true;
  }
  void run() {
    cudaSetDevice(device);
    for (;;) {
      if (active.load(std::memory_order_acquire) == 0) {
        if (stop.load()) return;
        const uint32_t k = kick.load(std::memory_order_acquire);
        if (active.load(std::memory_order_acquire) == 0 && !stop.load()) futex(&kick, FUTEX_WAIT, k);
        continue;
      }
      const int nl = n_lanes.load(std::memory_order_acquire);
      const int64_t t = now_ns();
      bool idle = true;
      for (int li = 0; li < nl; li++) {
        Lane& L = *lanes[li];
        const int bi = L.open.load(std::memory_order_acquire);
        if (bi >= 0) {
          Batch& B = L.ring[bi];
          uint64_t w = B.word.load(std::memory_order_acquire);
          if (w_state(w) == ST_OPEN && w_count(w) > 0 && (t - B.t_open.load(std::memory_order_relaxed) >= (int64_t)window_us * 1000 || stop.load())) close_batch(L, bi, w);
        }
        // closed batches (by the window above, or by the submitter that found them full) are launched as soon as every claimant
        // has finished copying its body in; the dispatcher never waits for one
        for (int k = 0; k < kRing; k++) {
          Batch& B = L.ring[k];
          const uint64_t w = B.word.load(std::memory_order_acquire);
          if (w_state(w) != ST_CLOSED || w_count(w) == 0 || B.launched.load(std::memory_order_relaxed)) continue;
          if (B.filled.load(std::memory_order_acquire) != w_count(w)) continue;
          launch(L, B, w_count(w));
          idle = false;
        }
        for (int k = 0; k < kRing; k++) {
          Batch& B = L.ring[k];
          if (!B.in_flight) continue;
          const cudaError_t q = cudaEventQuery(B.ev);
          if (q == cudaErrorNotReady) continue;
          if (q != cudaSuccess) cudaGetLastError();
          publish(B, (int)q);
          idle = false;
        }
      }
      if (idle) cpu_relax();
    }
  }

  // ---- submitter
  int open_batch(Lane& L) {   // returns the lane's open batch index, opening one if none is
    for (;;) {
      int bi = L.open.load(std::memory_order_acquire);
      if (bi >= 0) return bi;
      for (int k = 0; k < kRing; k++) {
        Batch& B = L.ring[k];
        uint64_t w = B.word.load(std::memory_order_acquire);
        if (w_state(w) != ST_FREE) continue;
        if (!B.word.compare_exchange_strong(w, mk(w_gen(w) + 1, ST_CLOSED, 0), std::memory_order_acq_rel)) continue;   // reserved (not claimable yet)
        B.filled.store(0, std::memory_order_relaxed); B.done.store(0, std::memory_order_relaxed); B.launched.store(0, std::memory_order_relaxed); B.t_open.store(now_ns(), std::memory_order_relaxed);
        int expect = -1;
        if (L.open.compare_exchange_strong(expect, k, std::memory_order_acq_rel)) {
          active.fetch_add(1, std::memory_order_acq_rel);
          B.word.store(mk(w_gen(w) + 1, ST_OPEN, 0), std::memory_order_release);
          kick.fetch_add(1, std::memory_order_release); futex(&kick, FUTEX_WAKE, 1);
          return k;
        }
        B.word.store(mk(w_gen(w) + 1, ST_FREE, 0), std::memory_order_release);   // somebody else opened one first
        return expect;
      }
      if (stop.load()) return -1;
      // every buffer is busy: wait for a release
      const uint32_t f = freed.load(std::memory_order_acquire);
      if (L.open.load(std::memory_order_acquire) >= 0) continue;
      bool any_free = false;
      for (int k = 0; k < kRing; k++) any_free |= w_state(L.ring[k].word.load(std::memory_order_acquire)) == ST_FREE;
      if (!any_free) futex(&freed, FUTEX_WAIT, f);
    }
  }

  int translate_big(Lane& L, const uint8_t* body, uint32_t len, uint8_t* out, uint32_t out_cap, aigw_doc_result* res) {
    std::lock_guard<std::mutex> g(big_mu);
    const uint64_t off = 0; aigw_batch_out bo;
    std::vector<uint8_t> padded((size_t)len + 32, ' ');
    memcpy(padded.data(), body, len);
    aigw_doc_result r; memset(&r, 0, sizeof r); r.status = AIGW_DECLINED; r.in_len = len;
    const int rc = aigw_chat_translate_host(ctx, &L.cfg, padded.data(), &off, &len, 1, &bo);
    n_batches++; n_docs++;
    if (!rc) {
      r = bo.results[0];
      if (r.status == AIGW_OK) {
        const uint32_t bytes = (uint32_t)r.path_len + r.body_len;
        if (bytes > out_cap) { r.status = AIGW_DECLINED; r.reason = AIGW_R_OUT_SPACE; r.out_off = 0; if (res) *res = r; return -4; }
        memcpy(out, bo.out + r.out_off, bytes); r.out_off = 0;
      }
    } else { r.reason = AIGW_R_ARENA_FULL; }
    if (res) *res = r;
    return 0;
  }

  int translate(int lane, const uint8_t* body, uint32_t len, uint8_t* out, uint32_t out_cap, aigw_doc_result* res) {
    aigw_doc_result r; memset(&r, 0, sizeof r); r.status = AIGW_DECLINED; r.in_len = len;
    if (res) *res = r;
    if (lane < 0 || lane >= n_lanes.load(std::memory_order_acquire)) return -2;
    if (stop.load()) return -5;
    Lane& L = *lanes[lane];
    if (len == 0 || !aigw::chat_small_fits(&L.cfg, len)) return translate_big(L, body, len, out, out_cap, res);
    // claim a slot of the open batch
    Batch* Bp = nullptr; uint32_t slot = 0;
    for (;;) {
      const int bi = open_batch(L);
      if (bi < 0) return -5;
      Batch& B = L.ring[bi];
      uint64_t w = B.word.load(std::memory_order_acquire);
      if (w_state(w) != ST_OPEN) { cpu_relax(); continue; }
      if (w_count(w) >= max_batch) { close_batch(L, bi, w); continue; }   // full: close it here, the next iteration opens the next one
      if (!B.word.compare_exchange_weak(w, w + 1, std::memory_order_acq_rel)) continue;
      Bp = &B; slot = w_count(w);
      break;
    }
    Batch& B = *Bp;
    memcpy(B.in + (size_t)slot * kInStride, body, len);
    B.lens[slot] = len;
    B.filled.fetch_add(1, std::memory_order_release);
    // wait for the batch
    for (int spin = 0; spin < 64 && !B.done.load(std::memory_order_acquire); spin++) cpu_relax();
    while (!B.done.load(std::memory_order_acquire)) futex(&B.done, FUTEX_WAIT, 0);
    int rc = 0;
    if (B.rc == 0) {
      r = B.res[slot];
      if (r.status == AIGW_OK) {
        const uint32_t bytes = (uint32_t)r.path_len + r.body_len;
        if (bytes > out_cap) { r.status = AIGW_DECLINED; r.reason = AIGW_R_OUT_SPACE; r.body_len = bytes; rc = -4; }   // body_len = the size to retry with
        else memcpy(out, B.out + r.out_off, bytes);
        r.out_off = 0;
      }
    } else { r.status = AIGW_DECLINED; r.reason = AIGW_R_ARENA_FULL; }   // the launch failed: the stock path takes the request
    if (B.readers.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      const uint64_t w = B.word.load(std::memory_order_relaxed);
      B.word.store(mk(w_gen(w), ST_FREE, 0), std::memory_order_release);
      active.fetch_sub(1, std::memory_order_acq_rel);
      freed.fetch_add(1, std::memory_order_release); futex(&freed, FUTEX_WAKE, INT_MAX);
    }
    if (res) *res = r;
    return rc;
  }
};

extern "C" {

int aigw_batcher_start(aigw_ctx* ctx, const aigw_backend_cfg* cfg, uint32_t max_batch, uint32_t window_us, aigw_batcher** out) {
  if (!ctx || !cfg || !out) return -2;
  aigw_batcher* b = new aigw_batcher();
  b->ctx = ctx; b->device = aigw::ctx_device(ctx);
  b->max_batch = max_batch ? (max_batch > 4096 ? 4096 : max_batch) : 256; b->window_us = window_us;
  const int l0 = b->add_lane(cfg);
  if (l0 != 0) { delete b; return l0 < 0 ? l0 : -3; }
  b->th = std::thread([b] { b->run(); });
  *out = b;
  return 0;
}

int aigw_batcher_add_backend(aigw_batcher* b, const aigw_backend_cfg* cfg) { return (b && cfg) ? b->add_lane(cfg) : -2; }

int aigw_batcher_translate(aigw_batcher* b, const uint8_t* body, uint32_t len, uint8_t* out, uint32_t out_cap, aigw_doc_result* res) {
  return b ? b->translate(0, body, len, out, out_cap, res) : -2;
}
int aigw_batcher_translate_to(aigw_batcher* b, int backend, const uint8_t* body, uint32_t len, uint8_t* out, uint32_t out_cap, aigw_doc_result* res) {
  return b ? b->translate(backend, body, len, out, out_cap, res) : -2;
}

int aigw_batcher_get_stats(aigw_batcher* b, aigw_batcher_stats* s) {
  if (!b || !s) return -2;
  s->batches = b->n_batches.load(); s->requests = b->n_docs.load(); s->max_batch_seen = (uint32_t)b->max_seen.load(); s->_pad = 0;
  return 0;
}

void aigw_batcher_stop(aigw_batcher* b) {
  if (!b) return;
  b->stop.store(true);
  b->kick.fetch_add(1); futex(&b->kick, FUTEX_WAKE, 1);
  // open batches are closed by the dispatcher (stop makes every open batch due); wait until nothing is active
  while (b->active.load() != 0) { b->kick.fetch_add(1); futex(&b->kick, FUTEX_WAKE, 1); std::this_thread::sleep_for(std::chrono::microseconds(50)); }
  if (b->th.joinable()) b->th.join();
  b->freed.fetch_add(1); futex(&b->freed, FUTEX_WAKE, INT_MAX);
  cudaSetDevice(b->device);
  for (int i = 0; i < b->n_lanes.load(); i++) {
    Lane* L = b->lanes[i];
    for (int k = 0; k < kRing; k++) { if (L->ring[k].st) cudaStreamDestroy(L->ring[k].st); if (L->ring[k].ev) cudaEventDestroy(L->ring[k].ev); }
    if (L->pinned) cudaFreeHost(L->pinned);
    delete L;
  }
  delete b;
}

}  // extern "C"
