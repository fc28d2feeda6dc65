// aigw_b200 — per-GPU request batcher: the synchronous single-request call a cgo shim makes from thousands of goroutines
// ("a synchronous single-doc wrapper for the Go shim sits on top of the batch queue"; the reference calls
// ParseBody / RequestBody once per request on the request's own goroutine, internal/extproc/processor_impl.go:211-398).
//
// Submitters enqueue (pointer, length) under a mutex and sleep on their ticket; one batcher thread per context wakes on the
// first request, keeps collecting until `max_batch` requests are queued or `window_us` has passed since the first one, packs
// the bodies into the pinned input arena, runs ONE aigw_chat_translate_host for the whole batch and wakes the tickets; each
// woken submitter copies its own record out of the shared output arena, so the copies are not serialised on the batcher thread.  A slow or failed batch fails every ticket in it with the CUDA error code; there
// is no CPU path here either.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aigw_b200.h"

namespace {
// (A wake tree — each woken ticket waking four more — was measured: same median, worse tails than waking every ticket
// from the batcher thread, so the batcher wakes them all itself.)
struct Ticket {
  const uint8_t* body; uint32_t len;
  uint8_t* out; uint32_t out_cap;
  // filled by the batcher thread before the wake
  const aigw_doc_result* src_res = nullptr; const uint8_t* src_out = nullptr; int batch_rc = 0;
  std::atomic<uint32_t>* pending = nullptr;
  bool woken = false;
  std::mutex m; std::condition_variable cv;
};
inline void wake(Ticket* t) { std::lock_guard<std::mutex> g(t->m); t->woken = true; t->cv.notify_one(); }
}  // namespace

struct aigw_batcher {
  aigw_ctx* ctx = nullptr;
  aigw_backend_cfg cfg{};
  std::string s_override, s_prefix, s_version, s_rid;
  uint32_t max_batch = 256, window_us = 50;
  std::mutex m; std::condition_variable cv;
  std::deque<Ticket*> q;
  bool stop = false;
  std::thread th;
  uint8_t* arena = nullptr; size_t arena_cap = 0;   // pinned staging for the packed bodies
  std::vector<uint64_t> offs; std::vector<uint32_t> lens;
  std::atomic<uint32_t> pending{0};                 // tickets of the last batch that have not copied their record out yet
  std::atomic<uint64_t> n_batches{0}, n_docs{0}, max_seen{0};

  void run() {
    std::vector<Ticket*> batch;
    for (;;) {
      batch.clear();
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (stop && q.empty()) return;
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us);
        while (q.size() < max_batch && !stop) { if (cv.wait_until(lk, deadline) == std::cv_status::timeout) break; }
        while (!q.empty() && batch.size() < max_batch) { batch.push_back(q.front()); q.pop_front(); }
      }
      process(batch);
    }
  }
  // hand the batch's outcome to its tickets
  void release(std::vector<Ticket*>& batch, int rc, const aigw_batch_out* bo) {
    const size_t n = batch.size();
    pending.store((uint32_t)n, std::memory_order_release);
    for (size_t i = 0; i < n; i++) {
      Ticket* t = batch[i];
      t->batch_rc = rc; t->src_res = bo ? bo->results + i : nullptr; t->src_out = bo ? bo->out : nullptr; t->pending = &pending;
    }
    for (size_t i = 0; i < n; i++) wake(batch[i]);
    // the output arena belongs to the next GPU call only after every ticket has copied its record
    while (pending.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }
  void process(std::vector<Ticket*>& batch) {
    const uint32_t n = (uint32_t)batch.size();
    offs.resize(n); lens.resize(n);
    uint64_t o = 0;
    for (uint32_t i = 0; i < n; i++) { offs[i] = o; lens[i] = batch[i]->len; o += ((uint64_t)batch[i]->len + 15u) & ~15ull; }
    if (o + 64 > arena_cap) {
      if (arena) aigw_host_free(ctx, arena);
      arena_cap = (size_t)(o + 64) * 2; arena = (uint8_t*)aigw_host_alloc(ctx, arena_cap);
      if (!arena) { arena_cap = 0; release(batch, -3, nullptr); return; }
    }
    for (uint32_t i = 0; i < n; i++) { memcpy(arena + offs[i], batch[i]->body, batch[i]->len); const uint64_t pad = ((((uint64_t)lens[i] + 15u) & ~15ull)) - lens[i]; memset(arena + offs[i] + lens[i], ' ', (size_t)pad); }
    memset(arena + o, ' ', 16);
    aigw_batch_out bo;
    const int rc = aigw_chat_translate_host(ctx, &cfg, arena, offs.data(), lens.data(), n, &bo);
    if (!rc) { n_batches++; n_docs += n; if (n > max_seen.load()) max_seen = n; }
    release(batch, rc, rc ? nullptr : &bo);
  }
};

extern "C" {

int aigw_batcher_start(aigw_ctx* ctx, const aigw_backend_cfg* cfg, uint32_t max_batch, uint32_t window_us, aigw_batcher** out) {
  if (!ctx || !cfg || !out) return -2;
  aigw_batcher* b = new aigw_batcher();
  b->ctx = ctx; b->cfg = *cfg;
  // the configuration strings must outlive the caller's
  auto keep = [](std::string& dst, const char*& p) { if (p) { dst = p; p = dst.c_str(); } };
  keep(b->s_override, b->cfg.model_name_override); keep(b->s_prefix, b->cfg.openai_prefix); keep(b->s_version, b->cfg.api_version); keep(b->s_rid, b->cfg.response_id);
  b->max_batch = max_batch ? max_batch : 256; b->window_us = window_us;
  b->th = std::thread([b] { b->run(); });
  *out = b;
  return 0;
}

int aigw_batcher_translate(aigw_batcher* b, const uint8_t* body, uint32_t len, uint8_t* out, uint32_t out_cap, aigw_doc_result* res) {
  Ticket t; t.body = body; t.len = len; t.out = out; t.out_cap = out_cap;
  {
    std::lock_guard<std::mutex> g(b->m);
    if (b->stop) return -5;
    b->q.push_back(&t);
  }
  b->cv.notify_one();
  { std::unique_lock<std::mutex> lk(t.m); t.cv.wait(lk, [&] { return t.woken; }); }
  int rc = t.batch_rc;
  aigw_doc_result r; memset(&r, 0, sizeof r);
  if (!rc) {
    r = *t.src_res;
    if (r.status == AIGW_OK) {
      const uint32_t bytes = (uint32_t)r.path_len + r.body_len;
      if (bytes > out_cap) rc = -4;   // caller's buffer too small
      else { memcpy(out, t.src_out + r.out_off, bytes); r.out_off = 0; }
    }
  }
  t.pending->fetch_sub(1, std::memory_order_acq_rel);
  if (res) *res = r;
  return rc;
}

int aigw_batcher_get_stats(aigw_batcher* b, aigw_batcher_stats* s) {
  s->batches = b->n_batches.load(); s->requests = b->n_docs.load(); s->max_batch_seen = (uint32_t)b->max_seen.load(); s->_pad = 0;
  return 0;
}

void aigw_batcher_stop(aigw_batcher* b) {
  if (!b) return;
  { std::lock_guard<std::mutex> g(b->m); b->stop = true; }
  b->cv.notify_all();
  if (b->th.joinable()) b->th.join();
  if (b->arena) aigw_host_free(b->ctx, b->arena);
  delete b;
}

}  // extern "C"
