// Load generator for the request batcher: T threads, each submitting its share of the bodies through aigw_batcher_translate
// (what T goroutines in the cgo shim would do).  Prints one JSON line: throughput, latency percentiles, batch statistics.
// usage: batcher_load <bodies.bin> <offsets.u64> <n> <threads> <requests_per_thread> <max_batch> <window_us> [schema]
// Build: g++ -O2 -std=c++17 -I include tools/batcher_load.cpp -L aigw_b200 -laigw_b200 -Wl,-rpath,$PWD/aigw_b200 -lpthread -o /tmp/batcher_load
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "aigw_b200.h"

static std::vector<uint8_t> slurp(const char* p) { FILE* f = fopen(p, "rb"); if (!f) { perror(p); exit(2); } fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<uint8_t> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(2); fclose(f); return v; }

int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
  std::vector<uint8_t> bodies = slurp(argv[1]), offb = slurp(argv[2]);
  const uint64_t* offs = (const uint64_t*)offb.data();
  const uint32_t n = atoi(argv[3]); const int T = atoi(argv[4]); const int R = atoi(argv[5]);
  const uint32_t max_batch = atoi(argv[6]), window_us = atoi(argv[7]);
  const int schema = argc > 8 ? atoi(argv[8]) : AIGW_SCHEMA_AWS_BEDROCK;
  aigw_ctx* ctx = nullptr;
  if (int rc = aigw_init(0, &ctx)) { fprintf(stderr, "aigw_init: %d\n", rc); return 1; }
  aigw_backend_cfg cfg; memset(&cfg, 0, sizeof cfg); cfg.schema = schema;
  aigw_batcher* b = nullptr;
  if (aigw_batcher_start(ctx, &cfg, max_batch, window_us, &b)) return 1;
  std::vector<std::vector<float>> lat(T);
  std::atomic<uint64_t> ok{0}, declined{0}, failed{0}, out_bytes{0};
  std::atomic<int> first_reason{-1}, first_rc{0};
  auto worker = [&](int t, int reqs, bool record) {
    std::vector<uint8_t> out(1 << 17);
    for (int r = 0; r < reqs; r++) {
      const uint32_t i = (uint32_t)((uint64_t)t * 7919u + (uint64_t)r * 104729u) % n;
      aigw_doc_result res;
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = aigw_batcher_translate(b, bodies.data() + offs[i], (uint32_t)(offs[i + 1] - offs[i]), out.data(), (uint32_t)out.size(), &res);
      const float us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (record) lat[t].push_back(us);
      if (rc) { failed++; first_rc = rc; } else if (res.status == AIGW_OK) { ok++; out_bytes += res.path_len + res.body_len; } else { declined++; int e = -1; first_reason.compare_exchange_strong(e, (int)res.reason * 16 + (int)res.status); }
    }
  };
  { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(worker, t, std::max(1, R / 10), false); for (auto& x : th) x.join(); }  // warm-up
  ok = 0; declined = 0; failed = 0; out_bytes = 0;
  aigw_batcher_stats s0; aigw_batcher_get_stats(b, &s0);
  const auto w0 = std::chrono::steady_clock::now();
  { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(worker, t, R, true); for (auto& x : th) x.join(); }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
  aigw_batcher_stats s1; aigw_batcher_get_stats(b, &s1);
  std::vector<float> all; for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end());
  auto pct = [&](double p) { return all.empty() ? 0.f : all[(size_t)(p * (all.size() - 1))]; };
  printf("{\"threads\": %d, \"requests\": %zu, \"ok\": %llu, \"declined\": %llu, \"failed\": %llu, \"wall_s\": %.4f, \"requests_per_s\": %.1f, "
         "\"p50_us\": %.1f, \"p90_us\": %.1f, \"p99_us\": %.1f, \"max_us\": %.1f, \"batches\": %llu, \"mean_batch\": %.2f, \"max_batch_seen\": %u, \"max_batch\": %u, \"window_us\": %u, \"first_declined_reason_status\": [%d, %d], \"first_rc\": %d}\n",
         T, all.size(), (unsigned long long)ok.load(), (unsigned long long)declined.load(), (unsigned long long)failed.load(), wall, all.size() / wall,
         pct(0.5), pct(0.9), pct(0.99), all.empty() ? 0.f : all.back(), (unsigned long long)(s1.batches - s0.batches),
         (s1.batches - s0.batches) ? (double)(s1.requests - s0.requests) / (double)(s1.batches - s0.batches) : 0.0, s1.max_batch_seen, max_batch, window_us, first_reason.load() < 0 ? -1 : first_reason.load() / 16, first_reason.load() < 0 ? -1 : first_reason.load() % 16, first_rc.load());
  aigw_batcher_stop(b);
  aigw_destroy(ctx);
  return failed.load() ? 1 : 0;
}
