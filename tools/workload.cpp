// Synthetic workload generators for BASELINE.json's configs.  Test/bench tooling:
// not part of the product path and not part of the oracle.  Deterministic per (seed, index), so
// any rank can generate its own shard.
//   C1/C2  wl_chat_*  OpenAI ChatCompletion bodies (target bytes ± jitter, 4–8 messages, 10 % with a
//                     tool definition + assistant tool call + tool result, the shape of
//                     tests/data-plane/testupstream_test.go:1570-1596)
//   C4     wl_sse_*   OpenAI SSE streams: 80-byte events, usage in the penultimate event, [DONE] last,
//                     20 % of the streams with chunk boundaries that do not fall on event boundaries
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
  double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};
uint64_t mix(uint64_t a, uint64_t b) { Rng r(a * 0x9e3779b97f4a7c15ull + b); r.next(); return r.next(); }

const char* kWords[] = {"the", "of", "and", "to", "in", "is", "that", "for", "it", "as", "was", "with", "be", "by", "on", "not", "he", "this", "are", "or", "his",
  "from", "at", "which", "but", "have", "an", "had", "they", "you", "were", "their", "one", "all", "we", "can", "her", "has", "there", "been", "if", "more", "when",
  "will", "would", "who", "so", "no", "gateway", "request", "response", "token", "model", "stream", "latency", "backend", "translate", "schema", "payload", "bedrock",
  "converse", "kernel", "throughput", "bandwidth", "buffer", "pipeline", "summarize", "following", "document", "please", "explain", "difference", "between", "example",
  "function", "return", "value", "error", "context", "assistant", "question", "answer", "detail", "because", "however", "therefore", "analysis", "results"};
const int kNumWords = sizeof(kWords) / sizeof(kWords[0]);
const char* kMulti[] = {"\xC3\xA9", "\xC3\xBC", "\xE4\xB8\xAD", "\xE6\x96\x87", "\xE2\x82\xAC", "\xF0\x9F\x99\x82", "\xC3\xB1", "\xE3\x81\x82"};

// exactly `n` bytes of JSON-string-safe text: Zipf-ish words, ~5 % multi-byte UTF-8, ~1 % escapes (\n, \")
void text(std::string& o, size_t n, Rng& r) {
  size_t start = o.size();
  while (o.size() - start < n) {
    size_t left = n - (o.size() - start);
    uint32_t roll = r.below(100);
    if (roll < 5 && left >= 5) { o += kMulti[r.below(8)]; o.push_back(' '); }
    else if (roll < 6 && left >= 3) { o += (r.below(2) ? "\\n" : "\\\""); }
    else {
      double u = r.unit(); int w = (int)(u * u * kNumWords);  // skewed toward frequent words
      const char* s = kWords[w]; size_t l = strlen(s);
      if (l + 1 <= left) { o += s; o.push_back(' '); } else o.push_back('x');
    }
    // multi-byte / escape pieces may overshoot by at most 4 bytes: trim by rewinding
    if (o.size() - start > n) { o.resize(start + n - (n - (start + n > o.size() ? 0 : 0))); }
  }
  if (o.size() - start > n) o.resize(start + n);
  // never end on a dangling backslash or a cut multi-byte sequence
  size_t e = o.size();
  while (e > start && ((unsigned char)o[e - 1] >= 0x80 || o[e - 1] == '\\')) e--;
  for (size_t i = e; i < o.size(); i++) o[i] = '.';
}

struct ChatShape { uint32_t target; int n_msgs; bool tools; bool stream; };
ChatShape shape(uint64_t seed, uint64_t idx, uint32_t target, uint32_t jitter) {
  Rng r(mix(seed, idx));
  ChatShape s;
  s.target = target - jitter + r.below(2 * jitter + 1);
  s.n_msgs = 4 + (int)r.below(5);
  s.tools = r.below(10) == 0;
  s.stream = r.below(4) == 0;
  return s;
}

void chat_body(std::string& o, uint64_t seed, uint64_t idx, uint32_t target, uint32_t jitter) {
  ChatShape s = shape(seed, idx, target, jitter);
  Rng r(mix(seed ^ 0x5151, idx));
  // structure first, with %N placeholders for the free-text slots
  struct Slot { size_t pos; };
  std::vector<size_t> slots;
  std::string t = "{\"model\":\"gpt-4o-mini\",\"messages\":[";
  auto add_msg = [&](const char* role) { if (t.back() != '[') t += ","; t += "{\"role\":\""; t += role; t += "\",\"content\":\""; slots.push_back(t.size()); t += "\"}"; };
  add_msg("system");
  int body_msgs = s.n_msgs - 1;
  if (s.tools) {
    add_msg("user");
    t += ",{\"role\":\"assistant\",\"content\":null,\"tool_calls\":[{\"id\":\"call_"; t += std::to_string(100000 + r.below(900000));
    std::string id = t.substr(t.size() - 11);
    t += "\",\"type\":\"function\",\"function\":{\"name\":\"list_files\",\"arguments\":\"{\\\"path\\\": \\\"/tmp/d"; t += std::to_string(r.below(1000)); t += "\\\", \\\"depth\\\": "; t += std::to_string(1 + r.below(9)); t += "}\"}}]}";
    t += ",{\"role\":\"tool\",\"tool_call_id\":\""; t += id; t += "\",\"content\":\""; slots.push_back(t.size()); t += "\"}";
    body_msgs -= 3;
    for (int k = 0; k < body_msgs; k++) add_msg((k & 1) ? "user" : "assistant");
    if (body_msgs <= 0 || ((body_msgs - 1) & 1) == 0) add_msg("user");
  } else {
    for (int k = 0; k < body_msgs; k++) add_msg((k & 1) ? "assistant" : "user");
    if (((body_msgs - 1) & 1) == 1) add_msg("user");
  }
  t += "]";
  if (s.tools) t += ",\"tools\":[{\"type\":\"function\",\"function\":{\"name\":\"list_files\",\"description\":\"List the files in a directory\",\"parameters\":{\"type\":\"object\",\"properties\":{\"path\":{\"type\":\"string\",\"description\":\"Directory to list\"},\"depth\":{\"type\":\"integer\"}},\"required\":[\"path\"]}}}],\"tool_choice\":\"auto\"";
  t += ",\"max_completion_tokens\":"; t += std::to_string(64 << r.below(5));
  t += ",\"temperature\":0.7";
  if (r.below(3) == 0) t += ",\"top_p\":0.95";
  if (s.stream) t += ",\"stream\":true";
  t += "}";
  // distribute the remaining bytes over the slots
  size_t fixed = t.size();
  size_t free_bytes = s.target > fixed + slots.size() * 8 ? s.target - fixed : slots.size() * 8;
  std::vector<size_t> share(slots.size());
  { double tot = 0; std::vector<double> w(slots.size()); for (auto& x : w) { x = 0.25 + r.unit(); tot += x; }
    size_t used = 0; for (size_t k = 0; k < slots.size(); k++) { share[k] = (size_t)(free_bytes * (w[k] / tot)); used += share[k]; }
    share[0] += free_bytes - used; }
  o.clear(); o.reserve(s.target + 64);
  size_t prev = 0;
  for (size_t k = 0; k < slots.size(); k++) { o.append(t, prev, slots[k] - prev); text(o, share[k], r); prev = slots[k]; }
  o.append(t, prev, std::string::npos);
}

void sse_stream(std::string& o, std::vector<uint32_t>& chunk_ends, uint64_t seed, uint64_t idx, int chunks, int chunk_bytes) {
  Rng r(mix(seed, idx));
  o.clear(); chunk_ends.clear();
  const std::string head = "data: {\"model\":\"gpt-4o-mini\",\"choices\":[{\"index\":0,\"delta\":{\"content\":\"";
  const std::string tail = "\"}}]}\n\n";
  std::vector<uint32_t> event_ends;
  int events = chunks - 2;
  for (int e = 0; e < events; e++) {
    o += head;
    int pad = chunk_bytes - (int)head.size() - (int)tail.size(); if (pad < 1) pad = 1;
    for (int k = 0; k < pad; k++) o.push_back("abcdefghijklmnopqrstuvwxyz "[r.below(27)]);
    o += tail; event_ends.push_back((uint32_t)o.size());
  }
  uint32_t pt = 10 + r.below(4000), ct = 1 + r.below(2000);
  o += "data: {\"model\":\"gpt-4o-mini\",\"choices\":[],\"usage\":{\"prompt_tokens\":" + std::to_string(pt) + ",\"completion_tokens\":" + std::to_string(ct) + ",\"total_tokens\":" + std::to_string(pt + ct);
  if (r.below(2)) o += ",\"prompt_tokens_details\":{\"cached_tokens\":" + std::to_string(r.below(pt)) + "}";
  if (r.below(4) == 0) o += ",\"completion_tokens_details\":{\"reasoning_tokens\":" + std::to_string(r.below(ct)) + "}";
  o += "}}\n\n"; event_ends.push_back((uint32_t)o.size());
  o += "data: [DONE]\n\n"; event_ends.push_back((uint32_t)o.size());
  if (r.below(5) == 0) {  // ragged: same number of chunks, boundaries anywhere
    uint32_t n = (uint32_t)o.size(); uint32_t prev = 0;
    for (int c = 0; c < chunks - 1; c++) { uint32_t left = n - prev; uint32_t remain = chunks - c; uint32_t sz = 1 + r.below(2 * left / remain > 1 ? 2 * left / remain - 1 : 1); if (prev + sz > n - (remain - 1)) sz = 1; prev += sz; chunk_ends.push_back(prev); }
    chunk_ends.push_back(n);
  } else chunk_ends = event_ends;
}
}  // namespace

extern "C" {
// lens[i] = exact byte length of body first+i
void wl_chat_lens(uint64_t seed, uint64_t first, uint32_t n, uint32_t target, uint32_t jitter, uint32_t* lens, int threads) {
  auto work = [&](uint32_t b, uint32_t e) { std::string s; for (uint32_t i = b; i < e; i++) { chat_body(s, seed, first + i, target, jitter); lens[i] = (uint32_t)s.size(); } };
  if (threads <= 1) { work(0, n); return; }
  std::vector<std::thread> th; uint32_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) { uint32_t b = t * per, e = b + per > n ? n : b + per; if (b < e) th.emplace_back(work, b, e); }
  for (auto& t : th) t.join();
}
// writes body i at out + offsets[i]; offsets are caller-provided (16-byte aligned), lens from wl_chat_lens
void wl_chat_fill(uint64_t seed, uint64_t first, uint32_t n, uint32_t target, uint32_t jitter, uint8_t* out, const uint64_t* offsets, const uint32_t* lens, int threads) {
  auto work = [&](uint32_t b, uint32_t e) { std::string s; for (uint32_t i = b; i < e; i++) { chat_body(s, seed, first + i, target, jitter); memcpy(out + offsets[i], s.data(), s.size()); size_t pad = ((s.size() + 15) & ~15ull) - s.size(); memset(out + offsets[i] + s.size(), ' ', pad); } };
  if (threads <= 1) { work(0, n); return; }
  std::vector<std::thread> th; uint32_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) { uint32_t b = t * per, e = b + per > n ? n : b + per; if (b < e) th.emplace_back(work, b, e); }
  for (auto& t : th) t.join();
}
// multi-threaded memcpy (the bench's stand-in for the goroutines that copy bodies into / results out of the pinned arenas)
void wl_memcpy_mt(uint8_t* dst, const uint8_t* src, uint64_t n, int threads) {
  if (threads <= 1 || n < (1u << 20)) { memcpy(dst, src, n); return; }
  std::vector<std::thread> th; const uint64_t per = ((n + threads - 1) / threads + 4095) & ~4095ull;
  for (int t = 0; t < threads; t++) { const uint64_t b = (uint64_t)t * per; if (b >= n) break; const uint64_t l = b + per > n ? n - b : per; th.emplace_back([=] { memcpy(dst + b, src + b, l); }); }
  for (auto& t : th) t.join();
}
// SSE: returns total bytes; chunk_off has n*chunks+1 entries (absolute byte offsets); out may be NULL to size
uint64_t wl_sse_fill(uint64_t seed, uint64_t first, uint32_t n, int chunks, int chunk_bytes, uint8_t* out, uint64_t cap, uint64_t* chunk_off) {
  std::string s; std::vector<uint32_t> ends; uint64_t pos = 0; uint64_t c = 0;
  if (chunk_off) chunk_off[0] = 0;
  for (uint32_t i = 0; i < n; i++) {
    sse_stream(s, ends, seed, first + i, chunks, chunk_bytes);
    if (out && pos + s.size() <= cap) memcpy(out + pos, s.data(), s.size());
    for (uint32_t e : ends) { c++; if (chunk_off) chunk_off[c] = pos + e; }
    pos += s.size();
  }
  return pos;
}
}
