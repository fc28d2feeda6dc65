// TEST INFRASTRUCTURE (oracle): CPU restatement of the response side of the native /v1/messages translators.  Only tests/,
// __graft_entry__.smoke() and bench.py's CPU legs may use it; the product never does.
//
//   anthropicToAnthropicTranslator.ResponseBody       internal/translator/anthropic_anthropic.go:86-130 (buffered + stream branch)
//   extractUsageFromBufferEvent / reflectStreamingEvent / updateTotalTokens   :133-201
//   anthropic.MessagesStreamChunk.UnmarshalJSON       internal/apischema/anthropic/anthropic.go:1696-1748
//   metrics.ExtractTokenUsageFromExplicitCaching      internal/metrics/metrics.go:292-307
//
// The body is never rewritten (nil mutation); the calls return the accumulated TokenUsage and the response model.
// Only message_start and message_delta events touch the state; every other line has no effect whether it decodes or not.
// Those two are decoded over a restated subset — a message_start whose `content` is empty, integer token counts — and the
// stream answers DECLINED (sticky) outside it, because the content-block unions are not restated.
// Pinned by the data-plane cases "anthropic - /anthropic/v1/messages[ - streaming]" (pass-through bodies) and by
// anthropic_anthropic_test.go's usage tables through tests/test_oracle_golden.py.  Parity unpinned: nothing is emitted.
#pragma once
#include "stream.hpp"

namespace oracle {

struct NativeAnthropicStream {
  std::string buffered, request_model, response_model;
  TokenUsage u;
  bool dead = false;
};

namespace native {

// a token count: integer literal (optional '-'), magnitude below 2^31; null counts as 0.  0 ok, 1 outside the subset
inline int count_field(const Value* v, int64_t& out) {
  out = 0;
  if (!v || v->is_null()) return 0;
  if (!v->is_num()) return 1;
  const std::string& s = v->s;
  size_t i = 0; bool neg = false;
  if (i < s.size() && s[i] == '-') { neg = true; i++; }
  if (i >= s.size()) return 1;
  int64_t x = 0;
  for (; i < s.size(); i++) { if (s[i] < '0' || s[i] > '9') return 1; x = x * 10 + (s[i] - '0'); if (x >= (1ll << 31)) return 1; }
  out = neg ? -x : x;
  return 0;
}
struct UsageFields { int64_t in = 0, out = 0, rd = 0, cr = 0; };
inline int usage_obj(const Value& v, UsageFields& f) {
  if (!v.is_obj()) return 1;
  int seen = 0;
  for (auto& kv : v.obj) {
    int bit = kv.first == "input_tokens" ? 1 : kv.first == "output_tokens" ? 2 : kv.first == "cache_read_input_tokens" ? 4 : kv.first == "cache_creation_input_tokens" ? 8 : 0;
    if (!bit) continue;                 // unknown members are ignored by the decoder
    if (seen & bit) return 1;
    seen |= bit;
    int64_t x;
    if (count_field(&kv.second, x)) return 1;
    if (bit == 1) f.in = x; else if (bit == 2) f.out = x; else if (bit == 4) f.rd = x; else f.cr = x;
  }
  return 0;
}
inline bool str_or_null(const Value& v) { return v.is_str() || v.is_null(); }

// one `data: ` payload: 0 handled (or ignored), 1 outside the restated subset
inline int event(NativeAnthropicStream& S, std::string_view payload) {
  Value root; std::string err;
  if (!oj::parse(payload, root, err)) return 0;            // "continue"
  if (!root.is_obj()) return 0;                            // gjson finds no type: "missing type field"
  const Value* ty = root.get_first("type");
  if (!ty) return 0;
  if (!ty->is_str()) return 1;
  auto dup_known = [](const Value& o, std::initializer_list<const char*> known) {   // a member the decoder reads, spelled twice
    for (const char* k : known) { int c = 0; for (auto& kv : o.obj) if (kv.first == k) c++; if (c > 1) return true; }
    return false;
  };
  if (dup_known(root, {"type", "message", "usage", "delta"})) return 1;
  if (ty->s == "message_start") {
    const Value* m = root.get_first("message");
    if (!m) return 0;                                      // decoder.Decode("") fails: the event is skipped
    if (m->is_null()) return 0;                            // zero message: nothing to take
    if (!m->is_obj()) return 1;
    if (dup_known(*m, {"id", "stop_reason", "stop_sequence", "model", "type", "role", "content", "usage"})) return 1;
    std::string model; const Value* usage = nullptr;
    for (auto& kv : m->obj) {
      const std::string& k = kv.first; const Value& v = kv.second;
      if (k == "id" || k == "stop_reason" || k == "stop_sequence") { if (!str_or_null(v)) return 1; }
      else if (k == "model") { if (!str_or_null(v)) return 1; if (v.is_str()) model = v.s; }
      else if (k == "type") { if (!(v.is_str() && v.s == "message")) return 1; }
      else if (k == "role") { if (!(v.is_str() && v.s == "assistant")) return 1; }
      else if (k == "content") { if (!(v.is_null() || (v.is_arr() && v.arr.empty()))) return 1; }
      else if (k == "usage") { if (!v.is_null()) usage = &v; }
    }
    UsageFields f;
    if (usage && usage_obj(*usage, f)) return 1;
    if (usage && (f.in < 0 || f.out < 0 || f.rd < 0 || f.cr < 0 || f.in + f.rd + f.cr >= (1ll << 31) || f.in + f.rd + f.cr + f.out >= (1ll << 31))) return 1;
    if (!model.empty()) S.response_model = model;
    if (usage) {   // ExtractTokenUsageFromExplicitCaching + Override
      TokenUsage t; t.input = (uint32_t)(f.in + f.rd + f.cr); t.output = (uint32_t)f.out; t.total = t.input + t.output; t.cached = (uint32_t)f.rd; t.cache_creation = (uint32_t)f.cr;
      t.mask = TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL | TokenUsage::CACHED | TokenUsage::CACHE_CREATION;
      S.u.override_with(t);
    }
    return 0;
  }
  if (ty->s == "message_delta") {
    UsageFields f;
    if (const Value* us = root.get_first("usage")) { if (!us->is_null() && usage_obj(*us, f)) return 1; }
    if (const Value* d = root.get_first("delta")) {
      if (!d->is_null()) { if (!d->is_obj()) return 1; if (dup_known(*d, {"stop_reason", "stop_sequence"})) return 1; for (auto& kv : d->obj) if ((kv.first == "stop_reason" || kv.first == "stop_sequence") && !str_or_null(kv.second)) return 1; }
    }
    if (f.out >= 0) { S.u.output = (uint32_t)f.out; S.u.mask |= TokenUsage::OUT; }
    return 0;
  }
  return 0;
}

inline void update_total(NativeAnthropicStream& S) {   // updateTotalTokens, anthropic_anthropic.go:176-201
  TokenUsage& u = S.u;
  const bool out_set = u.mask & TokenUsage::OUT;
  if (out_set && !(u.mask & TokenUsage::IN)) { u.input = 0; u.mask |= TokenUsage::IN; }
  if (out_set) {
    if (!(u.mask & TokenUsage::CACHED)) { u.cached = 0; u.mask |= TokenUsage::CACHED; }
    if (!(u.mask & TokenUsage::CACHE_CREATION)) { u.cache_creation = 0; u.mask |= TokenUsage::CACHE_CREATION; }
  }
  if ((u.mask & TokenUsage::IN) && out_set) { u.total = u.input + u.output; u.mask |= TokenUsage::TOTAL; }
}

}  // namespace native

// one ResponseBody(stream) call: OK or DECLINED (sticky); usage / model = what the call returns
inline Status native_anthropic_feed(NativeAnthropicStream& S, std::string_view bytes, TokenUsage& usage, std::string& model) {
  if (S.dead) return DECLINED;
  S.buffered.append(bytes);
  for (;;) {
    const size_t i = S.buffered.find('\n');
    if (i == std::string::npos) break;
    const std::string line = S.buffered.substr(0, i);
    S.buffered.erase(0, i + 1);
    if (line.rfind("data: ", 0) != 0) continue;
    if (native::event(S, std::string_view(line).substr(6))) { S.dead = true; return DECLINED; }
  }
  native::update_total(S);
  usage = S.u; model = S.response_model.empty() ? S.request_model : S.response_model;
  return OK;
}

// buffered anthropic.MessagesResponse: usage + response model (anthropic_anthropic.go:103-130)
inline Status native_anthropic_response(std::string_view body, const std::string& request_model, TokenUsage& usage, std::string& model) {
  // json.NewDecoder(body).Decode: the first JSON value, trailing bytes ignored
  Value root;
  { oj::Parser ps(body.data(), body.size()); ps.ws(); if (!ps.value(root)) return INTERNAL; }   // "failed to unmarshal body"
  if (root.is_null()) { usage = TokenUsage(); usage.mask = 31; model = request_model; return OK; }
  if (!root.is_obj()) return INTERNAL;
  for (const char* k : {"id", "stop_reason", "stop_sequence", "model", "type", "role", "content", "usage"}) { int c = 0; for (auto& kv : root.obj) if (kv.first == k) c++; if (c > 1) return DECLINED; }
  std::string m; const Value* us = nullptr;
  for (auto& kv : root.obj) {
    const std::string& k = kv.first; const Value& v = kv.second;
    if (k == "id" || k == "stop_reason" || k == "stop_sequence") { if (!native::str_or_null(v)) return DECLINED; }
    else if (k == "model") { if (!native::str_or_null(v)) return DECLINED; if (v.is_str()) m = v.s; }
    else if (k == "type") { if (!(v.is_str() && v.s == "message")) return DECLINED; }
    else if (k == "role") { if (!(v.is_str() && v.s == "assistant")) return DECLINED; }
    else if (k == "content") {   // MessagesContentBlock.UnmarshalJSON (anthropic.go:1505-1557) over the block types text / tool_use /
      if (v.is_null()) continue;  // server_tool_use / thinking / redacted_thinking with members of the declared JSON types; the rest is DECLINED
      if (!v.is_arr()) return DECLINED;
      for (auto& b : v.arr) {
        if (!b.is_obj()) return DECLINED;
        for (const char* kk : {"type", "text", "citations", "id", "name", "input", "thinking", "signature", "data"}) { int c = 0; for (auto& kv2 : b.obj) if (kv2.first == kk) c++; if (c > 1) return DECLINED; }
        const Value* t = b.get_first("type");
        if (!t || !t->is_str()) return DECLINED;
        auto son = [&](const char* kk) { const Value* x = b.get_first(kk); return !x || x->is_null() || x->is_str(); };
        if (t->s == "text") { const Value* c = b.get_first("citations"); if (!son("text") || (c && !c->is_null())) return DECLINED; }
        else if (t->s == "tool_use" || t->s == "server_tool_use") { const Value* in = b.get_first("input"); if (!son("id") || !son("name") || (in && !in->is_null() && !in->is_obj())) return DECLINED; }
        else if (t->s == "thinking") { if (!son("thinking") || !son("signature")) return DECLINED; }
        else if (t->s == "redacted_thinking") { if (!son("data")) return DECLINED; }
        else return DECLINED;
      }
    }
    else if (k == "usage") { if (!v.is_null()) us = &v; }
  }
  native::UsageFields f;
  if (us && native::usage_obj(*us, f)) return DECLINED;
  if (f.in < 0 || f.out < 0 || f.rd < 0 || f.cr < 0 || f.in + f.rd + f.cr + f.out >= (1ll << 31)) return DECLINED;
  usage = TokenUsage();
  usage.input = (uint32_t)(f.in + f.rd + f.cr); usage.output = (uint32_t)f.out; usage.total = usage.input + usage.output; usage.cached = (uint32_t)f.rd; usage.cache_creation = (uint32_t)f.cr;
  usage.mask = 31;
  model = m.empty() ? request_model : m;
  return OK;
}

}  // namespace oracle
