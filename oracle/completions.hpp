// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the response side of the legacy /v1/completions endpoint (S1 variant + R1):
//   ResponseBody, stream branch + extractUsageFromBufferEvent   internal/translator/openai_completions.go:80-96,157-203
//   ResponseBody, buffered branch                               internal/translator/openai_completions.go:98-150
//   CompletionResponse / CompletionChoice / CompletionLogprobs  internal/apischema/openai/openai.go:2000-2055
//   Usage                                                       internal/apischema/openai/openai.go:2064-2081
// Pinned by the reference's own known answers: openai_completions_test.go:107-247 (five buffered bodies),
// :249-309 (a stream in three ResponseBody calls), :311-337 (reasoning tokens), lifted into tests/golden/completions_cases.json.
#pragma once
#include "stream.hpp"

namespace oracle {

// json.Unmarshal(data, &openai.CompletionResponse{}) — true when it would succeed
inline bool decode_completion(std::string_view text, std::string& model, bool& has_usage, TokenUsage& tu, std::vector<int64_t>* raw_usage = nullptr) {
  Value v; std::string err;
  if (!oj::parse(text, v, err)) return false;
  model.clear(); has_usage = false;
  if (v.is_null()) return true;
  if (!v.is_obj()) return false;
  for (const char* k : {"id", "object", "model", "system_fingerprint"}) if (!str_or_null(v.get(k))) return false;
  if (const Value* c = v.get("created")) {  // JSONUNIXTime.UnmarshalJSON, called for null too (openai.go:1789-1807)
    std::string raw(text.substr(c->b, c->e - c->b));
    size_t dot = raw.find('.'); if (dot != std::string::npos) raw.resize(dot);
    int64_t q; if (!oj::num_to_i64(raw, q)) return false;
  }
  const Value* ch = v.get("choices");
  if (!arr_or_null(ch)) return false;
  if (ch && ch->is_arr()) for (auto& c : ch->arr) {
    if (c.is_null()) continue;
    if (!c.is_obj()) return false;
    int64_t idx;
    if (!str_or_null(c.get("text")) || !int_field(c.get("index"), idx) || !str_or_null(c.get("finish_reason"))) return false;
    const Value* lp = c.get("logprobs"); if (!obj_or_null(lp)) return false;
    if (lp && lp->is_obj()) {
      const Value* a = lp->get("tokens"); if (!arr_or_null(a)) return false;
      if (a && a->is_arr()) for (auto& t : a->arr) if (!(t.is_null() || t.is_str())) return false;
      a = lp->get("token_logprobs"); if (!arr_or_null(a)) return false;
      if (a && a->is_arr()) for (auto& t : a->arr) if (!(t.is_null() || t.is_num())) return false;
      a = lp->get("text_offset"); if (!arr_or_null(a)) return false;
      if (a && a->is_arr()) for (auto& t : a->arr) { int64_t q; if (!int_field(&t, q)) return false; }
      a = lp->get("top_logprobs"); if (!arr_or_null(a)) return false;   // []map[string]float64
      if (a && a->is_arr()) for (auto& m : a->arr) {
        if (m.is_null()) continue;
        if (!m.is_obj()) return false;
        for (auto& kv : m.obj) if (!(kv.second.is_null() || kv.second.is_num())) return false;
      }
    }
  }
  if (const Value* m = v.get("model"); m && m->is_str()) model = m->s;
  const Value* u = v.get("usage");
  if (!obj_or_null(u)) return false;
  if (u && u->is_obj()) {
    tu = TokenUsage{}; if (!decode_usage(*u, tu)) return false; has_usage = true;
    if (raw_usage) {  // the signed values, for the buffered branch's ">= 0" guards
      auto iv = [&](const Value* o, const char* k) { int64_t x = 0; if (o && o->is_obj()) int_field(o->get(k), x); return x; };
      const Value* ptd = u->get("prompt_tokens_details"); const Value* ctd = u->get("completion_tokens_details");
      *raw_usage = {iv(u, "prompt_tokens"), iv(u, "completion_tokens"), iv(u, "total_tokens"), iv(ptd, "cached_tokens"), iv(ptd, "cache_creation_input_tokens"), iv(ctd, "reasoning_tokens")};
    }
  }
  return true;
}

struct SSECompletionsState {  // openai_completions.go:32-43
  std::string buffered;
  std::string streaming_model;
};

// One ResponseBody(stream) call (openai_completions.go:80-96,157-203): append, scan complete lines, latest usage of THIS call;
// responseModel = streamingResponseModel (no fallback to the request model).
inline TokenUsage sse_completions_feed(SSECompletionsState& st, std::string_view chunk) {
  st.buffered.append(chunk);
  TokenUsage out;
  size_t pos = 0;
  for (;;) {
    size_t nl = st.buffered.find('\n', pos);
    if (nl == std::string::npos) break;
    std::string_view line(st.buffered.data() + pos, nl - pos);
    pos = nl + 1;
    if (line.substr(0, 6) != "data: ") continue;
    std::string_view data = line.substr(6);
    if (data == "[DONE]") continue;
    std::string model; bool has_usage; TokenUsage tu;
    if (!decode_completion(data, model, has_usage, tu)) continue;
    if (!model.empty()) st.streaming_model = model;
    if (has_usage) {  // :188-199
      out.input = tu.input; out.output = tu.output; out.total = tu.total; out.mask |= TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
      if (tu.mask & TokenUsage::CACHED) { out.cached = tu.cached; out.cache_creation = tu.cache_creation; out.mask |= TokenUsage::CACHED | TokenUsage::CACHE_CREATION; }
      if (tu.mask & TokenUsage::REASONING) { out.reasoning = tu.reasoning; out.mask |= TokenUsage::REASONING; }
    }
  }
  st.buffered.erase(0, pos);
  return out;
}

// Buffered branch (openai_completions.go:98-150): json.Unmarshal of the whole body (trailing bytes ARE an error, unlike the
// chat translator's Decoder), responseModel = resp.Model, every counter only when it is >= 0.
inline bool response_completions(std::string_view body, TokenUsage& tu, std::string& response_model) {
  tu = TokenUsage{}; response_model.clear();
  std::string model; bool has_usage = false; TokenUsage du; std::vector<int64_t> raw;
  if (!decode_completion(body, model, has_usage, du, &raw)) return false;
  response_model = model;
  if (has_usage) {
    if (raw[0] >= 0) { tu.input = (uint32_t)raw[0]; tu.mask |= TokenUsage::IN; }
    if (raw[1] >= 0) { tu.output = (uint32_t)raw[1]; tu.mask |= TokenUsage::OUT; }
    if (raw[2] >= 0) { tu.total = (uint32_t)raw[2]; tu.mask |= TokenUsage::TOTAL; }
    if (du.mask & TokenUsage::CACHED) {
      if (raw[3] >= 0) { tu.cached = (uint32_t)raw[3]; tu.mask |= TokenUsage::CACHED; }
      if (raw[4] >= 0) { tu.cache_creation = (uint32_t)raw[4]; tu.mask |= TokenUsage::CACHE_CREATION; }
    }
    if ((du.mask & TokenUsage::REASONING) && raw[5] >= 0) { tu.reasoning = (uint32_t)raw[5]; tu.mask |= TokenUsage::REASONING; }
  }
  return true;
}

}  // namespace oracle
