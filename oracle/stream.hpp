// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the response side for OpenAI-schema backends and of the cost step:
//   S1 extractUsageFromBufferEvent          internal/translator/openai_openai.go:179-215
//   R1 non-stream ResponseBody (OpenAI)     internal/translator/openai_openai.go:146-174
//   ChatCompletionResponseChunk / Usage     internal/apischema/openai/openai.go:1497-1565,2064-2083,1464-1495,1789-1807
//   C1 TokenUsage + Override                internal/metrics/metrics.go:143-158,258-283
//   C2 evalCost (the six direct selectors)  internal/extproc/processor_impl.go:707-756
#pragma once
#include "chat.hpp"

namespace oracle {

struct TokenUsage {  // metrics.go:143-158; mask bit i ⇔ the i-th "…Set" flag
  uint32_t input = 0, output = 0, total = 0, cached = 0, cache_creation = 0, reasoning = 0;
  uint32_t mask = 0;
  enum { IN = 1, OUT = 2, TOTAL = 4, CACHED = 8, CACHE_CREATION = 16, REASONING = 32 };
  void override_with(const TokenUsage& o) {  // metrics.go:258-283
    if (o.mask & IN) { input = o.input; mask |= IN; }
    if (o.mask & OUT) { output = o.output; mask |= OUT; }
    if (o.mask & TOTAL) { total = o.total; mask |= TOTAL; }
    if (o.mask & CACHED) { cached = o.cached; mask |= CACHED; }
    if (o.mask & CACHE_CREATION) { cache_creation = o.cache_creation; mask |= CACHE_CREATION; }
    if (o.mask & REASONING) { reasoning = o.reasoning; mask |= REASONING; }
  }
};

// `int` struct field decode; uint32(x) conversion wraps (Go conversion semantics).
inline bool int_field(const Value* v, int64_t& out) {
  out = 0;
  if (!v || v->is_null()) return true;
  return v->is_num() && oj::num_to_i64(v->s, out);
}
inline bool str_or_null(const Value* v) { return !v || v->is_null() || v->is_str(); }
inline bool obj_or_null(const Value* v) { return !v || v->is_null() || v->is_obj(); }
inline bool arr_or_null(const Value* v) { return !v || v->is_null() || v->is_arr(); }

// Usage (openai.go:2064-2083)
inline bool decode_usage(const Value& u, TokenUsage& tu) {
  int64_t p, c, t;
  if (!int_field(u.get("prompt_tokens"), p) || !int_field(u.get("completion_tokens"), c) || !int_field(u.get("total_tokens"), t)) return false;
  tu.input = (uint32_t)p; tu.output = (uint32_t)c; tu.total = (uint32_t)t; tu.mask |= TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
  const Value* ctd = u.get("completion_tokens_details");
  const Value* ptd = u.get("prompt_tokens_details");
  if (!obj_or_null(ctd) || !obj_or_null(ptd)) return false;
  if (ctd && ctd->is_obj()) {
    int64_t x, r = 0;
    for (const char* k : {"text_tokens", "accepted_prediction_tokens", "audio_tokens", "rejected_prediction_tokens"}) if (!int_field(ctd->get(k), x)) return false;
    if (!int_field(ctd->get("reasoning_tokens"), r)) return false;
    tu.reasoning = (uint32_t)r; tu.mask |= TokenUsage::REASONING;
  }
  if (ptd && ptd->is_obj()) {
    int64_t x, cch = 0, ccr = 0;
    for (const char* k : {"text_tokens", "audio_tokens"}) if (!int_field(ptd->get(k), x)) return false;
    if (!int_field(ptd->get("cached_tokens"), cch) || !int_field(ptd->get("cache_creation_input_tokens"), ccr)) return false;
    tu.cached = (uint32_t)cch; tu.cache_creation = (uint32_t)ccr; tu.mask |= TokenUsage::CACHED | TokenUsage::CACHE_CREATION;
  }
  return true;
}

// json.Unmarshal(line, &ChatCompletionResponseChunk{}) — true when it would succeed.
inline bool decode_chunk(std::string_view line, std::string& model, bool& has_usage, TokenUsage& tu) {
  Value v; std::string err;
  if (!oj::parse(line, v, err)) return false;
  model.clear(); has_usage = false;
  if (v.is_null()) return true;
  if (!v.is_obj()) return false;
  for (const char* k : {"id", "model", "service_tier", "system_fingerprint", "object", "obfuscation"}) if (!str_or_null(v.get(k))) return false;
  if (const Value* c = v.get("created")) {  // JSONUNIXTime.UnmarshalJSON, called for null too (openai.go:1789-1807)
    std::string raw(line.substr(c->b, c->e - c->b));
    size_t dot = raw.find('.'); if (dot != std::string::npos) raw.resize(dot);
    int64_t q; if (!oj::num_to_i64(raw, q)) return false;
  }
  const Value* ch = v.get("choices");
  if (!arr_or_null(ch)) return false;
  if (ch && ch->is_arr()) for (auto& c : ch->arr) {
    if (c.is_null()) continue;
    if (!c.is_obj()) return false;
    int64_t idx; if (!int_field(c.get("index"), idx)) return false;
    if (!str_or_null(c.get("finish_reason"))) return false;
    const Value* lp = c.get("logprobs"); if (!obj_or_null(lp)) return false;
    if (lp && lp->is_obj()) for (const char* k : {"content", "refusal"}) {  // []ChatCompletionTokenLogprob (openai.go:1323-1361)
      const Value* a = lp->get(k); if (!arr_or_null(a)) return false;
      if (a && a->is_arr()) for (auto& t : a->arr) {
        if (t.is_null()) continue;
        if (!t.is_obj()) return false;
        auto tok_ok = [&](const Value& x) -> bool {
          if (!str_or_null(x.get("token"))) return false;
          const Value* lpv = x.get("logprob"); if (lpv && !lpv->is_null() && !lpv->is_num()) return false;
          const Value* by = x.get("bytes"); if (!arr_or_null(by)) return false;
          if (by && by->is_arr()) for (auto& bb : by->arr) { int64_t q; if (!int_field(&bb, q)) return false; }
          return true; };
        if (!tok_ok(t)) return false;
        const Value* tl = t.get("top_logprobs"); if (!arr_or_null(tl)) return false;
        if (tl && tl->is_arr()) for (auto& u : tl->arr) { if (u.is_null()) continue; if (!u.is_obj() || !tok_ok(u)) return false; }
      }
    }
    const Value* d = c.get("delta"); if (!obj_or_null(d)) return false;
    if (d && d->is_obj()) {
      if (!str_or_null(d->get("content")) || !str_or_null(d->get("role"))) return false;
      const Value* tcs = d->get("tool_calls"); if (!arr_or_null(tcs)) return false;
      if (tcs && tcs->is_arr()) for (auto& tc : tcs->arr) {
        if (tc.is_null()) continue;
        if (!tc.is_obj()) return false;
        int64_t ti; if (!int_field(tc.get("index"), ti)) return false;
        if (!str_or_null(tc.get("id")) || !str_or_null(tc.get("type"))) return false;
        const Value* f = tc.get("function"); if (!obj_or_null(f)) return false;
        if (f && f->is_obj() && (!str_or_null(f->get("arguments")) || !str_or_null(f->get("name")))) return false;
      }
      const Value* an = d->get("annotations"); if (!arr_or_null(an)) return false;
      if (an && an->is_arr()) for (auto& a : an->arr) { if (!(a.is_null() || a.is_obj())) return false; if (a.is_obj()) { if (!str_or_null(a.get("type")) || !obj_or_null(a.get("url_citation"))) return false;
          const Value* uc = a.get("url_citation");
          if (uc && uc->is_obj()) { int64_t q; if (!int_field(uc->get("end_index"), q) || !int_field(uc->get("start_index"), q) || !str_or_null(uc->get("url")) || !str_or_null(uc->get("title"))) return false; } } }
      const Value* rc = d->get("reasoning_content"); if (!obj_or_null(rc)) return false;
      if (rc && rc->is_obj()) {
        if (!str_or_null(rc->get("text")) || !str_or_null(rc->get("signature"))) return false;
        if (const Value* red = rc->get("redactedContent"); red && !red->is_null()) { std::string tmp; if (!red->is_str() || !oj::b64dec(red->s, tmp)) return false; }
      }
    }
  }
  if (const Value* m = v.get("model"); m && m->is_str()) model = m->s;
  const Value* u = v.get("usage");
  if (!obj_or_null(u)) return false;
  if (u && u->is_obj()) { tu = TokenUsage{}; if (!decode_usage(*u, tu)) return false; has_usage = true; }
  return true;
}

struct SSEOpenAIState {  // the translator's per-stream fields (openai_openai.go:36-50)
  std::string buffered;
  std::string streaming_model;
};

// One ResponseBody(stream) call: append chunk, scan complete lines, return the latest usage of THIS call.
inline TokenUsage sse_openai_feed(SSEOpenAIState& st, std::string_view chunk) {
  st.buffered.append(chunk);
  TokenUsage out;
  size_t pos = 0;
  for (;;) {
    size_t nl = st.buffered.find('\n', pos);
    if (nl == std::string::npos) break;
    std::string_view line(st.buffered.data() + pos, nl - pos);
    pos = nl + 1;
    if (line.substr(0, 6) != "data: ") continue;
    std::string model; bool has_usage; TokenUsage tu;
    if (!decode_chunk(line.substr(6), model, has_usage, tu)) continue;
    if (!model.empty()) st.streaming_model = model;
    if (has_usage) {  // every Set… in :201-211 overwrites; flags accumulate within the call
      out.input = tu.input; out.output = tu.output; out.total = tu.total; out.mask |= TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
      if (tu.mask & TokenUsage::CACHED) { out.cached = tu.cached; out.cache_creation = tu.cache_creation; out.mask |= TokenUsage::CACHED | TokenUsage::CACHE_CREATION; }
      if (tu.mask & TokenUsage::REASONING) { out.reasoning = tu.reasoning; out.mask |= TokenUsage::REASONING; }
    }
  }
  st.buffered.erase(0, pos);
  return out;
}

// R1 (OpenAI): json.NewDecoder(body).Decode(&ChatCompletionResponse{}) then usage/model (openai_openai.go:146-174).
// Type checks follow internal/apischema/openai/openai.go:1269-1306 (response), :1365-1422 (choice, message), :1323-1361 (logprobs),
// :1424-1452 (annotations, audio), :1456-1461 (thinking blocks), :1840-1873 (reasoning_content union); the two genai-typed
// Gemini extras (safety_ratings, grounding_metadata) are only syntax-checked (their types live in google.golang.org/genai).
inline bool logprobs_ok(const Value* lp) {
  if (!obj_or_null(lp)) return false;
  if (lp && lp->is_obj()) for (const char* k : {"content", "refusal"}) {
    const Value* a = lp->get(k); if (!arr_or_null(a)) return false;
    if (a && a->is_arr()) for (auto& t : a->arr) {
      if (t.is_null()) continue;
      if (!t.is_obj()) return false;
      auto tok_ok = [&](const Value& x) -> bool {
        if (!str_or_null(x.get("token"))) return false;
        const Value* lpv = x.get("logprob"); if (lpv && !lpv->is_null() && !lpv->is_num()) return false;
        const Value* by = x.get("bytes"); if (!arr_or_null(by)) return false;
        if (by && by->is_arr()) for (auto& bb : by->arr) { int64_t q; if (!int_field(&bb, q)) return false; }
        return true; };
      if (!tok_ok(t)) return false;
      const Value* tl = t.get("top_logprobs"); if (!arr_or_null(tl)) return false;
      if (tl && tl->is_arr()) for (auto& u : tl->arr) { if (u.is_null()) continue; if (!u.is_obj() || !tok_ok(u)) return false; }
    }
  }
  return true;
}
inline bool annotations_ok(const Value* an) {
  if (!arr_or_null(an)) return false;
  if (an && an->is_arr()) for (auto& a : an->arr) {
    if (a.is_null()) continue;
    if (!a.is_obj()) return false;
    if (!str_or_null(a.get("type")) || !obj_or_null(a.get("url_citation"))) return false;
    const Value* uc = a.get("url_citation");
    if (uc && uc->is_obj()) { int64_t q; if (!int_field(uc->get("end_index"), q) || !int_field(uc->get("start_index"), q) || !str_or_null(uc->get("url")) || !str_or_null(uc->get("title"))) return false; }
  }
  return true;
}
inline bool response_openai(std::string_view body, const std::string& request_model, TokenUsage& tu, std::string& response_model) {
  Value v;
  oj::Parser ps(body.data(), body.size());  // json.Decoder reads ONE value; trailing bytes are not an error
  if (!ps.value(v)) return false;
  tu = TokenUsage{}; response_model = request_model;
  TokenUsage u2; u2.mask = TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
  if (v.is_obj()) {
    for (const char* k : {"id", "model", "service_tier", "system_fingerprint", "object", "obfuscation"}) if (!str_or_null(v.get(k))) return false;
    if (const Value* c = v.get("created")) {
      std::string raw(body.substr(c->b, c->e - c->b));
      size_t dot = raw.find('.'); if (dot != std::string::npos) raw.resize(dot);
      int64_t q; if (!oj::num_to_i64(raw, q)) return false;
    }
    const Value* ch = v.get("choices"); if (!arr_or_null(ch)) return false;
    if (ch && ch->is_arr()) for (auto& c : ch->arr) {
      if (c.is_null()) continue;
      if (!c.is_obj()) return false;
      int64_t idx; if (!int_field(c.get("index"), idx) || !str_or_null(c.get("finish_reason")) || !logprobs_ok(c.get("logprobs"))) return false;
      const Value* m = c.get("message"); if (!obj_or_null(m)) return false;
      if (m && m->is_obj()) {
        if (!str_or_null(m->get("content")) || !str_or_null(m->get("role")) || !annotations_ok(m->get("annotations"))) return false;
        const Value* tcs = m->get("tool_calls"); if (!arr_or_null(tcs)) return false;
        if (tcs && tcs->is_arr()) for (auto& tc : tcs->arr) {
          if (tc.is_null()) continue;
          if (!tc.is_obj() || !str_or_null(tc.get("id")) || !str_or_null(tc.get("type"))) return false;
          const Value* f = tc.get("function"); if (!obj_or_null(f)) return false;
          if (f && f->is_obj() && (!str_or_null(f->get("arguments")) || !str_or_null(f->get("name")))) return false;
          const Value* cc = tc.get("cache_control"); if (!obj_or_null(cc)) return false;
          if (cc && cc->is_obj() && (!str_or_null(cc->get("type")) || !str_or_null(cc->get("ttl")))) return false;
        }
        const Value* au = m->get("audio"); if (!obj_or_null(au)) return false;
        if (au && au->is_obj()) { int64_t q; if (!str_or_null(au->get("data")) || !int_field(au->get("expires_at"), q) || !str_or_null(au->get("id")) || !str_or_null(au->get("transcript"))) return false; }
        if (const Value* rc = m->get("reasoning_content"); rc && !rc->is_null() && !rc->is_str()) {
          if (!rc->is_obj()) return false;
          const Value* blk = rc->get("reasoningContent"); if (!obj_or_null(blk)) return false;
          if (blk && blk->is_obj()) {
            const Value* rt = blk->get("reasoningText"); if (!obj_or_null(rt)) return false;
            if (rt && rt->is_obj() && (!str_or_null(rt->get("text")) || !str_or_null(rt->get("signature")))) return false;
            if (const Value* red = blk->get("redactedContent"); red && !red->is_null()) { std::string tmp; if (!red->is_str() || !oj::b64dec(red->s, tmp)) return false; }
          }
        }
        const Value* tb = m->get("thinking_blocks"); if (!arr_or_null(tb)) return false;
        if (tb && tb->is_arr()) for (auto& t : tb->arr) { if (t.is_null()) continue; if (!t.is_obj()) return false; for (const char* k : {"type", "thinking", "signature", "data"}) if (!str_or_null(t.get(k))) return false; }
      }
    }
    const Value* u = v.get("usage");
    if (!obj_or_null(u)) return false;
    if (u && u->is_obj()) { u2 = TokenUsage{}; if (!decode_usage(*u, u2)) return false; }
    if (const Value* m = v.get("model"); m && m->is_str() && !m->s.empty()) response_model = m->s;
  } else if (!v.is_null()) return false;
  tu = u2;
  return true;
}

// evalCost's direct selectors (processor_impl.go:707-756; filterapi.LLMRequestCostType)
enum CostType : int { COST_INPUT = 0, COST_CACHED_INPUT = 1, COST_CACHE_CREATION_INPUT = 2, COST_OUTPUT = 3, COST_TOTAL = 4, COST_REASONING = 5, COST_CEL = 6 };
inline uint64_t eval_cost(int type, const TokenUsage& u) {
  switch (type) {
    case COST_INPUT: return u.input; case COST_CACHED_INPUT: return u.cached; case COST_CACHE_CREATION_INPUT: return u.cache_creation;
    case COST_OUTPUT: return u.output; case COST_TOTAL: return u.total; case COST_REASONING: return u.reasoning;
  }
  return 0;
}

// R1 (embeddings): json.NewDecoder(body).Decode(&openai.EmbeddingResponse{}) then usage / model
// (internal/translator/openai_embeddings.go:70-88; types internal/apischema/openai/openai.go:1683-1731,1772-1778).
// Only input and total tokens are set; the response model is resp.Model.
inline bool response_embeddings(std::string_view body, TokenUsage& tu, std::string& response_model) {
  Value v;
  oj::Parser ps(body.data(), body.size());
  if (!ps.value(v)) return false;
  tu = TokenUsage{}; response_model.clear();
  int64_t prompt = 0, total = 0;
  if (v.is_obj()) {
    for (const char* k : {"object", "model"}) if (!str_or_null(v.get(k))) return false;
    const Value* d = v.get("data"); if (!arr_or_null(d)) return false;
    if (d && d->is_arr()) for (auto& e : d->arr) {
      if (e.is_null()) continue;
      if (!e.is_obj()) return false;
      if (!str_or_null(e.get("object"))) return false;
      int64_t q; if (!int_field(e.get("index"), q)) return false;
      if (const Value* em = e.get("embedding")) {   // EmbeddingUnion: []float64 first, then string
        if (em->is_arr()) { for (auto& x : em->arr) if (!x.is_null() && !x.is_num()) return false; }
        else if (!em->is_null() && !em->is_str()) return false;
      }
    }
    const Value* u = v.get("usage"); if (!obj_or_null(u)) return false;
    if (u && u->is_obj()) { if (!int_field(u->get("prompt_tokens"), prompt) || !int_field(u->get("total_tokens"), total)) return false; }
    if (const Value* m = v.get("model"); m && m->is_str()) response_model = m->s;
  } else if (!v.is_null()) return false;
  tu.input = (uint32_t)prompt; tu.total = (uint32_t)total; tu.mask = TokenUsage::IN | TokenUsage::TOTAL;
  return true;
}

}  // namespace oracle
