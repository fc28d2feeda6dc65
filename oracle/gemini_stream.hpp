// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the GCP Vertex AI Gemini → OpenAI response side:
//   ResponseBody (buffered) / geminiResponseToOpenAIMessage   internal/translator/openai_gcpvertexai.go:139-198,582-600
//   handleStreamingResponse                                    internal/translator/openai_gcpvertexai.go:202-280
//   parseGCPStreamingChunks / detectSSEDelimiter               internal/translator/openai_gcpvertexai.go:43-54,283-337
//   geminiCandidatesToOpenAIStreamingChoices / convertGCPChunkToOpenAI   internal/translator/openai_gcpvertexai.go:390-482
//   geminiCandidatesToOpenAIChoices                            internal/translator/gemini_helper.go:741-828
//   geminiFinishReasonToOpenAI                                 internal/translator/gemini_helper.go:836-868
//   extractTextAndThoughtSummaryFromGeminiParts                internal/translator/gemini_helper.go:871-899
//   geminiUsageToOpenAIUsage                                   internal/translator/gemini_helper.go:942-958
// genai.GenerateContentResponse is google.golang.org/genai v1.55.0 (not in tree).  Pinned by the streamed text at
// tests/data-plane/testupstream_test.go:523 and (JSONEq) the buffered responses at :313,327,341.  Parity unpinned ⇒ DECLINED:
// functionCall parts (the tool-call id is a fresh uuid.New()), thoughtSignature, safetyRatings / groundingMetadata / logprobsResult,
// a candidate without content in a buffered response, wrong JSON types, createTime that is not RFC 3339.
#pragma once
#include "anthropic_stream.hpp"

namespace oracle {

struct GeminiCfg { std::string request_model; };
struct GeminiStreamState { std::string buffered; std::string delim; bool dead = false; Status dead_status = OK; };

// RFC 3339 → Unix seconds (time.Time.UnmarshalJSON + Unix()); false when malformed
inline bool rfc3339_unix(std::string_view s, int64_t& out) {
  auto dig = [&](size_t i) { return i < s.size() && s[i] >= '0' && s[i] <= '9'; };
  auto num = [&](size_t i, int n, int& v) { v = 0; for (int k = 0; k < n; k++) { if (!dig(i + k)) return false; v = v * 10 + (s[i + k] - '0'); } return true; };
  int Y, M, D, h, m, sec;
  if (s.size() < 20 || !num(0, 4, Y) || s[4] != '-' || !num(5, 2, M) || s[7] != '-' || !num(8, 2, D) || (s[10] != 'T' && s[10] != 't') || !num(11, 2, h) || s[13] != ':' || !num(14, 2, m) || s[16] != ':' || !num(17, 2, sec)) return false;
  size_t i = 19;
  if (i < s.size() && s[i] == '.') { i++; if (!dig(i)) return false; while (dig(i)) i++; }
  int off = 0;
  if (i < s.size() && (s[i] == 'Z' || s[i] == 'z')) i++;
  else { int oh, om; if (i + 6 > s.size() || (s[i] != '+' && s[i] != '-') || !num(i + 1, 2, oh) || s[i + 3] != ':' || !num(i + 4, 2, om)) return false; off = (oh * 60 + om) * 60 * (s[i] == '-' ? -1 : 1); i += 6; }
  if (i != s.size() || M < 1 || M > 12 || D < 1 || D > 31 || h > 23 || m > 59 || sec > 59) return false;
  int64_t y = Y - (M <= 2); const int64_t era = (y >= 0 ? y : y - 399) / 400; const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (M + (M > 2 ? -3 : 9)) + 2) / 5 + D - 1; const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  out = (era * 146097 + doe - 719468) * 86400 + h * 3600 + m * 60 + sec - off;
  return true;
}

inline const char* gemini_finish_reason(const std::string& r, bool tool_calls) {   // gemini_helper.go:836-868
  if (r == "STOP") return tool_calls ? "tool_calls" : "stop";
  if (r == "MAX_TOKENS") return "length";
  if (r == "SAFETY" || r == "BLOCKLIST" || r == "PROHIBITED_CONTENT" || r == "SPII" || r == "IMAGE_SAFETY" || r == "IMAGE_PROHIBITED_CONTENT") return "content_filter";
  if (r == "RECITATION" || r == "IMAGE_RECITATION") return "recitation";
  if (r == "MALFORMED_FUNCTION_CALL") return "malformed_function_call";
  if (r == "UNEXPECTED_TOOL_CALL") return "unexpected_tool_call";
  if (r == "LANGUAGE") return "language";
  if (r == "NO_IMAGE") return "no_image";
  if (r.empty()) return "";
  return "error";
}

struct GemCandidate { bool null = false, has_content = false; std::string text, thought, finish; };
struct GemResponse {
  std::string id, model_version; bool has_time = false; int64_t created = 0;
  std::vector<GemCandidate> cands;
  bool has_usage = false; int64_t prompt = 0, cand = 0, total = 0, cached = 0, thoughts = 0;
};
// 0 decoded, 1 not decodable (syntax / not an object), 2 unpinned
inline int gemini_decode(std::string_view js, GemResponse& g, bool buffered) {
  Value v; std::string err;
  if (buffered) { oj::Parser ps(js.data(), js.size()); if (!ps.value(v)) return 1; }
  else if (!oj::parse(js, v, err)) return 1;
  if (v.is_null()) return 0;
  if (!v.is_obj()) return 2;
  bool unp = false;
  auto str_of = [&](const Value* x, std::string& dst) { dst.clear(); if (!x || x->is_null()) return; if (!x->is_str()) { unp = true; return; } dst = x->s; };
  auto int_of = [&](const Value* x, int64_t& dst) { dst = 0; if (!x || x->is_null()) return; if (!int_field(x, dst) || dst < 0 || dst >= (1ll << 31)) unp = true; };
  str_of(v.get("responseId"), g.id); str_of(v.get("modelVersion"), g.model_version);
  if (const Value* t = v.get("createTime"); t && !t->is_null()) { if (!t->is_str() || !rfc3339_unix(t->s, g.created)) unp = true; else g.has_time = true; }
  if (const Value* u = v.get("usageMetadata"); u && !u->is_null()) {
    if (!u->is_obj()) unp = true;
    else { g.has_usage = true; int_of(u->get("promptTokenCount"), g.prompt); int_of(u->get("candidatesTokenCount"), g.cand); int_of(u->get("totalTokenCount"), g.total); int_of(u->get("cachedContentTokenCount"), g.cached); int_of(u->get("thoughtsTokenCount"), g.thoughts); }
  }
  if (const Value* cs = v.get("candidates"); cs && !cs->is_null()) {
    if (!cs->is_arr()) unp = true;
    else for (const Value& c : cs->arr) {
      GemCandidate gc;
      if (c.is_null()) { gc.null = true; g.cands.push_back(gc); continue; }
      if (!c.is_obj()) { unp = true; continue; }
      str_of(c.get("finishReason"), gc.finish);
      for (const char* k : {"safetyRatings", "groundingMetadata", "logprobsResult"}) if (const Value* x = c.get(k); x && !x->is_null() && buffered) unp = true;
      if (const Value* ct = c.get("content"); ct && !ct->is_null()) {
        if (!ct->is_obj()) { unp = true; continue; }
        gc.has_content = true;
        if (const Value* ps = ct->get("parts"); ps && !ps->is_null()) {
          if (!ps->is_arr()) unp = true;
          else for (const Value& p : ps->arr) {
            if (p.is_null()) continue;
            if (!p.is_obj()) { unp = true; continue; }
            if (const Value* fc = p.get("functionCall"); fc && !fc->is_null()) unp = true;
            if (const Value* ts = p.get("thoughtSignature"); ts && !ts->is_null()) unp = true;
            std::string text; str_of(p.get("text"), text);
            bool thought = false;
            if (const Value* th = p.get("thought"); th && !th->is_null()) { if (th->t == oj::T::True) thought = true; else if (th->t != oj::T::False) unp = true; }
            if (!text.empty()) (thought ? gc.thought : gc.text) += text;
          }
        }
      }
      g.cands.push_back(gc);
    }
  }
  return unp ? 2 : 0;
}

inline void gemini_usage_json(std::string& out, const GemResponse& g) {
  bool f = true;
  auto num = [&](const char* k, long long x) { if (x) { if (!f) out.push_back(','); f = false; out += std::string("\"") + k + "\":" + std::to_string(x); } };
  out += "{"; num("prompt_tokens", g.prompt); num("completion_tokens", g.cand + g.thoughts); num("total_tokens", g.total);
  if (!f) out.push_back(',');
  out += "\"completion_tokens_details\":{"; if (g.thoughts) out += "\"reasoning_tokens\":" + std::to_string(g.thoughts);
  out += "},\"prompt_tokens_details\":{"; if (g.cached) out += "\"cached_tokens\":" + std::to_string(g.cached);
  out += "}}";
}

inline void gemini_chunk_head(std::string& out, const GemResponse& g) { out += "data: {"; if (!g.id.empty()) { out += "\"id\":"; oj::enc_str(out, g.id); out.push_back(','); } }
inline void gemini_chunk_tail(std::string& out, const GemResponse& g, const GeminiCfg& cfg) {
  if (g.has_time) out += ",\"created\":" + std::to_string(g.created);
  if (!cfg.request_model.empty()) { out += ",\"model\":"; oj::enc_str(out, cfg.request_model); }
  out += ",\"object\":\"chat.completion.chunk\"";
}

// One ResponseBody(stream) call.
inline Status gemini_stream_feed(GeminiStreamState& st, const GeminiCfg& cfg, std::string_view chunk, bool eos, std::string& out, TokenUsage& usage) {
  usage = TokenUsage{};
  if (st.dead) return st.dead_status;
  std::string all = st.buffered; all.append(chunk);
  if (!all.empty()) {
    if (st.delim.empty()) { for (const char* d : {"\r\n\r\n", "\n\n", "\r\r"}) if (all.find(d) != std::string::npos) { st.delim = d; break; } }
    std::vector<std::string_view> parts;
    if (!st.delim.empty()) { size_t pos = 0; for (;;) { const size_t q = all.find(st.delim, pos); if (q == std::string::npos) { parts.emplace_back(all.data() + pos, all.size() - pos); break; } parts.emplace_back(all.data() + pos, q - pos); pos = q + st.delim.size(); } }
    else parts.emplace_back(all);
    std::string new_buf = st.buffered; bool touched = false;
    for (std::string_view part : parts) {
      part = trim_space(part);
      if (part.empty()) continue;
      if (part.substr(0, 6) == "data: ") part.remove_prefix(6);
      GemResponse g;
      const int rc = gemini_decode(part, g, false);
      if (rc == 2) { st.dead = true; st.dead_status = DECLINED; out.clear(); return DECLINED; }
      if (rc == 1) { new_buf.assign(part); touched = true; continue; }
      new_buf.clear(); touched = true;
      gemini_chunk_head(out, g); out += "\"choices\":[";
      bool cf = true; int idx = -1;
      for (const GemCandidate& c : g.cands) {
        idx++;
        if (c.null) continue;
        if (!cf) out.push_back(','); cf = false;
        out += "{\"index\":" + std::to_string(idx) + ",\"delta\":{";
        if (c.has_content) {
          if (!c.text.empty()) { out += "\"content\":"; oj::enc_str(out, c.text); out.push_back(','); }
          out += "\"role\":\"assistant\"";
          if (!c.thought.empty()) { out += ",\"reasoning_content\":{\"text\":"; oj::enc_str(out, c.thought); out += "}"; }
        }
        out += "}";
        const char* fr = gemini_finish_reason(c.finish, false);
        if (*fr) { out += ",\"finish_reason\":\""; out += fr; out += "\""; }
        out += "}";
      }
      out += "]"; gemini_chunk_tail(out, g, cfg); out += "}\n\n";
      if (g.has_usage && g.prompt > 0) {
        gemini_chunk_head(out, g); out += "\"choices\":[]"; gemini_chunk_tail(out, g, cfg); out += ",\"usage\":"; gemini_usage_json(out, g); out += "}\n\n";
        usage.input = (uint32_t)g.prompt; usage.output = (uint32_t)g.cand; usage.total = (uint32_t)g.total; usage.cached = (uint32_t)g.cached; usage.reasoning = (uint32_t)g.thoughts;
        usage.mask = TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL | TokenUsage::CACHED | TokenUsage::REASONING;
      }
    }
    if (touched) st.buffered = new_buf;
  }
  if (eos) out += "data: [DONE]\n";
  return OK;
}

// buffered response → ChatCompletionResponse
inline Status gemini_response(std::string_view body, const GeminiCfg& cfg, std::string& out, TokenUsage& usage, std::string& response_model) {
  out.clear(); usage = TokenUsage{}; response_model = cfg.request_model;
  GemResponse g;
  const int rc = gemini_decode(body, g, true);
  if (rc == 1) return INTERNAL;   // "error decoding GCP response"
  if (rc == 2) return DECLINED;
  for (const GemCandidate& c : g.cands) if (!c.null && !c.has_content) return DECLINED;   // zero message struct layout: stock path
  if (!g.model_version.empty()) response_model = g.model_version;
  out = "{";
  if (!g.id.empty()) { out += "\"id\":"; oj::enc_str(out, g.id); out.push_back(','); }
  out += "\"choices\":[";
  bool cf = true; int idx = -1;
  for (const GemCandidate& c : g.cands) {
    idx++;
    if (c.null) continue;
    if (!cf) out.push_back(','); cf = false;
    out += "{\"finish_reason\":\""; out += gemini_finish_reason(c.finish, false); out += "\",\"index\":" + std::to_string(idx) + ",\"message\":{";
    if (!c.text.empty()) { out += "\"content\":"; oj::enc_str(out, c.text); out.push_back(','); }
    out += "\"role\":\"assistant\"";
    if (!c.thought.empty()) { out += ",\"reasoning_content\":"; oj::enc_str(out, c.thought); }
    out += "}}";
  }
  out += "]";
  if (g.has_time) out += ",\"created\":" + std::to_string(g.created);
  if (!response_model.empty()) { out += ",\"model\":"; oj::enc_str(out, response_model); }
  out += ",\"object\":\"chat.completion\"";
  if (g.has_usage) { out += ",\"usage\":"; gemini_usage_json(out, g); }
  out += "}";
  usage.input = (uint32_t)g.prompt; usage.output = (uint32_t)(g.cand + g.thoughts); usage.total = (uint32_t)g.total;
  usage.mask = TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
  if (g.has_usage) { usage.cached = (uint32_t)g.cached; usage.reasoning = (uint32_t)g.thoughts; usage.mask |= TokenUsage::CACHED | TokenUsage::REASONING; }
  return OK;
}

}  // namespace oracle
