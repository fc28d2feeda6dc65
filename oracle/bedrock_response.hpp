// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the buffered Bedrock Converse → OpenAI ChatCompletion response translation (Bedrock):
//   ResponseBody (non-stream)        internal/translator/openai_awsbedrock.go:734-824
//   bedrockToolUseToOpenAICalls      internal/translator/openai_awsbedrock.go:623-641
//   stop-reason mapping              internal/translator/openai_awsbedrock.go:601-621
//   awsbedrock.ConverseResponse      internal/apischema/awsbedrock/awsbedrock.go:178-182,264-279,330-366,370-432
//   openai.ChatCompletionResponse    internal/apischema/openai/openai.go:1269-1306,1365-1422,596-614,1836-1873,2064-2082
// Output key order = Go struct declaration order (the same encoder produces the streamed chunks whose bytes ARE pinned,
// tests/data-plane/testupstream_test.go:427-447); the reference compares buffered responses with JSONEq only
// (testupstream_test.go:1473-1482), so the byte layout of this row is "struct order by convention", the content is pinned
// by the goldens at :238-241 and :286-289.  `created` is time.Now() in the reference; here it is an input.
// Not restated (DECLINED): content blocks carrying document / image / toolResult / cachePoint members, and the inputs on
// which the reference dereferences a nil pointer (missing "output", null content element).
#pragma once
#include "bedrock_stream.hpp"
#include "translate.hpp"

namespace oracle {

// encoding/json-style re-marshal of a decoded `any` value: sorted keys, float64 numbers (translate.hpp enc_any)
// to_anthropic: the /v1/messages form of the same response (anthropicToAWSBedrockTranslator.ResponseBody, non-stream:
// internal/translator/anthropic_awsbedrock.go:429-510, stop reasons :720-735; anthropic.MessagesResponse field order
// internal/apischema/anthropic/anthropic.go:1413-1440,1450-1480,1603-1612) — every block in order (text, else tool_use, else thinking /
// redacted_thinking), id = the x-amzn-requestid header, model = the request model, usage as the four float64 counters.
// Pinned by the data-plane golden "aws-bedrock - /anthropic/v1/messages" (exact expResponseBody; the reference compares it with JSONEq).
inline Status bedrock_response(std::string_view body, const BedrockStreamCfg& cfg, std::string& out, TokenUsage& usage, bool to_anthropic = false) {
  out.clear(); usage = TokenUsage{};
  Value v;
  oj::Parser ps(body.data(), body.size());  // json.Decoder reads ONE value; trailing bytes are not an error
  if (!ps.value(v)) return INTERNAL;
  if (v.is_null()) return DECLINED;         // zero struct: Output == nil, the reference panics
  if (!v.is_obj()) return INTERNAL;
  // metrics
  if (const Value* m = v.get("metrics")) { if (!obj_or_null(m)) return INTERNAL; if (m->is_obj()) { int64_t x; if (!int_field(m->get("latencyMs"), x)) return INTERNAL; } }
  std::optional<std::string> stop; if (!opt_str(v.get("stopReason"), stop)) return INTERNAL;
  std::string tier; bool has_tier = false;
  if (const Value* sv = v.get("serviceTier")) { if (!obj_or_null(sv)) return INTERNAL; if (sv->is_obj()) { has_tier = true; if (!plain_str(sv->get("type"), tier)) return INTERNAL; } }
  bool has_usage = false; int64_t in_tok = 0, out_tok = 0, tot = 0; std::optional<int64_t> rd, wr;
  if (const Value* u = v.get("usage")) {
    if (!obj_or_null(u)) return INTERNAL;
    if (u->is_obj()) {
      has_usage = true;
      if (!int_field(u->get("inputTokens"), in_tok) || !int_field(u->get("outputTokens"), out_tok) || !int_field(u->get("totalTokens"), tot)) return INTERNAL;
      for (auto kv : {std::make_pair("cacheReadInputTokens", &rd), std::make_pair("cacheWriteInputTokens", &wr)}) {
        const Value* c = u->get(kv.first); if (c && !c->is_null()) { int64_t x; if (!int_field(c, x)) return INTERNAL; *kv.second = x; }
      }
    }
  }
  const Value* o = v.get("output");
  if (o && !obj_or_null(o)) return INTERNAL;
  // message
  std::string role; std::optional<std::string> content; std::string tool_calls, reasoning; bool has_reasoning = false;
  std::string an_blocks; bool an_any = false;
  bool decl = false;
  if (o && o->is_obj()) {
    const Value* m = o->get("message");
    if (m && !obj_or_null(m)) return INTERNAL;
    if (m && m->is_obj()) {
      if (!plain_str(m->get("role"), role)) return INTERNAL;
      const Value* c = m->get("content");
      if (c && !arr_or_null(c)) return INTERNAL;
      if (c && c->is_arr()) {
        for (const Value& blk : c->arr) {
          if (blk.is_null()) { decl = true; continue; }   // nil *ContentBlock: the reference dereferences it
          if (!blk.is_obj()) return INTERNAL;
          for (const char* k : {"document", "image", "toolResult", "cachePoint"}) if (const Value* x = blk.get(k); x && !x->is_null()) decl = true;
          std::optional<std::string> text; if (!opt_str(blk.get("text"), text)) return INTERNAL;
          const Value* tu = blk.get("toolUse"); if (!obj_or_null(tu)) return INTERNAL;
          const Value* rc = blk.get("reasoningContent"); if (!obj_or_null(rc)) return INTERNAL;
          std::string name, id, args = "null"; bool has_tu = false;
          if (tu && tu->is_obj()) {
            has_tu = true;
            if (!plain_str(tu->get("name"), name) || !plain_str(tu->get("toolUseId"), id)) return INTERNAL;
            const Value* in = tu->get("input"); if (!obj_or_null(in)) return INTERNAL;
            if (in && in->is_obj()) { args.clear(); oj::enc_any(args, *in); }
          }
          std::string rtext, rsig, red; bool has_rt = false, has_rc = false;
          if (rc && rc->is_obj()) {
            has_rc = true;
            const Value* rt = rc->get("reasoningText"); if (!obj_or_null(rt)) return INTERNAL;
            if (rt && rt->is_obj()) { has_rt = true; if (!plain_str(rt->get("text"), rtext) || !plain_str(rt->get("signature"), rsig)) return INTERNAL; }
            if (const Value* rd2 = rc->get("redactedContent"); rd2 && !rd2->is_null()) { if (!rd2->is_str() || !oj::b64dec(rd2->s, red)) return INTERNAL; }
          }
          if (to_anthropic) {
            std::string b;
            if (text) { b = "{\"type\":\"text\",\"text\":"; oj::enc_str(b, *text); b.push_back('}'); }
            else if (has_tu) { b = "{\"type\":\"tool_use\",\"id\":"; oj::enc_str(b, id); b += ",\"name\":"; oj::enc_str(b, name); b += ",\"input\":" + args + "}"; }
            else if (has_rc && has_rt) { b = "{\"type\":\"thinking\",\"thinking\":"; oj::enc_str(b, rtext); if (!rsig.empty()) { b += ",\"signature\":"; oj::enc_str(b, rsig); } b.push_back('}'); }
            else if (has_rc) { if (const Value* rd2 = rc->get("redactedContent"); rd2 && !rd2->is_null()) { b = "{\"type\":\"redacted_thinking\",\"data\":"; oj::enc_str(b, red); b.push_back('}'); } }
            if (!b.empty()) { if (an_any) an_blocks.push_back(','); an_any = true; an_blocks += b; }
          }
          if (has_tu) {
            if (!tool_calls.empty()) tool_calls.push_back(',');
            tool_calls += "{\"id\":"; oj::enc_str(tool_calls, id); tool_calls += ",\"function\":{\"arguments\":"; oj::enc_str(tool_calls, args);
            tool_calls += ",\"name\":"; oj::enc_str(tool_calls, name); tool_calls += "},\"type\":\"function\"}";
          } else if (text) { if (!content) content = *text; }
          else if (has_rc) {
            has_reasoning = true; reasoning = "{\"reasoningContent\":{"; bool f = true;
            if (has_rt) { reasoning += "\"reasoningText\":{\"text\":"; oj::enc_str(reasoning, rtext); if (!rsig.empty()) { reasoning += ",\"signature\":"; oj::enc_str(reasoning, rsig); } reasoning += "}"; f = false; }
            if (!red.empty()) { if (!f) reasoning.push_back(','); reasoning += "\"redactedContent\":\"" + oj::b64enc(red) + "\""; }
            reasoning += "}}";
          }
        }
      }
    }
  }
  if (!o || o->is_null()) return DECLINED;  // bedrockResp.Output.Message on a nil Output panics
  if (decl) return DECLINED;
  if (to_anthropic) {
    if (has_usage) usage = explicit_caching_usage(in_tok, out_tok, rd, wr);
    out = "{\"id\":"; oj::enc_str(out, cfg.response_id); out += ",\"type\":\"message\",\"role\":\"assistant\",\"content\":[" + an_blocks + "],\"model\":"; oj::enc_str(out, cfg.request_model);
    if (stop) {
      const std::string& s = *stop;
      out += ",\"stop_reason\":\""; out += (s == "max_tokens" || s == "stop_sequence" || s == "tool_use") ? s : std::string("end_turn"); out += "\"";
    }
    if (has_usage) {
      out += ",\"usage\":{\"cache_creation_input_tokens\":" + std::to_string((long long)(wr ? *wr : 0)) + ",\"cache_read_input_tokens\":" + std::to_string((long long)(rd ? *rd : 0));
      out += ",\"input_tokens\":" + std::to_string((long long)in_tok) + ",\"output_tokens\":" + std::to_string((long long)out_tok) + "}";
    }
    out += "}";
    return OK;
  }
  // usage
  std::string usage_json;
  if (has_usage) {
    usage = explicit_caching_usage(in_tok, out_tok, rd, wr);
    bool f = true;
    auto num = [&](const char* k, long long x) { if (x) { if (!f) usage_json.push_back(','); f = false; usage_json += std::string("\"") + k + "\":" + std::to_string(x); } };
    num("prompt_tokens", (long long)usage.input); num("completion_tokens", (long long)usage.output); num("total_tokens", (long long)usage.total);
    if (rd || wr) {
      if (!f) usage_json.push_back(','); f = false;
      usage_json += "\"prompt_tokens_details\":{"; bool g = true;
      if (rd && *rd) { usage_json += "\"cached_tokens\":" + std::to_string((long long)*rd); g = false; }
      if (wr && *wr) { if (!g) usage_json.push_back(','); usage_json += "\"cache_creation_input_tokens\":" + std::to_string((long long)*wr); }
      usage_json += "}";
    }
  }
  out = "{";
  if (!cfg.response_id.empty()) { out += "\"id\":"; oj::enc_str(out, cfg.response_id); out.push_back(','); }
  out += "\"choices\":[{\"finish_reason\":\""; out += bedrock_finish_reason(stop); out += "\",\"index\":0,\"message\":{";
  bool f = true;
  auto sep = [&] { if (!f) out.push_back(','); f = false; };
  if (content) { sep(); out += "\"content\":"; oj::enc_str(out, *content); }
  if (!role.empty()) { sep(); out += "\"role\":"; oj::enc_str(out, role); }
  if (!tool_calls.empty()) { sep(); out += "\"tool_calls\":[" + tool_calls + "]"; }
  if (has_reasoning) { sep(); out += "\"reasoning_content\":" + reasoning; }
  out += "}}],\"created\":" + std::to_string(cfg.created);
  if (!cfg.request_model.empty()) { out += ",\"model\":"; oj::enc_str(out, cfg.request_model); }
  if (has_tier && !tier.empty()) { out += ",\"service_tier\":"; oj::enc_str(out, tier); }
  out += ",\"object\":\"chat.completion\"";
  if (!usage_json.empty()) out += ",\"usage\":{" + usage_json + "}";
  out += "}";
  return OK;
}

}  // namespace oracle
