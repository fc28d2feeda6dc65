// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the Anthropic → OpenAI response side:
//   anthropicStreamParser.Process            internal/translator/anthropic_helper.go:826-919
//   parseAndHandleEvent                      internal/translator/anthropic_helper.go:921-942
//   handleAnthropicStreamEvent               internal/translator/anthropic_helper.go:944-1135
//   constructOpenAIChatCompletionChunk       internal/translator/anthropic_helper.go:1138-1162
//   messageToChatCompletion                  internal/translator/anthropic_helper.go:1165-1255
//   anthropicToOpenAIFinishReason            internal/translator/anthropic_helper.go:45-63
//   anthropicToolUseToOpenAICalls            internal/translator/anthropic_helper.go:750-769
//   ResponseBody (GCP Anthropic, buffered)   internal/translator/openai_gcpanthropic.go:237-278
//   chunk / response structs                 internal/apischema/openai/openai.go:1269-1306,1365-1422,1497-1565,2064-2083
//   usage arithmetic                         internal/metrics/metrics.go:226-254,292-307
// The event structs are anthropic-sdk-go v1.38.0 (not in tree).  Pinned by the exact streamed texts at
// tests/data-plane/testupstream_test.go:575,637.  Parity unpinned (the oracle answers DECLINED): a known field of the
// wrong JSON type (the SDK's apijson decoder is lenient in ways no in-tree test fixes), a tool_use block whose "input"
// is a non-empty object in content_block_start (Go map order), more than one `data:` line per event.
#pragma once
#include "bedrock_stream.hpp"

namespace oracle {

struct AnthropicStreamCfg { std::string request_model; int64_t created = 0; };
struct AnthropicStreamState {
  std::string buffered;
  std::string message_id;
  bool have_created = false;
  int64_t tool_index = -1;
  bool tool_active = false; int64_t active_index = -1;   // activeToolCalls holds at most the tool at tool_index
  TokenUsage usage;
  std::string stop_reason;
  bool sent_first = false;
  bool dead = false; Status dead_status = OK;
};

inline bool go_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
// bytes.TrimSpace (ASCII white space plus U+0085 and U+00A0 in their UTF-8 form)
inline std::string_view trim_space(std::string_view s) {
  for (;;) {
    if (!s.empty() && go_space((unsigned char)s.front())) { s.remove_prefix(1); continue; }
    if (s.size() >= 2 && (unsigned char)s[0] == 0xC2 && ((unsigned char)s[1] == 0x85 || (unsigned char)s[1] == 0xA0)) { s.remove_prefix(2); continue; }
    break;
  }
  for (;;) {
    if (!s.empty() && go_space((unsigned char)s.back())) { s.remove_suffix(1); continue; }
    if (s.size() >= 2 && (unsigned char)s[s.size() - 2] == 0xC2 && ((unsigned char)s.back() == 0x85 || (unsigned char)s.back() == 0xA0)) { s.remove_suffix(2); continue; }
    break;
  }
  return s;
}

inline const char* anthropic_finish_reason(const std::string& r) {  // anthropic_helper.go:45-63; nullptr = "received invalid stop reason"
  if (r == "end_turn" || r == "stop_sequence" || r == "pause_turn") return "stop";
  if (r == "max_tokens") return "length";
  if (r == "tool_use") return "tool_calls";
  if (r == "refusal") return "content_filter";
  return nullptr;
}

// {input_tokens, output_tokens, cache_read_input_tokens, cache_creation_input_tokens} of an anthropic Usage / MessageDeltaUsage
inline bool anthropic_usage_fields(const Value* u, int64_t& in, int64_t& out, int64_t& rd, int64_t& cr, bool& unpinned) {
  in = out = rd = cr = 0;
  if (!u || u->is_null()) return true;
  if (!u->is_obj()) { unpinned = true; return true; }
  if (!int_field(u->get("input_tokens"), in) || !int_field(u->get("output_tokens"), out) || !int_field(u->get("cache_read_input_tokens"), rd) ||
      !int_field(u->get("cache_creation_input_tokens"), cr)) unpinned = true;
  return true;
}

// serializeOpenAIChatCompletionChunk(constructOpenAIChatCompletionChunk(delta, finish))
struct AnDelta { std::optional<std::string> content; std::string tool_calls; /* serialized array body */ };
inline void anthropic_chunk(AnthropicStreamState& st, const AnthropicStreamCfg& cfg, const AnDelta& d, const char* finish, std::string& out) {
  bool role = false;
  if (!st.sent_first && (d.content || !d.tool_calls.empty())) { role = true; st.sent_first = true; }
  out += "data: {";
  if (!st.message_id.empty()) { out += "\"id\":"; oj::enc_str(out, st.message_id); out.push_back(','); }
  out += "\"choices\":[{\"index\":0,\"delta\":{";
  bool f = true;
  if (d.content) { out += "\"content\":"; oj::enc_str(out, *d.content); f = false; }
  if (role) { if (!f) out.push_back(','); out += "\"role\":\"assistant\""; f = false; }
  if (!d.tool_calls.empty()) { if (!f) out.push_back(','); out += "\"tool_calls\":[" + d.tool_calls + "]"; }
  out += "}";
  if (finish && *finish) { out += ",\"finish_reason\":\""; out += finish; out += "\""; }
  out += "}]";
  if (st.have_created) out += ",\"created\":" + std::to_string(cfg.created);
  if (!cfg.request_model.empty()) { out += ",\"model\":"; oj::enc_str(out, cfg.request_model); }
  out += ",\"object\":\"chat.completion.chunk\"}\n\n";
}

// One event block.  Returns OK, INTERNAL (the reference returns an error: the stream fails) or DECLINED (unpinned).
inline Status anthropic_event(AnthropicStreamState& st, const AnthropicStreamCfg& cfg, std::string_view block, std::string& out) {
  std::string_view etype; std::string data; int ndata = 0;
  size_t pos = 0;
  for (;;) {
    const size_t nl = block.find('\n', pos);
    std::string_view line = block.substr(pos, nl == std::string_view::npos ? std::string_view::npos : nl - pos);
    if (line.substr(0, 7) == "event: ") etype = trim_space(line.substr(7));
    else if (line.substr(0, 6) == "data: ") { const std::string_view d = trim_space(line.substr(6)); data.append(d); ndata++; }
    if (nl == std::string_view::npos) break;
    pos = nl + 1;
  }
  if (etype.empty() || data.empty()) return OK;
  if (ndata > 1) return DECLINED;
  const bool known = etype == "message_start" || etype == "content_block_start" || etype == "message_delta" || etype == "content_block_delta" ||
                     etype == "content_block_stop" || etype == "message_stop" || etype == "error";
  if (!known) return OK;   // ping and unknown event types are ignored before any decode
  Value v; std::string err;
  if (!oj::parse(data, v, err)) return INTERNAL;   // "unmarshal <event>: …"
  if (!v.is_obj()) return DECLINED;
  bool unpinned = false;
  auto str_of = [&](const Value* x, std::string& dst) { dst.clear(); if (!x || x->is_null()) return; if (!x->is_str()) { unpinned = true; return; } dst = x->s; };
  if (etype == "message_start") {
    const Value* m = v.get("message");
    std::string id; int64_t in = 0, o = 0, rd = 0, cr = 0;
    if (m && m->is_obj()) { str_of(m->get("id"), id); anthropic_usage_fields(m->get("usage"), in, o, rd, cr, unpinned); }
    else if (m && !m->is_null()) unpinned = true;
    if (unpinned) return DECLINED;
    st.message_id = id; st.have_created = true;
    const TokenUsage u = explicit_caching_usage(in, o, rd, cr);
    st.usage.input = u.input; st.usage.cached = u.cached; st.usage.cache_creation = u.cache_creation;
    st.usage.mask |= TokenUsage::IN | TokenUsage::CACHED | TokenUsage::CACHE_CREATION;
    st.tool_index = -1;
    return OK;
  }
  if (etype == "content_block_start") {
    const Value* cb = v.get("content_block");
    std::string type, id, name;
    const Value* input = nullptr;
    if (cb && cb->is_obj()) { str_of(cb->get("type"), type); str_of(cb->get("id"), id); str_of(cb->get("name"), name); input = cb->get("input"); }
    else if (cb && !cb->is_null()) unpinned = true;
    if (unpinned) return DECLINED;
    if (type == "tool_use" || type == "server_tool_use") {
      st.tool_index++;
      if (input && !input->is_null()) {
        if (!input->is_obj()) return INTERNAL;          // "unexpected tool use input type"
        if (!input->obj.empty()) return DECLINED;       // json.Marshal(map): layout left to the stock path
      }
      st.tool_active = true; st.active_index = st.tool_index;
      AnDelta d;
      d.tool_calls = "{\"index\":" + std::to_string(st.tool_index) + ",\"id\":"; oj::enc_str(d.tool_calls, id);
      d.tool_calls += ",\"function\":{\"arguments\":\"\",\"name\":"; oj::enc_str(d.tool_calls, name); d.tool_calls += "},\"type\":\"function\"}";
      anthropic_chunk(st, cfg, d, "", out);
      return OK;
    }
    if (type == "thinking") { AnDelta d; d.content = std::string(); anthropic_chunk(st, cfg, d, "", out); }
    return OK;
  }
  if (etype == "message_delta") {
    int64_t in = 0, o = 0, rd = 0, cr = 0;
    anthropic_usage_fields(v.get("usage"), in, o, rd, cr, unpinned);
    std::string stop;
    const Value* dl = v.get("delta");
    if (dl && dl->is_obj()) str_of(dl->get("stop_reason"), stop); else if (dl && !dl->is_null()) unpinned = true;
    if (unpinned) return DECLINED;
    const TokenUsage u = explicit_caching_usage(in, o, rd, cr);
    st.usage.output += u.output; st.usage.input += u.cached; st.usage.cached += u.cached;
    st.usage.input += u.cache_creation; st.usage.cache_creation += u.cache_creation;
    st.usage.mask |= TokenUsage::OUT | TokenUsage::IN | TokenUsage::CACHED | TokenUsage::CACHE_CREATION;
    if (!stop.empty()) st.stop_reason = stop;
    return OK;
  }
  if (etype == "content_block_delta") {
    const Value* dl = v.get("delta");
    std::string type, text, partial;
    if (dl && dl->is_obj()) { str_of(dl->get("type"), type); str_of(dl->get("text"), text); str_of(dl->get("partial_json"), partial); }
    else if (dl && !dl->is_null()) unpinned = true;
    if (unpinned) return DECLINED;
    if (type == "text_delta" || type == "thinking_delta") { AnDelta d; d.content = text; anthropic_chunk(st, cfg, d, "", out); return OK; }
    if (type == "input_json_delta") {
      if (!(st.tool_active && st.active_index == st.tool_index)) return INTERNAL;   // "received input_json_delta for unknown tool"
      AnDelta d;
      d.tool_calls = "{\"index\":" + std::to_string(st.tool_index) + ",\"id\":null,\"function\":{\"arguments\":"; oj::enc_str(d.tool_calls, partial);
      d.tool_calls += ",\"name\":\"\"}}";
      anthropic_chunk(st, cfg, d, "", out);
    }
    return OK;
  }
  if (etype == "content_block_stop") { if (st.tool_active && st.active_index == st.tool_index) st.tool_active = false; return OK; }
  if (etype == "message_stop") {
    if (st.stop_reason.empty()) st.stop_reason = "end_turn";
    const char* fr = anthropic_finish_reason(st.stop_reason);
    if (!fr) return INTERNAL;
    anthropic_chunk(st, cfg, AnDelta{}, fr, out);
    return OK;
  }
  return INTERNAL;  // "error" event: "anthropic stream error: …"
}

// One ResponseBody(stream) call (Process).  `usage` = the accumulated TokenUsage the call returns.
inline Status anthropic_stream_feed(AnthropicStreamState& st, const AnthropicStreamCfg& cfg, std::string_view chunk, bool eos, std::string& out, TokenUsage& usage) {
  if (st.dead) { usage = TokenUsage{}; return st.dead_status; }
  st.buffered.append(chunk);
  auto fail = [&](Status s) { st.dead = true; st.dead_status = s; usage = TokenUsage{}; return s; };
  for (;;) {
    const size_t cut = st.buffered.find("\n\n");
    if (cut == std::string::npos) break;
    const std::string block = st.buffered.substr(0, cut);
    const Status s = anthropic_event(st, cfg, block, out);
    if (s != OK) return fail(s);
    st.buffered.erase(0, cut + 2);
  }
  if (eos && !st.buffered.empty()) {
    const std::string block = st.buffered; st.buffered.clear();
    const Status s = anthropic_event(st, cfg, block, out);
    if (s != OK) return fail(s);
  }
  if (eos) {
    if (st.tool_active) return fail(DECLINED);   // a tool call still open at end of stream is replayed in the final chunk: stock path
    st.usage.total = st.usage.input + st.usage.output; st.usage.mask |= TokenUsage::TOTAL;
    const int in = (int)st.usage.input, o = (int)st.usage.output, tot = (int)st.usage.total, cached = (int)st.usage.cached, cc = (int)st.usage.cache_creation;
    if (in > 0 || o > 0) {
      out += "data: {";
      if (!st.message_id.empty()) { out += "\"id\":"; oj::enc_str(out, st.message_id); out.push_back(','); }
      out += "\"choices\":[]";
      if (st.have_created) out += ",\"created\":" + std::to_string(cfg.created);
      if (!cfg.request_model.empty()) { out += ",\"model\":"; oj::enc_str(out, cfg.request_model); }
      out += ",\"object\":\"chat.completion.chunk\",\"usage\":{";
      bool f = true;
      auto num = [&](const char* k, long long x) { if (x) { if (!f) out.push_back(','); f = false; out += std::string("\"") + k + "\":" + std::to_string(x); } };
      num("prompt_tokens", in); num("completion_tokens", o); num("total_tokens", tot);
      if (!f) out.push_back(',');
      out += "\"prompt_tokens_details\":{"; f = true;
      num("cached_tokens", cached); num("cache_creation_input_tokens", cc);
      out += "}}}\n\n";
    }
    out += "data: [DONE]\n\n";
  }
  usage = st.usage;
  return OK;
}

// ---- buffered: anthropic.Message → openai.ChatCompletionResponse (openai_gcpanthropic.go:246-278 + messageToChatCompletion)
inline Status anthropic_response(std::string_view body, const AnthropicStreamCfg& cfg, std::string& out, TokenUsage& usage, std::string& response_model) {
  out.clear(); usage = TokenUsage{}; response_model = cfg.request_model;
  Value v;
  oj::Parser ps(body.data(), body.size());   // json.Decoder reads one value
  if (!ps.value(v)) return INTERNAL;
  if (!v.is_obj()) return DECLINED;
  bool unpinned = false;
  auto str_of = [&](const Value* x, std::string& dst) { dst.clear(); if (!x || x->is_null()) return; if (!x->is_str()) { unpinned = true; return; } dst = x->s; };
  std::string id, model, stop, role;
  str_of(v.get("id"), id); str_of(v.get("model"), model); str_of(v.get("stop_reason"), stop); str_of(v.get("role"), role);
  int64_t in = 0, o = 0, rd = 0, cr = 0;
  anthropic_usage_fields(v.get("usage"), in, o, rd, cr, unpinned);
  std::optional<std::string> content; std::string tool_calls, reasoning;
  if (const Value* c = v.get("content"); c && !c->is_null()) {
    if (!c->is_arr()) unpinned = true;
    else for (const Value& b : c->arr) {
      if (!b.is_obj()) { unpinned = true; continue; }
      std::string type, text, bid, name, thinking, sig, data;
      str_of(b.get("type"), type); str_of(b.get("text"), text); str_of(b.get("id"), bid); str_of(b.get("name"), name);
      str_of(b.get("thinking"), thinking); str_of(b.get("signature"), sig); str_of(b.get("data"), data);
      if (type == "tool_use") {
        if (!bid.empty()) {
          std::string args = "null";
          if (const Value* inp = b.get("input"); inp && !inp->is_null()) { args.clear(); oj::enc_any(args, *inp); }   // json.Marshal(any): sorted keys
          if (!tool_calls.empty()) tool_calls.push_back(',');
          tool_calls += "{\"id\":"; oj::enc_str(tool_calls, bid); tool_calls += ",\"function\":{\"arguments\":"; oj::enc_str(tool_calls, args);
          tool_calls += ",\"name\":"; oj::enc_str(tool_calls, name); tool_calls += "},\"type\":\"function\"}";
        }
      } else if (type == "text") { if (!text.empty() && !content) content = text; }
      else if (type == "thinking") {
        if (!thinking.empty()) { reasoning = "{\"reasoningContent\":{\"reasoningText\":{\"text\":"; oj::enc_str(reasoning, thinking); if (!sig.empty()) { reasoning += ",\"signature\":"; oj::enc_str(reasoning, sig); } reasoning += "}}}"; }
      } else if (type == "redacted_thinking") {
        if (!data.empty()) reasoning = "{\"reasoningContent\":{\"redactedContent\":\"" + oj::b64enc(data) + "\"}}";
      }
    }
  }
  if (unpinned) return DECLINED;
  if (!model.empty()) response_model = model;
  usage = explicit_caching_usage(in, o, rd, cr);
  const char* fr = anthropic_finish_reason(stop);
  if (!fr) return INTERNAL;                       // "received invalid stop reason"
  const char* orole = role == "assistant" ? "assistant" : role == "user" ? "user" : nullptr;
  if (!orole) return INTERNAL;                    // "invalid anthropic role"
  out = "{";
  if (!id.empty()) { out += "\"id\":"; oj::enc_str(out, id); out.push_back(','); }
  out += "\"choices\":[{\"finish_reason\":\""; out += fr; out += "\",\"index\":0,\"message\":{";
  bool f = true;
  auto sep = [&] { if (!f) out.push_back(','); f = false; };
  if (content) { sep(); out += "\"content\":"; oj::enc_str(out, *content); }
  sep(); out += "\"role\":\""; out += orole; out += "\"";
  if (!tool_calls.empty()) { sep(); out += "\"tool_calls\":[" + tool_calls + "]"; }
  if (!reasoning.empty()) { sep(); out += "\"reasoning_content\":" + reasoning; }
  out += "}}],\"created\":" + std::to_string(cfg.created);
  if (!response_model.empty()) { out += ",\"model\":"; oj::enc_str(out, response_model); }
  out += ",\"object\":\"chat.completion\",\"usage\":{";
  f = true;
  auto num = [&](const char* k, long long x) { if (x) { if (!f) out.push_back(','); f = false; out += std::string("\"") + k + "\":" + std::to_string(x); } };
  num("prompt_tokens", (long long)(int)usage.input); num("completion_tokens", (long long)(int)usage.output); num("total_tokens", (long long)(int)usage.total);
  if (!f) out.push_back(',');
  out += "\"prompt_tokens_details\":{"; f = true;
  num("cached_tokens", (long long)(int)usage.cached); num("cache_creation_input_tokens", (long long)(int)usage.cache_creation);
  out += "}}}";
  return OK;
}

}  // namespace oracle
