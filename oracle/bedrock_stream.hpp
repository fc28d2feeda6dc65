// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the streamed Bedrock ConverseStream → OpenAI SSE back-translation:
//   ResponseBody (stream)            internal/translator/openai_awsbedrock.go:695-732
//   extractAmazonEventStreamEvents   internal/translator/openai_awsbedrock.go:829-852
//   convertEvent                     internal/translator/openai_awsbedrock.go:858-1006
//   stop-reason mapping              internal/translator/openai_awsbedrock.go:601-621
//   event / chunk structs            internal/apischema/awsbedrock/awsbedrock.go:423-503, internal/apischema/openai/openai.go:1497-1565,2064-2083
//   usage arithmetic                 internal/metrics/metrics.go:292-307
// The frame codec is github.com/aws/aws-sdk-go-v2/aws/protocol/eventstream v1.7.10 (not in tree); its published wire format
// is restated: [total_len u32 BE][headers_len u32 BE][prelude CRC32][headers][payload][message CRC32], IEEE CRC-32,
// header = [name_len u8][name][type u8][value], type 7 = string with u16 BE length.  Pinned by the real capture at
// internal/translator/openai_awsbedrock_test.go:1859 with the exact expected output at :1338-1390.
#pragma once
#include "stream.hpp"

namespace oracle {

inline uint32_t crc32_ieee(const uint8_t* p, size_t n, uint32_t crc = 0) {
  static uint32_t tab[256]; static bool init = false;
  if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; tab[i] = c; } init = true; }
  crc = ~crc;
  for (size_t i = 0; i < n; i++) crc = tab[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return ~crc;
}
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct BedrockStreamCfg { std::string request_model, response_id; int64_t created = 0; };
struct BedrockStreamState {
  std::string buffered;
  std::string role;
  int64_t tool_index = 0;
  bool active_tool = false;
};

// value sizes of eventstream header types 0..9 (-1 = u16-length prefixed)
inline int header_value_size(int type) { static const int sz[10] = {0, 0, 1, 2, 4, 8, -1, -1, 8, 16}; return type >= 0 && type <= 9 ? sz[type] : -2; }

struct ConverseEvent {  // awsbedrock.ConverseStreamEvent after decode
  std::string event_type;
  bool has_delta = false; std::optional<std::string> delta_text; bool has_delta_tool = false; std::string delta_tool_input;
  bool has_delta_reasoning = false; std::string r_text, r_sig, r_redacted; bool r_has_redacted = false;
  std::optional<std::string> role, stop_reason;
  bool has_usage = false; int64_t in_tok = 0, out_tok = 0, total_tok = 0; std::optional<int64_t> cache_read, cache_write;
  bool has_start = false; bool has_start_tool = false; std::string start_name, start_id;
  bool has_service_tier = false; std::string service_tier;
};

inline bool opt_str(const Value* v, std::optional<std::string>& out) { if (!v || v->is_null()) return true; if (!v->is_str()) return false; out = v->s; return true; }
inline bool plain_str(const Value* v, std::string& out) { if (!v || v->is_null()) return true; if (!v->is_str()) return false; out = v->s; return true; }

inline bool decode_converse_event(std::string_view payload, ConverseEvent& e) {
  Value v; std::string err;
  if (!oj::parse(payload, v, err)) return false;
  if (v.is_null()) return true;
  if (!v.is_obj()) return false;
  { std::optional<std::string> t; if (!opt_str(v.get("eventType"), t)) return false; if (t) e.event_type = *t; }
  { int64_t x; if (!int_field(v.get("contentBlockIndex"), x)) return false; }
  if (const Value* d = v.get("delta")) {
    if (!obj_or_null(d)) return false;
    if (d->is_obj()) {
      e.has_delta = true;
      if (!opt_str(d->get("text"), e.delta_text)) return false;
      const Value* tu = d->get("toolUse"); if (!obj_or_null(tu)) return false;
      if (tu && tu->is_obj()) { e.has_delta_tool = true; if (!plain_str(tu->get("input"), e.delta_tool_input)) return false; }
      const Value* rc = d->get("reasoningContent"); if (!obj_or_null(rc)) return false;
      if (rc && rc->is_obj()) {
        e.has_delta_reasoning = true;
        if (!plain_str(rc->get("text"), e.r_text) || !plain_str(rc->get("signature"), e.r_sig)) return false;
        if (const Value* red = rc->get("redactedContent"); red && !red->is_null()) { if (!red->is_str() || !oj::b64dec(red->s, e.r_redacted)) return false; e.r_has_redacted = true; }
      }
    }
  }
  if (!opt_str(v.get("role"), e.role) || !opt_str(v.get("stopReason"), e.stop_reason)) return false;
  if (const Value* u = v.get("usage")) {
    if (!obj_or_null(u)) return false;
    if (u->is_obj()) {
      e.has_usage = true;
      if (!int_field(u->get("inputTokens"), e.in_tok) || !int_field(u->get("outputTokens"), e.out_tok) || !int_field(u->get("totalTokens"), e.total_tok)) return false;
      for (auto kv : {std::make_pair("cacheReadInputTokens", &e.cache_read), std::make_pair("cacheWriteInputTokens", &e.cache_write)}) {
        const Value* c = u->get(kv.first); if (c && !c->is_null()) { int64_t x; if (!int_field(c, x)) return false; *kv.second = x; }
      }
    }
  }
  if (const Value* st = v.get("start")) {
    if (!obj_or_null(st)) return false;
    if (st->is_obj()) {
      e.has_start = true;
      const Value* tu = st->get("toolUse"); if (!obj_or_null(tu)) return false;
      if (tu && tu->is_obj()) { e.has_start_tool = true; if (!plain_str(tu->get("name"), e.start_name) || !plain_str(tu->get("toolUseId"), e.start_id)) return false; }
    }
  }
  if (const Value* sv = v.get("serviceTier")) {
    if (!obj_or_null(sv)) return false;
    if (sv->is_obj()) { e.has_service_tier = true; if (!plain_str(sv->get("type"), e.service_tier)) return false; }
  }
  return true;
}

inline const char* bedrock_finish_reason(const std::optional<std::string>& r) {  // openai_awsbedrock.go:601-621
  if (!r) return "stop";
  if (*r == "stop_sequence" || *r == "end_turn") return "stop";
  if (*r == "max_tokens") return "length";
  if (*r == "content_filtered") return "content_filter";
  if (*r == "tool_use") return "tool_calls";
  return "stop";
}

inline TokenUsage explicit_caching_usage(int64_t in, int64_t out, const std::optional<int64_t>& read, const std::optional<int64_t>& write) {  // metrics.go:292-307
  TokenUsage u; int64_t tot_in = in;
  if (write) { tot_in += *write; u.cache_creation = (uint32_t)*write; u.mask |= TokenUsage::CACHE_CREATION; }
  if (read) { tot_in += *read; u.cached = (uint32_t)*read; u.mask |= TokenUsage::CACHED; }
  u.input = (uint32_t)tot_in; u.output = (uint32_t)out; u.total = (uint32_t)(tot_in + out);
  u.mask |= TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
  return u;
}

// serializeOpenAIChatCompletionChunk of the chunk convertEvent builds; false = "no chunk for this event"
inline bool convert_event(BedrockStreamState& st, const BedrockStreamCfg& cfg, const ConverseEvent& e, std::string& out) {
  std::string choices, service_tier, usage;
  auto delta_open = [&](std::string& c) { c += "{\"index\":0,\"delta\":{"; };
  const std::string& t = e.event_type;
  if (t == "metadata") {
    if (e.has_service_tier) service_tier = e.service_tier;
    if (!e.has_usage) return false;
    TokenUsage u = explicit_caching_usage(e.in_tok, e.out_tok, e.cache_read, e.cache_write);
    usage = "{"; bool f = true;
    auto num = [&](const char* k, long long v) { if (v) { if (!f) usage.push_back(','); f = false; usage += std::string("\"") + k + "\":" + std::to_string(v); } };
    num("prompt_tokens", (long long)(int)u.input); num("completion_tokens", (long long)(int)u.output); num("total_tokens", (long long)(int)u.total);
    if (e.cache_read || e.cache_write) {
      if (!f) usage.push_back(','); f = false;
      usage += "\"prompt_tokens_details\":{"; bool g = true;
      if (e.cache_read && (int)*e.cache_read) { usage += "\"cached_tokens\":" + std::to_string((long long)(int)*e.cache_read); g = false; }
      if (e.cache_write && (int)*e.cache_write) { if (!g) usage.push_back(','); usage += "\"cache_creation_input_tokens\":" + std::to_string((long long)(int)*e.cache_write); }
      usage += "}";
    }
    usage += "}";
  } else if (t == "messageStart") {
    if (!e.role) return false;
    delta_open(choices); choices += "\"content\":\"\"";
    if (!e.role->empty()) { choices += ",\"role\":"; oj::enc_str(choices, *e.role); }
    choices += "}}";
    st.role = *e.role;
  } else if (t == "contentBlockDelta") {
    if (!e.has_delta) return false;
    auto role_member = [&](bool lead_comma) { if (!st.role.empty()) { if (lead_comma) choices.push_back(','); choices += "\"role\":"; oj::enc_str(choices, st.role); return true; } return false; };
    if (e.delta_text) {
      delta_open(choices); choices += "\"content\":"; oj::enc_str(choices, *e.delta_text); role_member(true); choices += "}}";
    } else if (e.has_delta_tool) {
      delta_open(choices); bool r = role_member(false); if (r) choices.push_back(',');
      choices += "\"tool_calls\":[{\"index\":" + std::to_string(st.tool_index) + ",\"id\":null,\"function\":{\"arguments\":"; oj::enc_str(choices, e.delta_tool_input);
      choices += ",\"name\":\"\"},\"type\":\"function\"}]}}";
    } else if (e.has_delta_reasoning) {
      delta_open(choices); bool r = role_member(false); if (r) choices.push_back(',');
      choices += "\"reasoning_content\":{"; bool f = true;
      if (!e.r_text.empty()) { choices += "\"text\":"; oj::enc_str(choices, e.r_text); f = false; }
      if (!e.r_sig.empty()) { if (!f) choices.push_back(','); choices += "\"signature\":"; oj::enc_str(choices, e.r_sig); f = false; }
      if (e.r_has_redacted && !e.r_redacted.empty()) { if (!f) choices.push_back(','); choices += "\"redactedContent\":\"" + oj::b64enc(e.r_redacted) + "\""; }
      choices += "}}}";
    }
  } else if (t == "contentBlockStart") {
    if (!e.has_start) return false;
    if (e.has_start_tool) {
      st.active_tool = true;
      delta_open(choices); if (!st.role.empty()) { choices += "\"role\":"; oj::enc_str(choices, st.role); choices.push_back(','); }
      choices += "\"tool_calls\":[{\"index\":" + std::to_string(st.tool_index) + ",\"id\":"; oj::enc_str(choices, e.start_id);
      choices += ",\"function\":{\"arguments\":\"\",\"name\":"; oj::enc_str(choices, e.start_name); choices += "},\"type\":\"function\"}]}}";
    }
  } else if (t == "messageStop") {
    if (!e.stop_reason) return false;
    delta_open(choices); choices += "\"content\":\"\"";
    if (!st.role.empty()) { choices += ",\"role\":"; oj::enc_str(choices, st.role); }
    choices += "},\"finish_reason\":\""; choices += bedrock_finish_reason(e.stop_reason); choices += "\"}";
  } else if (t == "contentBlockStop") {
    if (st.active_tool) { st.tool_index++; st.active_tool = false; }
    return false;
  } else return false;
  out += "data: {";
  bool f = true;
  if (!cfg.response_id.empty()) { out += "\"id\":"; oj::enc_str(out, cfg.response_id); f = false; }
  if (!f) out.push_back(','); out += "\"choices\":[" + choices + "]";
  out += ",\"created\":" + std::to_string(cfg.created);
  if (!cfg.request_model.empty()) { out += ",\"model\":"; oj::enc_str(out, cfg.request_model); }
  if (!service_tier.empty()) { out += ",\"service_tier\":"; oj::enc_str(out, service_tier); }
  out += ",\"object\":\"chat.completion.chunk\"";
  if (!usage.empty()) out += ",\"usage\":" + usage;
  out += "}\n\n";
  return true;
}

// One ResponseBody(stream) call.  `usage` = tokenUsage returned by this call.
inline void bedrock_stream_feed(BedrockStreamState& st, const BedrockStreamCfg& cfg, std::string_view chunk, bool eos, std::string& out, TokenUsage& usage) {
  st.buffered.append(chunk);
  usage = TokenUsage{};
  const uint8_t* p = (const uint8_t*)st.buffered.data(); size_t n = st.buffered.size(), off = 0;
  for (;;) {
    if (n - off < 12) break;
    const uint32_t total = be32(p + off), hlen = be32(p + off + 4), pcrc = be32(p + off + 8);
    if (crc32_ieee(p + off, 8) != pcrc) break;                     // decoder error: this frame blocks the stream
    if (hlen > 128u * 1024 || total < 16u || hlen > total - 16u || total - hlen - 16u > 16u * 1024 * 1024) break;
    if (n - off < total) break;                                    // incomplete frame: wait for more bytes
    if (crc32_ieee(p + off, total - 4) != be32(p + off + total - 4)) break;
    // headers
    std::string etype; bool bad_headers = false;
    size_t h = off + 12, hend = h + hlen;
    while (h < hend) {
      const size_t nl = p[h]; h++;
      if (h + nl + 1 > hend) { bad_headers = true; break; }
      std::string_view name((const char*)p + h, nl); h += nl;
      const int type = p[h]; h++;
      int vs = header_value_size(type);
      if (vs == -2) { bad_headers = true; break; }
      if (vs == -1) { if (h + 2 > hend) { bad_headers = true; break; } vs = (p[h] << 8) | p[h + 1]; h += 2; }
      if (h + vs > hend) { bad_headers = true; break; }
      if (name == ":event-type" && type == 7) etype.assign((const char*)p + h, vs);
      h += vs;
    }
    if (bad_headers) break;
    ConverseEvent e; e.event_type = etype;
    const std::string_view payload((const char*)p + off + 12 + hlen, total - hlen - 16);
    if (decode_converse_event(payload, e)) {
      if (e.has_usage) usage = explicit_caching_usage(e.in_tok, e.out_tok, e.cache_read, e.cache_write);
      convert_event(st, cfg, e, out);
    }
    off += total;
  }
  st.buffered.erase(0, off);
  if (eos) out += "data: [DONE]\n";
}

}  // namespace oracle
