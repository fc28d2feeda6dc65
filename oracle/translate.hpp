// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the request-side translators:
//   T1 OpenAI→OpenAI RequestBody        internal/translator/openai_openai.go:55-84
//   P2 forced stream_options injection  internal/endpointspec/endpointspec.go:107-123
//   T2 OpenAI→AWS Bedrock Converse      internal/translator/openai_awsbedrock.go:57-226,229-585
//      output layout                    internal/apischema/awsbedrock/awsbedrock.go:41-70,126-200,310-367,536-604
//   sjson set semantics (module not in tree; published algorithm of tidwall/sjson v1.2.x `set`
//   / `appendRawPaths`, pinned by tests/data-plane/testupstream_test.go:383,1617,1635,1653,1671)
#pragma once
#include "chat.hpp"

namespace oracle {

// net/url.PathEscape (encodePathSegment): unreserved + "$&+:=@" stay, everything else %XX (upper hex).
inline std::string path_escape(std::string_view s) {
  static const char* hx = "0123456789ABCDEF";
  std::string o;
  for (unsigned char c : s) {
    bool keep = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '-' || c == '_' || c == '.' || c == '~' ||
                c == '$' || c == '&' || c == '+' || c == ':' || c == '=' || c == '@';
    if (keep) o.push_back((char)c); else { o.push_back('%'); o.push_back(hx[c >> 4]); o.push_back(hx[c & 15]); }
  }
  return o;
}

// sjson's string literal writer (appendStringify → mustMarshalString): plain quote when the
// string has only bytes in [' ', 0x7f] without '"' or '\\'; otherwise encoding/json.Marshal,
// which HTML-escapes <,>,& and U+2028/9 and writes � for invalid UTF-8.  Only the plain
// branch and the common escapes are restated; the rest is "parity unpinned".
inline void sjson_stringify(std::string& o, std::string_view s) {
  bool plain = true;
  for (unsigned char c : s) if (c < ' ' || c > 0x7f || c == '"' || c == '\\') { plain = false; break; }
  if (plain) { o.push_back('"'); o.append(s); o.push_back('"'); return; }
  static const char* hexd = "0123456789abcdef";
  o.push_back('"');
  for (size_t i = 0; i < s.size(); i++) {
    unsigned char c = s[i];
    if (c == '"') o += "\\\""; else if (c == '\\') o += "\\\\"; else if (c == '\n') o += "\\n"; else if (c == '\r') o += "\\r"; else if (c == '\t') o += "\\t";
    else if (c == '\b') o += "\\b"; else if (c == '\f') o += "\\f";
    else if (c < 0x20 || c == '<' || c == '>' || c == '&') { o += "\\u00"; o.push_back(hexd[c >> 4]); o.push_back(hexd[c & 15]); }
    else if (c == 0xE2 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x80 && ((unsigned char)s[i + 2] == 0xA8 || (unsigned char)s[i + 2] == 0xA9)) {
      o += ((unsigned char)s[i + 2] == 0xA8) ? "\\u2028" : "\\u2029"; i += 2;
    } else o.push_back((char)c);
  }
  o.push_back('"');
}

// sjson.SetBytesOptions(body, "<k1>[.<k2>]", raw, Optimistic) on a body that already parsed.
// `root` is the DOM of `body`.  Restates sjson `set`: existing path ⇒ splice the raw value;
// missing ⇒ appendRawPaths (append before the closing '}' of the deepest existing object;
// text before the first '{' and after the last '}' of a rebuilt object is dropped).
inline std::string sjson_set_raw(std::string_view body, const Value& root, std::string_view k1, std::string_view k2, std::string_view raw) {
  auto splice = [&](uint32_t b, uint32_t e, std::string_view with) { std::string o(body.substr(0, b)); o += with; o += body.substr(e); return o; };
  auto append_member = [&](const Value& objv, std::string_view member) {
    // objv is an object at [b,e): keep bytes up to its last '}', add ",member}" (no comma if empty)
    std::string o(body.substr(0, objv.b));
    std::string_view rawobj = body.substr(objv.b, objv.e - objv.b);
    size_t end = rawobj.size() - 1;  // objv.e-1 is '}'
    o += rawobj.substr(0, end);
    if (!objv.obj.empty()) o.push_back(',');
    o += member; o.push_back('}');
    o += body.substr(objv.e);
    return o;
  };
  if (!root.is_obj()) {  // gjson finds nothing; sjson rebuilds from "{}"
    std::string o = "{\""; o += k1; o += "\":";
    if (!k2.empty()) { o += "{\""; o += k2; o += "\":"; o += raw; o += "}"; } else o += raw;
    o += "}"; return o;
  }
  const Value* v1 = root.get_first(k1);
  if (k2.empty()) {
    if (v1) return splice(v1->b, v1->e, raw);
    std::string m = "\""; m += k1; m += "\":"; m += raw;
    // root rebuild drops bytes outside the outermost braces
    std::string o = append_member(root, m);
    return o.substr(root.b, o.size() - root.b - (body.size() - root.e));
  }
  if (v1) {
    if (v1->is_obj()) {
      if (const Value* v2 = v1->get_first(k2)) return splice(v2->b, v2->e, raw);
      std::string m = "\""; m += k2; m += "\":"; m += raw;
      return append_member(*v1, m);
    }
    // non-object (null, number, …): replaced by a fresh object
    std::string m = "{\""; m += k2; m += "\":"; m += raw; m += "}";
    return splice(v1->b, v1->e, m);
  }
  std::string m = "\""; m += k1; m += "\":{\""; m += k2; m += "\":"; m += raw; m += "}";
  std::string o = append_member(root, m);
  return o.substr(root.b, o.size() - root.b - (body.size() - root.e));
}

struct Header { std::string name, value; };
struct TranslateResult {
  Error err;
  int body_kind = UNCHANGED;
  std::string body;
  std::vector<Header> headers;
  std::string model;     // originalModel returned by ParseBody
  std::string request_model;  // after override
  bool stream = false;
  std::string mutated_body;  // ParseBody's mutatedBody (P2), empty ⇒ nil
  bool has_mutated = false;
};

// ---------------------------------------------------------------- T2: Bedrock Converse
namespace bedrock {

inline void cache_point(std::string& o) { o += "{\"cachePoint\":{\"type\":\"default\"}}"; }

// translator/util.go:33-50 parseDataURI + regexp `\Adata:(.+?)?(;base64)?,`
inline bool parse_data_uri(const std::string& uri, std::string& ctype, std::string& bin) {
  if (uri.rfind("data:", 0) != 0) return false;
  // Leftmost-first regexp semantics: group 1 is lazy, group 2 optional greedy, then ','.
  // The first position p ≥ 5 such that the text from p is ",..." or ";base64,...": with lazy
  // (.+?)? the engine prefers the empty/shortest group 1, and prefers matching (;base64) over skipping it.
  size_t n = uri.size();
  for (size_t p = 5; p <= n; p++) {
    // group1 = uri[5:p]
    if (uri.compare(p, 8, ";base64,") == 0) { ctype = uri.substr(5, p - 5); std::string rest = uri.substr(p + 8); return oj::b64dec(rest, bin); }
    if (p < n && uri[p] == ',') { ctype = uri.substr(5, p - 5); std::string rest = uri.substr(p + 1); return oj::b64dec(rest, bin); }
    if (p < n && uri[p] == '\n') return false;  // '.' does not match newline
  }
  return false;
}

inline Error user_content(const Message& m, std::string& o) {  // openai_awsbedrock.go:229-296
  if (m.ck == Message::String) { o += "[{\"text\":"; oj::enc_str(o, m.content_str); o += "}]"; return {}; }
  if (m.ck == Message::UserParts) {
    o.push_back('['); bool first = true;
    auto sep = [&] { if (!first) o.push_back(','); first = false; };
    for (auto& p : m.user_parts) {
      if (p.k == UserPart::Text) {
        sep(); o += "{\"text\":"; oj::enc_str(o, p.text.text); o.push_back('}');
        if (p.text.cache.ephemeral) { sep(); cache_point(o); }
      } else if (p.k == UserPart::ImageURL) {
        std::string ct, bin;
        if (!parse_data_uri(p.image_url, ct, bin)) return invalid("invalid image data URI");
        const char* fmt = ct == "image/png" ? "png" : ct == "image/jpeg" ? "jpeg" : ct == "image/gif" ? "gif" : ct == "image/webp" ? "webp" : nullptr;
        if (!fmt) return invalid("unsupported image format " + ct);
        sep(); o += "{\"image\":{\"format\":\""; o += fmt; o += "\",\"source\":{\"bytes\":\""; o += oj::b64enc(bin); o += "\"}}}";
        if (p.cache.ephemeral) { sep(); cache_point(o); }
      }
    }
    o.push_back(']'); return {};
  }
  return invalid("unexpected content type for user message");
}

inline Error assistant_content(const Message& m, std::string& o) {  // openai_awsbedrock.go:309-419
  std::vector<AsstPart> parts;
  if (m.ck == Message::String) { if (!m.content_str.empty()) { AsstPart p; p.type = "text"; p.text = m.content_str; parts.push_back(p); } }
  else if (m.ck == Message::AsstParts || m.ck == Message::AsstSingle) parts = m.asst_parts;
  o.push_back('['); bool first = true;
  auto sep = [&] { if (!first) o.push_back(','); first = false; };
  for (auto& p : parts) {
    if (p.type == "text") { if (p.text) { sep(); o += "{\"text\":"; oj::enc_str(o, *p.text); o.push_back('}'); if (p.cache.ephemeral) { sep(); cache_point(o); } } }
    else if (p.type == "thinking") {
      if (p.text) {
        sep(); o += "{\"reasoningContent\":{\"reasoningText\":{\"text\":"; oj::enc_str(o, *p.text);
        if (p.signature && !p.signature->empty()) { o += ",\"signature\":"; oj::enc_str(o, *p.signature); }
        o += "}}}"; if (p.cache.ephemeral) { sep(); cache_point(o); }
      }
    } else if (p.type == "redacted_thinking") {
      if (p.has_redacted) {
        if (!p.redacted_is_bytes) return invalid("redacted_content must be a binary/bytes value in bedrock");
        sep(); o += "{\"reasoningContent\":{";
        if (!p.redacted.empty()) { o += "\"redactedContent\":\""; o += oj::b64enc(p.redacted); o += "\""; }
        o += "}}"; if (p.cache.ephemeral) { sep(); cache_point(o); }
      }
    } else if (p.type == "refusal") {
      if (p.refusal) { sep(); o += "{\"text\":"; oj::enc_str(o, *p.refusal); o.push_back('}'); if (p.cache.ephemeral) { sep(); cache_point(o); } }
    }
  }
  for (auto& tc : m.tool_calls) {
    // unmarshalToolCallArguments (openai_awsbedrock.go:299-305): arguments → map[string]any
    Value args; std::string perr;
    if (!oj::parse(tc.arguments, args, perr) || !(args.is_obj() || args.is_null()))
      return internal("failed to unmarshal tool call arguments: " + perr);
    if (!tc.id) return internal("nil tool call id");  // *toolCall.ID nil dereference in the reference
    sep(); o += "{\"toolUse\":{\"name\":"; oj::enc_str(o, tc.name); o += ",\"input\":";
    if (args.is_null()) o += "null"; else oj::enc_any(o, args);
    o += ",\"toolUseId\":"; oj::enc_str(o, *tc.id); o += "}}";
  }
  o.push_back(']'); return {};
}

inline Error system_blocks(const Message& m, std::string& o, bool& first, const char* what) {  // :422-448, 522-551
  auto sep = [&] { if (!first) o.push_back(','); first = false; };
  if (m.ck == Message::String) { sep(); o += "{\"text\":"; oj::enc_str(o, m.content_str); o.push_back('}'); return {}; }
  if (m.ck == Message::TextParts) {
    for (auto& p : m.text_parts) { sep(); o += "{\"text\":"; oj::enc_str(o, p.text); o.push_back('}'); if (p.cache.ephemeral) { sep(); cache_point(o); } }
    return {};
  }
  return invalid(std::string("unexpected content type for ") + what + " message");
}

inline Error tool_result_block(const Message& m, std::string& o) {  // :451-486
  o += "{\"toolResult\":{\"content\":[";
  if (m.ck == Message::String) { o += "{\"text\":"; oj::enc_str(o, m.content_str); o.push_back('}'); }
  else if (m.ck == Message::TextParts) { bool f = true; for (auto& p : m.text_parts) { if (!f) o.push_back(','); f = false; o += "{\"text\":"; oj::enc_str(o, p.text); o.push_back('}'); } }
  else return invalid("message 'content' must be a string or an array");
  o += "],\"status\":null,\"toolUseId\":"; oj::enc_str(o, m.tool_call_id); o += "}}";
  return {};
}

inline TranslateResult request_body(const ChatReq& r, const std::string& model_override) {  // :91-159
  TranslateResult res;
  res.stream = r.stream; res.model = r.model;
  res.request_model = model_override.empty() ? r.model : model_override;
  std::string o = "{";
  if (r.thinking != ChatReq::ThNone) {  // :57-78,130-135; empty map (adaptive) is dropped by omitempty
    if (r.thinking == ChatReq::ThEnabled) { o += "\"additionalModelRequestFields\":{\"thinking\":{\"budget_tokens\":" + std::to_string(r.thinking_budget) + ",\"type\":\"enabled\"}},"; }
    else if (r.thinking == ChatReq::ThDisabled) o += "\"additionalModelRequestFields\":{\"thinking\":{\"type\":\"disabled\"}},";
  }
  o += "\"inferenceConfig\":{"; bool f = true;
  auto sep = [&] { if (!f) o.push_back(','); f = false; };
  std::optional<int64_t> mt = r.max_completion_tokens ? r.max_completion_tokens : r.max_tokens;  // cmp.Or on pointers
  if (mt) { sep(); o += "\"maxTokens\":" + std::to_string(*mt); }
  std::vector<std::string> stops;
  if (r.stop_is_string) stops = {r.stop_string}; else if (r.stop_array) stops = *r.stop_array;
  if (!stops.empty()) { sep(); o += "\"stopSequences\":["; for (size_t i = 0; i < stops.size(); i++) { if (i) o.push_back(','); oj::enc_str(o, stops[i]); } o.push_back(']'); }
  if (r.temperature) { sep(); o += "\"temperature\":"; oj::enc_f64(o, *r.temperature); }
  if (r.top_p) { sep(); o += "\"topP\":"; oj::enc_f64(o, *r.top_p); }
  o += "},\"messages\":[";
  std::string sys; bool sys_first = true; bool mfirst = true;
  for (size_t i = 0; i < r.messages.size(); i++) {  // :489-585
    const Message& m = r.messages[i];
    auto msep = [&] { if (!mfirst) o.push_back(','); mfirst = false; };
    switch (m.role) {
      case Message::User: { std::string c; if (auto e = user_content(m, c)) { res.err = e; return res; } msep(); o += "{\"content\":" + c + ",\"role\":"; oj::enc_str(o, m.role_str); o.push_back('}'); break; }
      case Message::Assistant: { std::string c; if (auto e = assistant_content(m, c)) { res.err = e; return res; } msep(); o += "{\"content\":" + c + ",\"role\":"; oj::enc_str(o, m.role_str); o.push_back('}'); break; }
      case Message::System: if (auto e = system_blocks(m, sys, sys_first, "system")) { res.err = e; return res; } break;
      case Message::Developer: if (auto e = system_blocks(m, sys, sys_first, "developer")) { res.err = e; return res; } break;
      case Message::Tool: {
        std::string c = "[";
        if (auto e = tool_result_block(m, c)) { res.err = e; return res; }
        while (i + 1 < r.messages.size() && r.messages[i + 1].role_str == "tool") {  // coalesce (:559-575)
          c.push_back(','); if (auto e = tool_result_block(r.messages[i + 1], c)) { res.err = e; return res; } i++;
        }
        c.push_back(']');
        msep(); o += "{\"content\":" + c + ",\"role\":\"user\"}"; break;
      }
    }
  }
  o.push_back(']');
  if (!sys.empty()) o += ",\"system\":[" + sys + "]";
  if (!r.service_tier.empty()) { o += ",\"serviceTier\":{\"type\":"; oj::enc_str(o, r.service_tier); o.push_back('}'); }
  if (!r.tools.empty()) {  // :162-226
    std::string tc;
    if (r.tool_choice == ChatReq::TCString) {
      if (r.tool_choice_str == "auto") tc = "{\"auto\":{}}";
      else if (r.tool_choice_str == "required") tc = "{\"any\":{}}";
      else if (r.model.find("anthropic") != std::string::npos && r.model.find("claude") != std::string::npos) { tc = "{\"tool\":{\"name\":"; oj::enc_str(tc, r.tool_choice_str); tc += "}}"; }
    } else if (r.tool_choice == ChatReq::TCNamed) { tc = "{\"tool\":{\"name\":"; oj::enc_str(tc, r.tool_choice_fn); tc += "}}"; }
    o += ",\"toolConfig\":{";
    if (!tc.empty()) o += "\"toolChoice\":" + tc + ",";
    o += "\"tools\":["; bool tf = true;
    for (auto& t : r.tools) {
      if (!t.has_function) continue;
      if (!tf) o.push_back(','); tf = false;
      o += "{\"toolSpec\":{";
      if (!t.description.empty()) { o += "\"description\":"; oj::enc_str(o, t.description); o.push_back(','); }
      o += "\"inputSchema\":{\"json\":"; if (t.parameters) oj::enc_any(o, *t.parameters); else o += "null";
      o += "},\"name\":"; oj::enc_str(o, t.name); o.push_back('}');
      if (t.cache.ephemeral) o += ",\"cachePoint\":{\"type\":\"default\"}";
      o.push_back('}');
    }
    o += "]}";
  }
  o.push_back('}');
  res.body_kind = BYTES; res.body = std::move(o);
  res.headers.push_back({":path", "/model/" + path_escape(res.request_model) + (r.stream ? "/converse-stream" : "/converse")});
  res.headers.push_back({"content-length", std::to_string(res.body.size())});
  return res;
}
}  // namespace bedrock

// ---------------------------------------------------------------- T1: OpenAI passthrough
namespace openai_passthrough {
// original = the body as left by ParseBody (i.e. after P2's mutation when that applied).
inline TranslateResult request_body(std::string_view original, const Value& root, const ChatReq& r, const std::string& prefix, const std::string& model_override, bool force) {
  TranslateResult res; res.stream = r.stream; res.model = r.model; res.request_model = r.model;
  std::string nb; bool has = false;
  if (!model_override.empty()) {
    std::string lit; sjson_stringify(lit, model_override);
    nb = sjson_set_raw(original, root, "model", "", lit); has = true; res.request_model = model_override;
  }
  // path.Join("/", prefix, "chat/completions")
  std::string path = "/"; {
    std::string pfx = prefix; while (!pfx.empty() && pfx.front() == '/') pfx.erase(0, 1); while (!pfx.empty() && pfx.back() == '/') pfx.pop_back();
    if (!pfx.empty()) path += pfx + "/"; path += "chat/completions";
  }
  res.headers.push_back({":path", path});
  if (force && (!has || nb.empty())) { nb.assign(original); has = true; }
  if (has && !nb.empty()) { res.body_kind = BYTES; res.body = nb; res.headers.push_back({"content-length", std::to_string(nb.size())}); }
  return res;
}
}  // namespace openai_passthrough

// ---------------------------------------------------------------- Azure OpenAI (openai_azureopenai.go:37-62)
namespace azure {
inline TranslateResult request_body(std::string_view original, const ChatReq& r, const std::string& api_version, const std::string& model_override, bool force) {
  TranslateResult res; res.stream = r.stream; res.model = r.model;
  res.request_model = model_override.empty() ? r.model : model_override;
  res.headers.push_back({":path", "/openai/deployments/" + res.request_model + "/chat/completions?api-version=" + api_version});
  if (force) res.headers.push_back({"content-length", std::to_string(original.size())});
  return res;  // body nil: never mutated
}
}  // namespace azure

// ---------------------------------------------------------------- T4: OpenAI → Anthropic (GCP rawPredict / AWS InvokeModel)
// internal/translator/anthropic_helper.go:131-221,265-569,661-747; openai_gcpanthropic.go:56-100; openai_awsanthropic.go:54-100.
// The body layout is produced by anthropic-sdk-go v1.38.0 (MessageNewParams marshal), which is not in the tree.  Only
// what the reference's goldens pin (tests/data-plane/testupstream_test.go:203,216,291,355,369,549) is restated:
// max_tokens, messages (text / tool_use / tool_result blocks, cache_control {"type":"ephemeral"}), system, then the
// sjson-appended "stream" (GCP) and "anthropic_version".  The sampling parameters (anthropic_helper.go:726-738: temperature with
// its 0..1 validation :67-72, top_p, stop -> stop_sequences) follow the declaration order of MessageNewParams in the SDK
// (required fields, then the param.Opt scalars temperature / top_k / top_p, then metadata … stop_sequences, system, thinking,
// tool_choice, tools) — the same rule the pinned fields obey (max_tokens, messages, …, system; text, cache_control, type);
// their CONTENT is pinned by openai_gcpanthropic_test.go:225-261 (gjson / struct compares), their byte position by no
// in-tree fixture ("parity unpinned" for the order).  tools, tool_choice, thinking, output_config: DECLINED, never a guess.
namespace anthropic {
inline Error unpinned(const char* what) { return Error{DECLINED, std::string("parity unpinned: ") + what}; }

inline void text_block(std::string& o, std::string_view text, bool cache) {
  o += "{\"text\":"; oj::enc_str(o, text);
  if (cache) o += ",\"cache_control\":{\"type\":\"ephemeral\"}";
  o += ",\"type\":\"text\"}";
}

inline TranslateResult request_body(const ChatReq& r, const Value& root, bool gcp, const std::string& model_override, const std::string& api_version) {
  TranslateResult res; res.stream = r.stream; res.model = r.model;
  res.request_model = model_override.empty() ? r.model : model_override;
  auto fail = [&](Error e) { res.err = e; return res; };
  if (r.stop_array && r.stop_array->empty()) return fail(unpinned("empty stop array (nil or [] after the openai-go union decode)"));
  if (!r.tools.empty()) return fail(unpinned("tools"));
  if (r.thinking != ChatReq::ThNone) return fail(unpinned("thinking"));
  if (!r.reasoning_effort.empty()) return fail(unpinned("output_config.effort"));
  if (r.response_format) return fail(unpinned("output_config.format"));
  std::optional<int64_t> mt = r.max_completion_tokens ? r.max_completion_tokens : r.max_tokens;
  std::string msgs, sys; bool mfirst = true, sfirst = true;
  auto cache_raw_ok = [&](const CacheCtl& c) { return !c.present || true; };
  (void)cache_raw_ok;
  for (size_t i = 0; i < r.messages.size();) {
    const Message& m = r.messages[i];
    auto msep = [&] { if (!mfirst) msgs.push_back(','); mfirst = false; };
    if (m.role == Message::System || m.role == Message::Developer) {  // one block per message, parts concatenated (:338-362)
      std::string text; bool cache = false;
      if (m.ck == Message::String) text = m.content_str;
      else if (m.ck == Message::TextParts) for (auto& p : m.text_parts) { text += p.text; if (p.cache.ephemeral) cache = true; }
      if (text.empty()) return fail(unpinned("empty system text block"));
      if (!sfirst) sys.push_back(','); sfirst = false;
      text_block(sys, text, cache);
      i++; continue;
    }
    if (m.role == Message::User) {
      std::string c;
      if (m.ck == Message::None) return fail(internal("unsupported OpenAI content type: <nil>"));
      if (m.ck == Message::String) { if (m.content_str.empty()) return fail(unpinned("nil content")); text_block(c, m.content_str, false); }
      else {
        if (m.user_parts.empty()) return fail(unpinned("empty content array"));
        bool f = true;
        for (auto& p : m.user_parts) {
          if (p.k != UserPart::Text) return fail(unpinned("non-text content part"));
          if (p.text.text.empty()) return fail(unpinned("empty text block"));
          if (!f) c.push_back(','); f = false;
          text_block(c, p.text.text, p.text.cache.ephemeral);
        }
      }
      msep(); msgs += "{\"content\":[" + c + "],\"role\":\"user\"}"; i++; continue;
    }
    if (m.role == Message::Assistant) {
      std::string c; bool f = true;
      auto sep = [&] { if (!f) c.push_back(','); f = false; };
      std::vector<AsstPart> parts;
      if (m.ck == Message::String) { if (!m.content_str.empty()) { sep(); text_block(c, m.content_str, false); } }
      else if (m.ck == Message::AsstParts || m.ck == Message::AsstSingle) parts = m.asst_parts;
      for (auto& p : parts) {
        if (p.type == "refusal") { if (p.refusal) { if (p.refusal->empty()) return fail(unpinned("empty text block")); sep(); text_block(c, *p.refusal, false); } }
        else if (p.type == "text") { if (p.text) { if (p.text->empty()) return fail(unpinned("empty text block")); sep(); text_block(c, *p.text, p.cache.ephemeral); } }
        else if (p.type == "thinking" || p.type == "redacted_thinking") return fail(unpinned("thinking blocks"));
        else return fail(internal("content type not supported: " + p.type));
      }
      for (auto& tc : m.tool_calls) {
        Value args; std::string perr;
        if (!oj::parse(tc.arguments, args, perr) || !(args.is_obj() || args.is_null())) return fail(internal("failed to unmarshal tool call arguments: " + perr));
        if (!tc.id) return fail(unpinned("nil tool call id (panics in the reference)"));
        if (tc.cache.ephemeral) return fail(unpinned("cache_control on tool_use"));
        sep(); c += "{\"id\":"; oj::enc_str(c, *tc.id); c += ",\"input\":";
        if (args.is_null()) c += "null"; else oj::enc_any(c, args);
        c += ",\"name\":"; oj::enc_str(c, tc.name); c += ",\"type\":\"tool_use\"}";
      }
      if (f) return fail(unpinned("empty assistant content"));
      msep(); msgs += "{\"content\":[" + c + "],\"role\":\"assistant\"}"; i++; continue;
    }
    // tool: consecutive tool messages aggregate into one user message (:498-560)
    std::string c; bool f = true;
    while (i < r.messages.size() && r.messages[i].role_str == "tool") {
      const Message& t = r.messages[i];
      std::string blocks;
      if (t.ck == Message::None) return fail(internal("unsupported ContentUnion value type: <nil>"));
      if (t.ck == Message::String) { if (t.content_str.empty()) return fail(unpinned("nil tool result content")); text_block(blocks, t.content_str, false); }
      else {
        if (t.text_parts.empty()) return fail(unpinned("nil tool result content"));
        bool bf = true;
        for (auto& p : t.text_parts) { if (p.cache.ephemeral) return fail(unpinned("cache_control on tool_result")); if (p.text.empty()) return fail(unpinned("empty text block")); if (!bf) blocks.push_back(','); bf = false; text_block(blocks, p.text, false); }
      }
      bool is_error = false;
      if (t.ck == Message::String) { Value cm; std::string pe; if (oj::parse(t.content_str, cm, pe) && cm.is_obj() && cm.get("error")) is_error = true; }
      if (!f) c.push_back(','); f = false;
      c += "{\"tool_use_id\":"; oj::enc_str(c, t.tool_call_id); c += is_error ? ",\"is_error\":true" : ",\"is_error\":false";
      c += ",\"content\":[" + blocks + "],\"type\":\"tool_result\"}";
      i++;
    }
    msep(); msgs += "{\"content\":[" + c + "],\"role\":\"user\"}";
  }
  if (mfirst) return fail(unpinned("no messages"));
  // validateTemperatureForAnthropic (anthropic_helper.go:67-72), after the message / tool translation errors (:671-679,726-731)
  if (r.temperature && (*r.temperature < 0.0 || *r.temperature > 1.0)) {
    char tb[64]; snprintf(tb, sizeof tb, "%.2f", *r.temperature);
    return fail(invalid(std::string("invalid request body: temperature ") + tb + " is not supported by Anthropic (must be between 0.0 and 1.0)"));
  }
  std::string o = "{\"max_tokens\":" + std::to_string(mt.value_or(0)) + ",\"messages\":[" + msgs + "]";
  if (r.temperature) { o += ",\"temperature\":"; oj::enc_f64(o, *r.temperature); }
  if (r.top_p) { o += ",\"top_p\":"; oj::enc_f64(o, *r.top_p); }
  if (r.stop_is_string || r.stop_array) {
    o += ",\"stop_sequences\":[";
    if (r.stop_is_string) oj::enc_str(o, r.stop_string);
    else for (size_t k = 0; k < r.stop_array->size(); k++) { if (k) o.push_back(','); oj::enc_str(o, (*r.stop_array)[k]); }
    o += "]";
  }
  if (!sys.empty()) o += ",\"system\":[" + sys + "]";
  if (gcp && r.stream) o += ",\"stream\":true";
  std::string ver = !api_version.empty() ? api_version : (gcp ? "vertex-2023-10-16" : "bedrock-2023-05-31");
  o += ",\"anthropic_version\":"; sjson_stringify(o, ver); o += "}";
  res.body_kind = BYTES; res.body = o;
  std::string path = gcp ? "publishers/anthropic/models/" + res.request_model + (r.stream ? ":streamRawPredict" : ":rawPredict")
                         : "/model/" + path_escape(res.request_model) + (r.stream ? "/invoke-with-response-stream" : "/invoke");
  res.headers.push_back({":path", path});
  res.headers.push_back({"content-length", std::to_string(o.size())});
  return res;
}
}  // namespace anthropic

TranslateResult gemini_request_body(const ChatReq& r, const std::string& model_override);   // gemini.hpp

enum Schema : int { SCHEMA_OPENAI = 0, SCHEMA_AWS_BEDROCK = 1, SCHEMA_AZURE_OPENAI = 2, SCHEMA_GCP_VERTEX = 3, SCHEMA_GCP_ANTHROPIC = 4, SCHEMA_AWS_ANTHROPIC = 5 };

// ParseBody + GetTranslator + RequestBody for one /v1/chat/completions body
// (routerProcessor.ProcessRequestBody processor_impl.go:211-295 then upstreamProcessor.ProcessRequestHeaders :307-398).
inline TranslateResult chat_translate(int schema, std::string_view body, const std::string& model_override, const std::string& prefix, bool cost_configured, bool force) {
  TranslateResult res;
  Value root; std::string perr;
  if (!oj::parse(body, root, perr)) { res.err = bad("malformed request: failed to parse JSON for /v1/chat/completions: " + perr); return res; }
  ChatReq r;
  if (auto e = parse_chat_request(root, r)) { res.err = bad("malformed request: failed to parse JSON for /v1/chat/completions: " + e.msg); return res; }
  std::string mutated; bool has_mut = false;
  std::string_view cur = body; Value root2; const Value* curroot = &root;
  if (r.stream && cost_configured && !(r.has_stream_options && r.include_usage)) {  // endpointspec.go:107-123
    mutated = sjson_set_raw(body, root, "stream_options", "include_usage", "true"); has_mut = true;
    r.has_stream_options = true; r.include_usage = true;
    cur = mutated; std::string e2; oj::parse(cur, root2, e2); curroot = &root2;
  }
  switch (schema) {
    case SCHEMA_AWS_BEDROCK: res = bedrock::request_body(r, model_override); break;
    case SCHEMA_OPENAI: res = openai_passthrough::request_body(cur, *curroot, r, prefix, model_override, force || has_mut); break;
    case SCHEMA_AZURE_OPENAI: res = azure::request_body(cur, r, prefix /* carries the api-version for this schema */, model_override, force || has_mut); break;
    case SCHEMA_GCP_ANTHROPIC: res = anthropic::request_body(r, root, true, model_override, prefix /* api version */); break;
    case SCHEMA_AWS_ANTHROPIC: res = anthropic::request_body(r, root, false, model_override, prefix /* api version */); break;
    case SCHEMA_GCP_VERTEX: res = gemini_request_body(r, model_override); break;
    default: res.err = Error{DECLINED, "schema not restated yet"}; break;
  }
  res.mutated_body = mutated; res.has_mutated = has_mut; res.model = r.model; res.stream = r.stream;
  return res;
}

}  // namespace oracle
