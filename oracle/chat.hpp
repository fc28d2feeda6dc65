// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the reference's /v1/chat/completions request decode:
//   endpointspec.ChatCompletionsEndpointSpec.ParseBody   internal/endpointspec/endpointspec.go:98-125
//   openai.ChatCompletionRequest                          internal/apischema/openai/openai.go:947-1131
//   message union / content unions                        internal/apischema/openai/openai.go:185-224,239-312,377-401,409-481
//   ContentUnion fast path                                internal/apischema/openai/openai.go:287-308, union.go:152-185
//   tool choice / thinking / response_format unions       internal/apischema/openai/openai.go:710-745,869-945,1194-1210
// Every decode error here is what json.Unmarshal would return to ParseBody, which wraps
// it as ErrMalformedRequest (HTTP 400).
#pragma once
#include <cerrno>
#include <cmath>
#include <optional>

#include "ojson.hpp"

namespace oracle {
using oj::T;
using oj::Value;

enum Status : int { OK = 0, MALFORMED_400 = 1, INVALID_422 = 2, INTERNAL = 3, DECLINED = 4 };
enum BodyKind : int { UNCHANGED = 0, BYTES = 1, EMPTY = 2 };

struct Error { int status = OK; std::string msg; explicit operator bool() const { return status != OK; } };
inline Error bad(std::string m) { return Error{MALFORMED_400, std::move(m)}; }
inline Error invalid(std::string m) { return Error{INVALID_422, std::move(m)}; }
inline Error internal(std::string m) { return Error{INTERNAL, std::move(m)}; }

// ---- typed decode primitives (null ⇒ zero value, no error: encoding/json & sonic)
inline Error want_str(const Value* v, const char* f, std::optional<std::string>& out) {
  if (!v || v->is_null()) return {};
  if (!v->is_str()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type string");
  out = v->s; return {};
}
inline Error want_bool(const Value* v, const char* f, std::optional<bool>& out) {
  if (!v || v->is_null()) return {};
  if (!v->is_bool()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type bool");
  out = v->t == T::True; return {};
}
inline Error want_int(const Value* v, const char* f, std::optional<int64_t>& out) {
  if (!v || v->is_null()) return {};
  int64_t x;
  if (!v->is_num() || !oj::num_to_i64(v->s, x)) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type int64");
  out = x; return {};
}
inline Error want_f64(const Value* v, const char* f, std::optional<double>& out) {
  if (!v || v->is_null()) return {};
  if (!v->is_num()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type float64");
  errno = 0; double d = strtod(v->s.c_str(), nullptr);
  if (errno == ERANGE && std::isinf(d)) return bad(std::string("json: number out of range for field ") + f);
  out = d; return {};
}
inline Error want_f32(const Value* v, const char* f, std::optional<float>& out) {
  if (!v || v->is_null()) return {};
  if (!v->is_num()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type float32");
  errno = 0; float d = strtof(v->s.c_str(), nullptr);
  if (errno == ERANGE && std::isinf(d)) return bad(std::string("json: number out of range for field ") + f);
  out = d; return {};
}
inline Error want_obj(const Value* v, const char* f, const Value*& out) {
  out = nullptr;
  if (!v || v->is_null()) return {};
  if (!v->is_obj()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type struct");
  out = v; return {};
}
inline Error want_arr(const Value* v, const char* f, const Value*& out) {
  out = nullptr;
  if (!v || v->is_null()) return {};
  if (!v->is_arr()) return bad(std::string("json: cannot unmarshal into Go struct field ") + f + " of type slice");
  out = v; return {};
}
inline Error want_str_arr(const Value* v, const char* f, std::optional<std::vector<std::string>>& out) {
  const Value* a; if (auto e = want_arr(v, f, a)) return e;
  if (!a) return {};
  std::vector<std::string> r;
  for (auto& x : a->arr) { std::optional<std::string> s; if (auto e = want_str(&x, f, s)) return e; r.push_back(s.value_or("")); }
  out = std::move(r); return {};
}

// gjson.GetBytes(data, key).String() on a raw JSON value (openai.go:186,418,711,912):
// first matching key; String() = decoded string / raw literal / "" for null.
inline bool gjson_top_string(const Value& v, const char* key, std::string& out) {
  if (!v.is_obj()) return false;
  const Value* r = v.get_first(key);
  if (!r) return false;
  switch (r->t) {
    case T::String: out = r->s; break;
    case T::Number: out = r->s; break;
    case T::True: out = "true"; break;
    case T::False: out = "false"; break;
    case T::Null: out = ""; break;
    default: out = "<json>"; break;
  }
  return true;
}

// AnthropicContentFields (openai.go:404-407): inline `cache_control` object; enabled when
// type == "ephemeral" (translator/anthropic_helper.go:261-263).
struct CacheCtl { bool present = false; bool ephemeral = false; std::string raw; };
inline Error parse_cache(const Value& o, CacheCtl& c) {
  const Value* cc = o.get("cache_control");
  const Value* co; if (auto e = want_obj(cc, "cache_control", co)) return e;
  if (co) {
    c.present = true;
    std::optional<std::string> t, ttl;
    if (auto e = want_str(co->get("type"), "cache_control.type", t)) return e;
    if (auto e = want_str(co->get("ttl"), "cache_control.ttl", ttl)) return e;
    c.ephemeral = t && *t == "ephemeral";
  }
  return {};
}

struct TextPart { std::string text; std::string type; CacheCtl cache; };
struct UserPart {
  enum K { Text, InputAudio, ImageURL, File } k = Text;
  TextPart text; std::string image_url, image_detail; CacheCtl cache;
  std::string audio_data, audio_format;
};
struct AsstPart {
  std::string type; std::optional<std::string> refusal, text, signature;
  bool has_redacted = false, redacted_is_bytes = false; std::string redacted; CacheCtl cache;
};
struct ToolCall { std::optional<std::string> id; std::string name, arguments, type; CacheCtl cache; };

struct Message {
  enum Role { User, Assistant, System, Developer, Tool } role = User;
  std::string role_str;
  // content
  enum CK { None, String, TextParts, UserParts, AsstParts, AsstSingle } ck = None;
  std::string content_str;
  std::vector<TextPart> text_parts;   // system / developer / tool
  std::vector<UserPart> user_parts;   // user
  std::vector<AsstPart> asst_parts;   // assistant (array or single object)
  std::string name, refusal, tool_call_id, audio_id;
  std::vector<ToolCall> tool_calls;
};

struct ToolDef {
  std::string type; bool has_function = false;
  std::string name, description; bool strict = false;
  const Value* parameters = nullptr;  // `any`
  CacheCtl cache;
  bool has_google_search = false;
};

struct ChatReq {
  std::vector<Message> messages;
  std::string model;
  std::optional<float> frequency_penalty, presence_penalty;
  std::optional<int64_t> max_tokens, max_completion_tokens, n, seed, top_logprobs;
  std::optional<bool> logprobs, parallel_tool_calls;
  std::string service_tier, user, reasoning_effort, verbosity, guided_regex;
  bool stop_is_string = false; std::string stop_string; std::optional<std::vector<std::string>> stop_array;
  bool stream = false;
  bool has_stream_options = false, include_usage = false;
  std::optional<double> temperature, top_p;
  std::vector<ToolDef> tools;
  enum TC { TCNone, TCString, TCNamed } tool_choice = TCNone;
  std::string tool_choice_str, tool_choice_fn, tool_choice_type;
  enum TK { ThNone, ThEnabled, ThDisabled, ThAdaptive } thinking = ThNone;
  int64_t thinking_budget = 0; bool thinking_include_thoughts = false;
  const Value* response_format = nullptr; std::string response_format_type;
  const Value* guided_json = nullptr;
  std::optional<std::vector<std::string>> guided_choice;
  const Value* generation_config = nullptr; const Value* safety_settings = nullptr;
};

// ChatCompletionContentPartTextParam (openai.go:92-99)
inline Error parse_text_part(const Value& v, TextPart& p) {
  const Value* o; if (auto e = want_obj(&v, "ChatCompletionContentPartTextParam", o)) return e;
  if (!o) return {};  // null element ⇒ zero struct
  std::optional<std::string> t, ty;
  if (auto e = want_str(o->get("text"), "text", t)) return e;
  if (auto e = want_str(o->get("type"), "type", ty)) return e;
  p.text = t.value_or(""); p.type = ty.value_or("");
  return parse_cache(*o, p.cache);
}

// ContentUnion.UnmarshalJSON (openai.go:287-308): string | []TextParam, anything else is an error.
inline Error parse_content_union(const Value& v, Message& m) {
  if (v.is_str()) { m.ck = Message::String; m.content_str = v.s; return {}; }
  if (v.is_arr()) {
    m.ck = Message::TextParts;
    for (auto& x : v.arr) {
      TextPart p;
      if (auto e = parse_text_part(x, p)) return bad("cannot unmarshal content as []ChatCompletionContentPartTextParam: " + e.msg);
      m.text_parts.push_back(std::move(p));
    }
    return {};
  }
  return bad("invalid content type (must be string or array of ChatCompletionContentPartTextParam)");
}

// ChatCompletionContentPartUserUnionParam.UnmarshalJSON (openai.go:185-224)
inline Error parse_user_part(const Value& v, UserPart& p) {
  std::string ty;
  if (!gjson_top_string(v, "type", ty)) return bad("chat content does not have type");
  if (ty == "text") { p.k = UserPart::Text; return parse_text_part(v, p.text); }
  if (ty == "input_audio") {
    p.k = UserPart::InputAudio;
    const Value* ia; if (auto e = want_obj(v.get("input_audio"), "input_audio", ia)) return e;
    if (ia) {
      std::optional<std::string> d, f;
      if (auto e = want_str(ia->get("data"), "data", d)) return e;
      if (auto e = want_str(ia->get("format"), "format", f)) return e;
      p.audio_data = d.value_or(""); p.audio_format = f.value_or("");
    }
    std::optional<std::string> t; if (auto e = want_str(v.get("type"), "type", t)) return e;
    return parse_cache(v, p.cache);
  }
  if (ty == "image_url") {
    p.k = UserPart::ImageURL;
    const Value* iu; if (auto e = want_obj(v.get("image_url"), "image_url", iu)) return e;
    if (iu) {
      std::optional<std::string> u, d;
      if (auto e = want_str(iu->get("url"), "url", u)) return e;
      if (auto e = want_str(iu->get("detail"), "detail", d)) return e;
      p.image_url = u.value_or(""); p.image_detail = d.value_or("");
    }
    std::optional<std::string> t; if (auto e = want_str(v.get("type"), "type", t)) return e;
    return parse_cache(v, p.cache);
  }
  if (ty == "file") {
    p.k = UserPart::File;
    const Value* f; if (auto e = want_obj(v.get("file"), "file", f)) return e;
    if (f) for (const char* k : {"file_data", "file_id", "filename"}) { std::optional<std::string> s; if (auto e = want_str(f->get(k), k, s)) return e; }
    std::optional<std::string> t; if (auto e = want_str(v.get("type"), "type", t)) return e;
    return parse_cache(v, p.cache);
  }
  return bad("unknown ChatCompletionContentPartUnionParam type: " + ty);
}

// ChatCompletionAssistantMessageParamContent (openai.go:546-560) + RedactedContentUnion (openai.go:1735-1755)
inline Error parse_asst_part(const Value& v, AsstPart& p) {
  const Value* o; if (auto e = want_obj(&v, "ChatCompletionAssistantMessageParamContent", o)) return e;
  if (!o) return {};
  std::optional<std::string> ty;
  if (auto e = want_str(o->get("type"), "type", ty)) return e;
  p.type = ty.value_or("");
  if (auto e = want_str(o->get("refusal"), "refusal", p.refusal)) return e;
  if (auto e = want_str(o->get("text"), "text", p.text)) return e;
  if (auto e = want_str(o->get("signature"), "signature", p.signature)) return e;
  if (const Value* rc = o->get("redactedContent"); rc && !rc->is_null()) {
    if (!rc->is_str()) return bad("redactedContent must be either []byte (base64 encoded) or string");
    p.has_redacted = true;
    std::string dec;
    if (oj::b64dec(rc->s, dec)) { p.redacted_is_bytes = true; p.redacted = dec; } else p.redacted = rc->s;
  }
  return parse_cache(*o, p.cache);
}

inline Error parse_tool_calls(const Value* v, std::vector<ToolCall>& out) {
  const Value* a; if (auto e = want_arr(v, "tool_calls", a)) return e;
  if (!a) return {};
  for (auto& x : a->arr) {
    ToolCall tc;
    const Value* o; if (auto e = want_obj(&x, "tool_calls[]", o)) return e;
    if (o) {
      if (auto e = want_str(o->get("id"), "id", tc.id)) return e;
      const Value* f; if (auto e = want_obj(o->get("function"), "function", f)) return e;
      if (f) {
        std::optional<std::string> a2, n2;
        if (auto e = want_str(f->get("arguments"), "arguments", a2)) return e;
        if (auto e = want_str(f->get("name"), "name", n2)) return e;
        tc.arguments = a2.value_or(""); tc.name = n2.value_or("");
      }
      std::optional<std::string> t; if (auto e = want_str(o->get("type"), "type", t)) return e;
      tc.type = t.value_or("");
      if (auto e = parse_cache(*o, tc.cache)) return e;
    }
    out.push_back(std::move(tc));
  }
  return {};
}

// ChatCompletionMessageParamUnion.UnmarshalJSON (openai.go:417-461)
inline Error parse_message(const Value& v, Message& m) {
  std::string role;
  if (!gjson_top_string(v, "role", role)) return bad("chat message does not have role");
  if (role == "user") m.role = Message::User;
  else if (role == "assistant") m.role = Message::Assistant;
  else if (role == "system") m.role = Message::System;
  else if (role == "developer") m.role = Message::Developer;
  else if (role == "tool") m.role = Message::Tool;
  else return bad("unknown ChatCompletionMessageParam type: " + role);
  // the role-specific struct decode: `role` must be a JSON string there (last occurrence)
  std::optional<std::string> rs; if (auto e = want_str(v.get("role"), "role", rs)) return e;
  m.role_str = rs.value_or("");
  const Value* c = v.get("content");
  switch (m.role) {
    case Message::User: {
      if (c) {  // StringOrUserRoleContentUnion (openai.go:377-397): string (null ⇒ "") | []part
        if (c->is_str() || c->is_null()) { m.ck = Message::String; m.content_str = c->is_str() ? c->s : ""; }
        else {
          bool ok = c->is_arr();
          if (ok) for (auto& x : c->arr) { UserPart p; if (parse_user_part(x, p)) { ok = false; break; } m.user_parts.push_back(std::move(p)); }
          if (!ok) return bad("cannot unmarshal JSON data as string or array of content parts");
          m.ck = Message::UserParts;
        }
      }
      std::optional<std::string> n; if (auto e = want_str(v.get("name"), "name", n)) return e; m.name = n.value_or("");
      break;
    }
    case Message::Assistant: {
      if (c) {  // StringOrAssistantRoleContentUnion (openai.go:243-266): string | []content | content
        if (c->is_str() || c->is_null()) { m.ck = Message::String; m.content_str = c->is_str() ? c->s : ""; }
        else if (c->is_arr()) {
          for (auto& x : c->arr) { AsstPart p; if (parse_asst_part(x, p)) return bad("cannot unmarshal JSON data as string or assistant content parts"); m.asst_parts.push_back(std::move(p)); }
          m.ck = Message::AsstParts;
        } else {
          AsstPart p; if (parse_asst_part(*c, p)) return bad("cannot unmarshal JSON data as string or assistant content parts");
          m.asst_parts.push_back(std::move(p)); m.ck = Message::AsstSingle;
        }
      }
      const Value* au; if (auto e = want_obj(v.get("audio"), "audio", au)) return e;
      if (au) { std::optional<std::string> id; if (auto e = want_str(au->get("id"), "id", id)) return e; m.audio_id = id.value_or(""); }
      std::optional<std::string> n, r;
      if (auto e = want_str(v.get("name"), "name", n)) return e;
      if (auto e = want_str(v.get("refusal"), "refusal", r)) return e;
      m.name = n.value_or(""); m.refusal = r.value_or("");
      if (auto e = parse_tool_calls(v.get("tool_calls"), m.tool_calls)) return e;
      break;
    }
    case Message::System: case Message::Developer: {
      if (c) if (auto e = parse_content_union(*c, m)) return e;
      std::optional<std::string> n; if (auto e = want_str(v.get("name"), "name", n)) return e; m.name = n.value_or("");
      break;
    }
    case Message::Tool: {
      if (c) if (auto e = parse_content_union(*c, m)) return e;
      std::optional<std::string> id; if (auto e = want_str(v.get("tool_call_id"), "tool_call_id", id)) return e; m.tool_call_id = id.value_or("");
      break;
    }
  }
  return {};
}

inline Error parse_tools(const Value* v, std::vector<ToolDef>& out) {
  const Value* a; if (auto e = want_arr(v, "tools", a)) return e;
  if (!a) return {};
  for (auto& x : a->arr) {
    ToolDef t;
    const Value* o; if (auto e = want_obj(&x, "tools[]", o)) return e;
    if (o) {
      std::optional<std::string> ty; if (auto e = want_str(o->get("type"), "type", ty)) return e; t.type = ty.value_or("");
      const Value* f; if (auto e = want_obj(o->get("function"), "function", f)) return e;
      if (f) {
        t.has_function = true;
        std::optional<std::string> n, d; std::optional<bool> st;
        if (auto e = want_str(f->get("name"), "name", n)) return e;
        if (auto e = want_str(f->get("description"), "description", d)) return e;
        if (auto e = want_bool(f->get("strict"), "strict", st)) return e;
        t.name = n.value_or(""); t.description = d.value_or(""); t.strict = st.value_or(false);
        const Value* p = f->get("parameters"); if (p && !p->is_null()) t.parameters = p;
        if (auto e = parse_cache(*f, t.cache)) return e;
      }
      const Value* gs; if (auto e = want_obj(o->get("google_search"), "google_search", gs)) return e;
      t.has_google_search = gs != nullptr;
    }
    out.push_back(std::move(t));
  }
  return {};
}

// json.Unmarshal(body, &openai.ChatCompletionRequest{}) — endpointspec.go:102-105
inline Error parse_chat_request(const Value& root, ChatReq& r) {
  if (root.is_null()) return {};  // `null` decodes to the zero request without error
  if (!root.is_obj()) return bad("json: cannot unmarshal non-object into Go value of type openai.ChatCompletionRequest");
  {
    const Value* a; if (auto e = want_arr(root.get("messages"), "messages", a)) return e;
    if (a) for (auto& x : a->arr) { Message m; if (auto e = parse_message(x, m)) return e; r.messages.push_back(std::move(m)); }
  }
  std::optional<std::string> s;
  auto str_field = [&](const char* k, std::string& dst) -> Error { std::optional<std::string> t; if (auto e = want_str(root.get(k), k, t)) return e; dst = t.value_or(""); return {}; };
  if (auto e = str_field("model", r.model)) return e;
  if (auto e = want_f32(root.get("frequency_penalty"), "frequency_penalty", r.frequency_penalty)) return e;
  if (const Value* lb = root.get("logit_bias"); lb && !lb->is_null()) {
    if (!lb->is_obj()) return bad("json: cannot unmarshal logit_bias");
    for (auto& kv : lb->obj) { std::optional<int64_t> x; if (auto e = want_int(&kv.second, "logit_bias", x)) return e; }
  }
  if (auto e = want_bool(root.get("logprobs"), "logprobs", r.logprobs)) return e;
  if (auto e = want_int(root.get("top_logprobs"), "top_logprobs", r.top_logprobs)) return e;
  if (auto e = want_int(root.get("max_tokens"), "max_tokens", r.max_tokens)) return e;
  if (auto e = want_int(root.get("max_completion_tokens"), "max_completion_tokens", r.max_completion_tokens)) return e;
  if (auto e = want_int(root.get("n"), "n", r.n)) return e;
  if (auto e = want_f32(root.get("presence_penalty"), "presence_penalty", r.presence_penalty)) return e;
  if (const Value* rf = root.get("response_format"); rf && !rf->is_null()) {  // openai.go:710-745
    std::string ty;
    if (!gjson_top_string(*rf, "type", ty)) return bad("response format does not have type");
    if (ty != "text" && ty != "json_schema" && ty != "json_object") return bad("unsupported ChatCompletionResponseFormatType");
    std::optional<std::string> t2; if (auto e = want_str(rf->get("type"), "type", t2)) return e;
    if (ty == "json_schema") {
      const Value* js; if (auto e = want_obj(rf->get("json_schema"), "json_schema", js)) return e;
      if (js) {
        std::optional<std::string> n, d; std::optional<bool> st;
        if (auto e = want_str(js->get("name"), "name", n)) return e;
        if (auto e = want_str(js->get("description"), "description", d)) return e;
        if (auto e = want_bool(js->get("strict"), "strict", st)) return e;
      }
    }
    r.response_format = rf; r.response_format_type = ty;
  }
  if (auto e = want_int(root.get("seed"), "seed", r.seed)) return e;
  if (auto e = str_field("reasoning_effort", r.reasoning_effort)) return e;
  if (auto e = str_field("service_tier", r.service_tier)) return e;
  if (auto e = str_field("verbosity", r.verbosity)) return e;
  if (const Value* st = root.get("stop"); st && !st->is_null()) {
    // openai-go ChatCompletionNewParamsStopUnion: string | []string (apijson union decode;
    // any other JSON type has no matching variant ⇒ error).  Not in tree; see DESIGN.md.
    if (st->is_str()) { r.stop_is_string = true; r.stop_string = st->s; }
    else if (st->is_arr()) { if (auto e = want_str_arr(st, "stop", r.stop_array)) return e; }
    else return bad("json: cannot unmarshal stop into ChatCompletionNewParamsStopUnion");
  }
  { std::optional<bool> b; if (auto e = want_bool(root.get("stream"), "stream", b)) return e; r.stream = b.value_or(false); }
  { const Value* so; if (auto e = want_obj(root.get("stream_options"), "stream_options", so)) return e;
    if (so) { r.has_stream_options = true; std::optional<bool> b; if (auto e = want_bool(so->get("include_usage"), "include_usage", b)) return e; r.include_usage = b.value_or(false); } }
  if (auto e = want_f64(root.get("temperature"), "temperature", r.temperature)) return e;
  if (auto e = want_f64(root.get("top_p"), "top_p", r.top_p)) return e;
  if (auto e = parse_tools(root.get("tools"), r.tools)) return e;
  if (const Value* tc = root.get("tool_choice"); tc && !tc->is_null()) {  // openai.go:1194-1210
    if (tc->is_str()) { r.tool_choice = ChatReq::TCString; r.tool_choice_str = tc->s; }
    else {
      bool ok = tc->is_obj();
      std::optional<std::string> ty, fn;
      if (ok && want_str(tc->get("type"), "type", ty)) ok = false;
      if (ok) { const Value* f; if (want_obj(tc->get("function"), "function", f)) ok = false; else if (f && want_str(f->get("name"), "name", fn)) ok = false; }
      if (!ok) return bad("tool choice must be either string or ChatCompletionNamedToolChoice");
      r.tool_choice = ChatReq::TCNamed; r.tool_choice_type = ty.value_or(""); r.tool_choice_fn = fn.value_or("");
    }
  }
  if (auto e = want_bool(root.get("parallel_tool_calls"), "parallel_tool_calls", r.parallel_tool_calls)) return e;
  if (auto e = str_field("user", r.user)) return e;
  { std::optional<std::vector<std::string>> m; if (auto e = want_str_arr(root.get("modalities"), "modalities", m)) return e; }
  { const Value* au; if (auto e = want_obj(root.get("audio"), "audio", au)) return e;
    if (au) for (const char* k : {"voice", "format"}) { std::optional<std::string> t; if (auto e = want_str(au->get(k), k, t)) return e; } }
  { const Value* pr; if (auto e = want_obj(root.get("prediction"), "prediction", pr)) return e;
    if (pr) { std::optional<std::string> t; if (auto e = want_str(pr->get("type"), "type", t)) return e;
      if (const Value* c = pr->get("content")) { Message tmp; if (auto e = parse_content_union(*c, tmp)) return e; } } }
  { const Value* ws; if (auto e = want_obj(root.get("web_search_options"), "web_search_options", ws)) return e;
    if (ws) {
      std::optional<std::string> t; if (auto e = want_str(ws->get("search_context_size"), "search_context_size", t)) return e;
      const Value* ul; if (auto e = want_obj(ws->get("user_location"), "user_location", ul)) return e;
      if (ul) { if (auto e = want_str(ul->get("type"), "type", t)) return e;
        const Value* ap; if (auto e = want_obj(ul->get("approximate"), "approximate", ap)) return e;
        if (ap) for (const char* k : {"city", "region", "country"}) if (auto e = want_str(ap->get(k), k, t)) return e; } } }
  { const Value* gc; if (auto e = want_obj(root.get("generationConfig"), "generationConfig", gc)) return e; r.generation_config = gc;
    const Value* ss; if (auto e = want_arr(root.get("safetySettings"), "safetySettings", ss)) return e; r.safety_settings = ss; }
  if (auto e = want_str_arr(root.get("guided_choice"), "guided_choice", r.guided_choice)) return e;
  if (auto e = str_field("guided_regex", r.guided_regex)) return e;
  if (const Value* gj = root.get("guided_json"); gj && !gj->is_null()) r.guided_json = gj;
  if (const Value* th = root.get("thinking"); th && !th->is_null()) {  // openai.go:911-945
    std::string ty;
    if (!gjson_top_string(*th, "type", ty)) return bad("thinking config does not have a type");
    std::optional<std::string> t2;
    if (ty == "enabled") {
      std::optional<int64_t> bt; std::optional<bool> it;
      if (auto e = want_int(th->get("budget_tokens"), "budget_tokens", bt)) return e;
      if (auto e = want_str(th->get("type"), "type", t2)) return e;
      if (auto e = want_bool(th->get("includeThoughts"), "includeThoughts", it)) return e;
      r.thinking = ChatReq::ThEnabled; r.thinking_budget = bt.value_or(0); r.thinking_include_thoughts = it.value_or(false);
    } else if (ty == "disabled") { if (auto e = want_str(th->get("type"), "type", t2)) return e; r.thinking = ChatReq::ThDisabled; }
    else if (ty == "adaptive") { if (auto e = want_str(th->get("type"), "type", t2)) return e; r.thinking = ChatReq::ThAdaptive; }
    else return bad("invalid thinking union type: " + ty);
  }
  return {};
}

}  // namespace oracle
