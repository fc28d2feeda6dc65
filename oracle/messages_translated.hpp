// TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's /v1/messages request path for the two backends that need a
// full re-map of the request (T5, second part).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
//
//   MessagesEndpointSpec.ParseBody                          internal/endpointspec/endpointspec.go:320-336
//   anthropic.MessagesRequest, content / tool unions        internal/apischema/anthropic/anthropic.go:26-140,214-232,285-358,392-452,1021-1041,1097-1162,1186-1268,1369-1390
//   anthropicToOpenAIV1ChatCompletionTranslator.RequestBody internal/translator/anthropic_openai.go:55-93
//   buildOpenAIChatCompletionRequest and helpers            internal/translator/openai_helper.go:27-261
//   openai.ChatCompletionRequest field order                internal/apischema/openai/openai.go:947-1131 (messages, model, max_completion_tokens, stream,
//                                                            stream_options, temperature, top_p, tools, tool_choice), message params :485-614
//   anthropicToAWSBedrockTranslator.RequestBody             internal/translator/anthropic_awsbedrock.go:52-163,176-368
//   awsbedrock.ConverseInput field order                    internal/apischema/awsbedrock/awsbedrock.go:41-70,126-200,310-367,516-605
//
// Subset: text / tool_use / tool_result content blocks, string or text-block system prompt, sampling parameters, stop sequences,
// custom tools, tool_choice.  Everything else (images, documents, thinking, built-in tools, service_tier, ...) is DECLINED, so
// nothing outside the restated rules is given a verdict.
// Pinned by the data-plane goldens tests/data-plane/testupstream_test.go: the six "anthropic-openai" cases (exact expRequestBody /
// expPath) and "aws-bedrock - /anthropic/v1/messages".  Parity unpinned: the position of temperature / top_p / tool_choice in the
// OpenAI body (Go struct order, same encoder as the pinned members); "stop" is appended by sjson, which marshals the []string with
// encoding/json (HTML-escaping <, >, & and U+2028/9): stop sequences outside printable ASCII-without-<>& are DECLINED.
#pragma once
#include "messages.hpp"

namespace oracle {

enum { SCHEMA_MSG_OPENAI = 0, SCHEMA_MSG_AWS_BEDROCK = 1 };

namespace msgx {

struct Block {
  enum Kind { Text, ToolUse, ToolResult } kind = Text;
  std::string text;                        // Text
  std::string id, name; const Value* input = nullptr;   // ToolUse (input: object or absent/null)
  std::string tool_use_id; bool has_content = false; std::string content_text; std::vector<std::string> content_items; bool content_is_array = false; bool is_error = false;  // ToolResult
};
struct Msg { std::string role; bool is_string = false; std::string text; std::vector<Block> blocks; };
struct Tool { std::string name, description; bool has_type = false; std::string schema_type; const Value* properties = nullptr; std::vector<std::string> required; };
struct Req {
  std::string model; bool stream = false;
  int64_t max_tokens = 0;
  const Value* temperature = nullptr; const Value* top_p = nullptr; bool has_top_k = false; int64_t top_k = 0;
  std::vector<std::string> stop;
  bool has_system = false; bool system_is_string = false; std::string system_text; std::vector<std::string> system_blocks;
  std::vector<Msg> messages; bool messages_null = true;
  std::vector<Tool> tools;
  enum TC { None_, Auto, Any, ToolNamed, NoneChoice } tc = None_; std::string tc_name;
};

inline bool cache_ok(const Value& c) { return c.is_obj() && c.obj.size() == 1 && c.obj[0].first == "type" && c.obj[0].second.is_str() && c.obj[0].second.s == "ephemeral"; }

// float64 → int64(): the subset takes non-negative integer literals, optionally followed by ".0…0", of at most 15 digits
inline bool plain_int(const std::string& lit, int64_t& out) {
  size_t i = 0; while (i < lit.size() && lit[i] >= '0' && lit[i] <= '9') i++;
  if (i == 0 || i > 15) return false;
  if (i > 1 && lit[0] == '0') return false;
  if (i < lit.size()) { if (lit[i] != '.' || i + 1 == lit.size()) return false; for (size_t k = i + 1; k < lit.size(); k++) if (lit[k] != '0') return false; }
  out = strtoll(lit.substr(0, i).c_str(), nullptr, 10); return true;
}

// 0 accepted, 1 DECLINED
inline int parse_block(const Value& b, Block& out) {
  if (!b.is_obj()) return 1;
  const Value* ty = nullptr;
  for (auto& kv : b.obj) if (kv.first == "type") { if (ty) return 1; ty = &kv.second; }
  if (!ty || !ty->is_str()) return 1;
  if (ty->s == "text") { if (msgs::text_block(b)) return 1; out.kind = Block::Text; out.text = b.get("text")->s; return 0; }
  std::vector<std::string> seen;
  auto dup = [&](const std::string& k) { for (auto& s : seen) if (s == k) return true; seen.push_back(k); return false; };
  if (ty->s == "tool_use") {
    out.kind = Block::ToolUse; bool has_id = false, has_name = false;
    for (auto& kv : b.obj) {
      if (dup(kv.first)) return 1;
      if (kv.first == "type") continue;
      else if (kv.first == "id") { if (!kv.second.is_str()) return 1; out.id = kv.second.s; has_id = true; }
      else if (kv.first == "name") { if (!kv.second.is_str()) return 1; out.name = kv.second.s; has_name = true; }
      else if (kv.first == "input") { if (kv.second.is_null()) continue; if (!kv.second.is_obj()) return 1; out.input = &kv.second; }
      else if (kv.first == "cache_control") { if (!cache_ok(kv.second)) return 1; }
      else return 1;
    }
    return (has_id && has_name) ? 0 : 1;
  }
  if (ty->s == "tool_result") {
    out.kind = Block::ToolResult; bool has_id = false;
    for (auto& kv : b.obj) {
      if (dup(kv.first)) return 1;
      if (kv.first == "type") continue;
      else if (kv.first == "tool_use_id") { if (!kv.second.is_str()) return 1; out.tool_use_id = kv.second.s; has_id = true; }
      else if (kv.first == "is_error") { if (!kv.second.is_bool()) return 1; out.is_error = kv.second.t == oj::T::True; }
      else if (kv.first == "cache_control") { if (!cache_ok(kv.second)) return 1; }
      else if (kv.first == "content") {
        const Value& c = kv.second;
        if (c.is_null()) continue;
        out.has_content = true;
        if (c.is_str()) { out.content_text = c.s; continue; }
        if (!c.is_arr()) return 1;
        out.content_is_array = true;
        for (auto& it : c.arr) { if (msgs::text_block(it)) return 1; out.content_items.push_back(it.get("text")->s); }
      }
      else return 1;
    }
    return has_id ? 0 : 1;
  }
  return 1;
}

// 0 accepted, 1 DECLINED, 2 a definite decode error (400); member order decides, as in msgs::check_request
inline int parse_request(const Value& root, Req& r) {
  if (root.is_null()) return 0;
  if (!root.is_obj()) return 2;
  std::vector<std::string> seen;
  for (auto& kv : root.obj) {
    for (auto& s : seen) if (s == kv.first) return 1;
    seen.push_back(kv.first);
    const std::string& k = kv.first; const Value& v = kv.second;
    if (k == "model") { if (v.is_null()) continue; if (!v.is_str()) return 2; r.model = v.s; }
    else if (k == "max_tokens") { if (v.is_null()) continue; if (!v.is_num()) return 2; if (!plain_int(v.s, r.max_tokens)) return 1; }
    else if (k == "temperature") { if (v.is_null()) continue; if (!v.is_num()) return 2; r.temperature = &v; }
    else if (k == "top_p") { if (v.is_null()) continue; if (!v.is_num()) return 2; r.top_p = &v; }
    else if (k == "top_k") { if (v.is_null()) continue; if (!v.is_num()) return 2; if (!oj::num_to_i64(v.s, r.top_k) || r.top_k < 0 || v.s.size() > 9) return 1; r.has_top_k = true; }
    else if (k == "stream") { if (v.is_null()) continue; if (!v.is_bool()) return 2; r.stream = v.t == oj::T::True; }
    else if (k == "stop_sequences") { if (v.is_null()) continue; if (!v.is_arr()) return 2; for (auto& e : v.arr) { if (!e.is_str()) return 1; r.stop.push_back(e.s); } }
    else if (k == "metadata") {
      if (v.is_null()) continue; if (!v.is_obj()) return 2;
      if (v.obj.size() > 1) return 1;
      if (v.obj.size() == 1 && (v.obj[0].first != "user_id" || !(v.obj[0].second.is_str() || v.obj[0].second.is_null()))) return 1;
    }
    else if (k == "system") {
      r.has_system = true;
      if (v.is_str()) { r.system_is_string = true; r.system_text = v.s; continue; }
      if (!v.is_arr()) return 1;
      for (auto& b : v.arr) { if (msgs::text_block(b)) return 1; r.system_blocks.push_back(b.get("text")->s); }
    }
    else if (k == "messages") {
      if (v.is_null()) continue; if (!v.is_arr()) return 2;
      r.messages_null = false;
      for (auto& m : v.arr) {
        if (!m.is_obj()) return 1;
        Msg g; bool has_role = false, has_content = false;
        for (auto& mk : m.obj) {
          if (mk.first == "role") { if (has_role) return 1; has_role = true; if (!mk.second.is_str()) return 1; g.role = mk.second.s; }
          else if (mk.first == "content") {
            if (has_content) return 1; has_content = true;
            if (mk.second.is_str()) { g.is_string = true; g.text = mk.second.s; continue; }
            if (!mk.second.is_arr()) return 1;
            for (auto& b : mk.second.arr) { Block bl; if (parse_block(b, bl)) return 1; g.blocks.push_back(std::move(bl)); }
          } else return 1;
        }
        if (!has_role || !has_content) return 1;
        r.messages.push_back(std::move(g));
      }
    }
    else if (k == "tools") {
      if (v.is_null()) continue; if (!v.is_arr()) return 2;
      for (auto& t : v.arr) {
        if (!t.is_obj()) return 1;
        Tool tl; bool has_name = false, has_schema = false; std::vector<std::string> ts;
        for (auto& tk : t.obj) {
          for (auto& s : ts) if (s == tk.first) return 1;
          ts.push_back(tk.first);
          const Value& tv = tk.second;
          if (tk.first == "type") { if (!tv.is_str() || !(tv.s.empty() || tv.s == "custom")) return 1; }
          else if (tk.first == "name") { if (!tv.is_str()) return 1; tl.name = tv.s; has_name = true; }
          else if (tk.first == "description") { if (!tv.is_str()) return 1; tl.description = tv.s; }
          else if (tk.first == "cache_control") { if (!cache_ok(tv)) return 1; }
          else if (tk.first == "input_schema") {
            if (!tv.is_obj()) return 1;
            has_schema = true; std::vector<std::string> ss;
            for (auto& sk : tv.obj) {
              for (auto& s : ss) if (s == sk.first) return 1;
              ss.push_back(sk.first);
              if (sk.first == "type") { if (!sk.second.is_str()) return 1; tl.has_type = true; tl.schema_type = sk.second.s; }
              else if (sk.first == "properties") { if (sk.second.is_null()) continue; if (!sk.second.is_obj()) return 1; tl.properties = &sk.second; }
              else if (sk.first == "required") { if (sk.second.is_null()) continue; if (!sk.second.is_arr()) return 1; for (auto& e : sk.second.arr) { if (!e.is_str()) return 1; tl.required.push_back(e.s); } }
              // any other member of the schema is dropped by the decoder (ToolInputSchema has three fields)
            }
          }
          else return 1;
        }
        if (!has_name || !has_schema) return 1;
        r.tools.push_back(std::move(tl));
      }
    }
    else if (k == "tool_choice") {
      if (v.is_null()) continue; if (!v.is_obj()) return 1;
      const Value* ty = nullptr; const Value* nm = nullptr; std::vector<std::string> cs;
      for (auto& ck : v.obj) {
        for (auto& s : cs) if (s == ck.first) return 1;
        cs.push_back(ck.first);
        if (ck.first == "type") ty = &ck.second;
        else if (ck.first == "name") nm = &ck.second;
        else if (ck.first == "disable_parallel_tool_use") { if (!(ck.second.is_bool() || ck.second.is_null())) return 1; }
        else return 1;
      }
      if (!ty || !ty->is_str()) return 1;
      if (ty->s == "auto") r.tc = Req::Auto; else if (ty->s == "any") r.tc = Req::Any; else if (ty->s == "none") r.tc = Req::NoneChoice;
      else if (ty->s == "tool") { if (!nm || !nm->is_str()) return 1; r.tc = Req::ToolNamed; r.tc_name = nm->s; }
      else return 1;
      if (ty->s != "tool" && nm) return 1;
    }
    else return 1;
  }
  return 0;
}

inline std::string content_text(const Msg& m) {   // anthropicContentToText, openai_helper.go:194-205
  if (m.is_string) return m.text;
  std::string s; for (auto& b : m.blocks) if (b.kind == Block::Text) s += b.text; return s;
}
inline std::string tool_result_text(const Block& b) {   // toolResultToText, :160-175
  if (!b.has_content) return "";
  if (!b.content_is_array) return b.content_text;
  std::string s; for (auto& t : b.content_items) s += t; return s;
}
// what encoding/json (sjson's marshaller for a []string) and sonic spell alike
inline bool plain_stop(const std::string& s) {
  for (unsigned char c : s) { if (c >= 0x7f || c == '<' || c == '>' || c == '&') return false; if (c < 0x20 && c != '\n' && c != '\r' && c != '\t') return false; }
  return true;
}
inline void schema_json(std::string& o, const Tool& t) {   // ToolInputSchema marshal: type, properties?, required?
  o += "{\"type\":"; oj::enc_str(o, t.schema_type);
  if (t.properties && !t.properties->obj.empty()) { o += ",\"properties\":"; oj::enc_any(o, *t.properties); }
  if (!t.required.empty()) { o += ",\"required\":["; for (size_t i = 0; i < t.required.size(); i++) { if (i) o.push_back(','); oj::enc_str(o, t.required[i]); } o.push_back(']'); }
  o.push_back('}');
}

inline void to_openai(const Req& r, const std::string& model, std::string& o) {
  o = "{\"messages\":";
  std::vector<std::string> ms;
  if (r.has_system) {
    std::string st = r.system_text;
    if (st.empty()) for (auto& b : r.system_blocks) st += b;
    if (!st.empty()) { std::string m = "{\"content\":"; oj::enc_str(m, st); m += ",\"role\":\"system\"}"; ms.push_back(m); }
  }
  for (auto& g : r.messages) {
    if (g.role == "user") {
      for (auto& b : g.blocks) if (b.kind == Block::ToolResult) {
        std::string m = "{\"content\":"; oj::enc_str(m, tool_result_text(b)); m += ",\"role\":\"tool\",\"tool_call_id\":"; oj::enc_str(m, b.tool_use_id); m.push_back('}'); ms.push_back(m);
      }
      const std::string t = content_text(g);
      if (!t.empty()) { std::string m = "{\"content\":"; oj::enc_str(m, t); m += ",\"role\":\"user\"}"; ms.push_back(m); }
    } else if (g.role == "assistant") {
      const std::string t = content_text(g);
      std::string m = "{\"role\":\"assistant\",\"content\":";
      if (t.empty()) m += "null"; else oj::enc_str(m, t);
      bool any = false;
      for (auto& b : g.blocks) if (b.kind == Block::ToolUse) {
        m += any ? "," : ",\"tool_calls\":["; any = true;
        std::string args; if (b.input) oj::enc_any(args, *b.input); else args = "null";
        m += "{\"id\":"; oj::enc_str(m, b.id); m += ",\"function\":{\"arguments\":"; oj::enc_str(m, args); m += ",\"name\":"; oj::enc_str(m, b.name); m += "},\"type\":\"function\"}";
      }
      if (any) m.push_back(']');
      m.push_back('}'); ms.push_back(m);
    }
  }
  if (ms.empty()) o += "null";
  else { o.push_back('['); for (size_t i = 0; i < ms.size(); i++) { if (i) o.push_back(','); o += ms[i]; } o.push_back(']'); }
  o += ",\"model\":"; oj::enc_str(o, model);
  o += ",\"max_completion_tokens\":" + std::to_string(r.max_tokens);
  if (r.stream) o += ",\"stream\":true,\"stream_options\":{\"include_usage\":true}";
  if (r.temperature) { o += ",\"temperature\":"; oj::enc_f64(o, strtod(r.temperature->s.c_str(), nullptr)); }
  if (r.top_p) { o += ",\"top_p\":"; oj::enc_f64(o, strtod(r.top_p->s.c_str(), nullptr)); }
  if (!r.tools.empty()) {
    o += ",\"tools\":[";
    for (size_t i = 0; i < r.tools.size(); i++) {
      const Tool& t = r.tools[i];
      if (i) o.push_back(',');
      o += "{\"type\":\"function\",\"function\":{\"name\":"; oj::enc_str(o, t.name);
      if (!t.description.empty()) { o += ",\"description\":"; oj::enc_str(o, t.description); }
      o += ",\"parameters\":"; schema_json(o, t); o += "}}";
    }
    o.push_back(']');
    switch (r.tc) {
      case Req::Auto: o += ",\"tool_choice\":\"auto\""; break;
      case Req::NoneChoice: o += ",\"tool_choice\":\"none\""; break;
      case Req::Any: o += ",\"tool_choice\":\"required\""; break;
      case Req::ToolNamed: o += ",\"tool_choice\":{\"type\":\"function\",\"function\":{\"name\":"; oj::enc_str(o, r.tc_name); o += "}}"; break;
      default: break;
    }
  }
  if (!r.stop.empty()) {   // sjson.SetBytesOptions(newBody, "stop", []string): appended as the last member
    o += ",\"stop\":["; for (size_t i = 0; i < r.stop.size(); i++) { if (i) o.push_back(','); oj::enc_str(o, r.stop[i]); } o.push_back(']');
  }
  o.push_back('}');
}

inline void bedrock_tool_result(std::string& o, const Block& b) {   // convertToolResultBlock, anthropic_awsbedrock.go:264-291
  o += "{\"toolResult\":{\"content\":";
  if (b.has_content && !b.content_is_array && !b.content_text.empty()) { o += "[{\"text\":"; oj::enc_str(o, b.content_text); o += "}]"; }
  else if (b.has_content && b.content_is_array && !b.content_items.empty()) {
    o.push_back('['); for (size_t i = 0; i < b.content_items.size(); i++) { if (i) o.push_back(','); o += "{\"text\":"; oj::enc_str(o, b.content_items[i]); o.push_back('}'); } o.push_back(']');
  } else o += "null";
  o += b.is_error ? ",\"status\":\"error\"" : ",\"status\":null";
  o += ",\"toolUseId\":"; oj::enc_str(o, b.tool_use_id); o += "}}";
}

// returns false with `bad_role` set when a message has a role other than user / assistant (422)
inline bool to_bedrock(const Req& r, std::string& o, std::string& bad_role) {
  o = "{";
  if (r.has_top_k) o += "\"additionalModelRequestFields\":{\"top_k\":" + std::to_string(r.top_k) + "},";
  o += "\"inferenceConfig\":{\"maxTokens\":" + std::to_string(r.max_tokens);
  if (!r.stop.empty()) { o += ",\"stopSequences\":["; for (size_t i = 0; i < r.stop.size(); i++) { if (i) o.push_back(','); oj::enc_str(o, r.stop[i]); } o.push_back(']'); }
  if (r.temperature) { o += ",\"temperature\":"; oj::enc_f64(o, strtod(r.temperature->s.c_str(), nullptr)); }
  if (r.top_p) { o += ",\"topP\":"; oj::enc_f64(o, strtod(r.top_p->s.c_str(), nullptr)); }
  o += "},\"messages\":[";
  bool first = true;
  for (auto& g : r.messages) {
    const bool user = g.role == "user";
    if (!user && g.role != "assistant") { bad_role = g.role; return false; }
    if (!first) o.push_back(','); first = false;
    o += "{\"content\":[";
    if (g.is_string ? !g.text.empty() : false) { o += "{\"text\":"; oj::enc_str(o, g.text); o.push_back('}'); }
    else {
      bool bf = true;
      for (auto& b : g.blocks) {
        std::string e;
        if (b.kind == Block::Text) { e = "{\"text\":"; oj::enc_str(e, b.text); e.push_back('}'); }
        else if (user && b.kind == Block::ToolResult) bedrock_tool_result(e, b);
        else if (!user && b.kind == Block::ToolUse) {
          e = "{\"toolUse\":{\"name\":"; oj::enc_str(e, b.name); e += ",\"input\":"; if (b.input) oj::enc_any(e, *b.input); else e += "null"; e += ",\"toolUseId\":"; oj::enc_str(e, b.id); e += "}}";
        } else continue;
        if (!bf) o.push_back(','); bf = false; o += e;
      }
    }
    o += "],\"role\":"; o += user ? "\"user\"}" : "\"assistant\"}";
  }
  o.push_back(']');
  if (r.has_system) {   // convertSystemPrompt :332-345; an empty slice is dropped by omitempty
    if (!r.system_text.empty()) { o += ",\"system\":[{\"text\":"; oj::enc_str(o, r.system_text); o += "}]"; }
    else if (!r.system_blocks.empty()) { o += ",\"system\":["; for (size_t i = 0; i < r.system_blocks.size(); i++) { if (i) o.push_back(','); o += "{\"text\":"; oj::enc_str(o, r.system_blocks[i]); o.push_back('}'); } o.push_back(']'); }
  }
  if (!r.tools.empty()) {   // convertTools :347-395; ToolConfiguration{tools, toolChoice?}, ToolSpecification{description?, inputSchema, name}
    o += ",\"toolConfig\":{";
    switch (r.tc) {
      case Req::Auto: o += "\"toolChoice\":{\"auto\":{}},"; break;
      case Req::Any: o += "\"toolChoice\":{\"any\":{}},"; break;
      case Req::ToolNamed: o += "\"toolChoice\":{\"tool\":{\"name\":"; oj::enc_str(o, r.tc_name); o += "}},"; break;
      default: break;
    }
    o += "\"tools\":[";
    for (size_t i = 0; i < r.tools.size(); i++) {
      const Tool& t = r.tools[i];
      if (i) o.push_back(',');
      o += "{\"toolSpec\":{";
      if (!t.description.empty()) { o += "\"description\":"; oj::enc_str(o, t.description); o.push_back(','); }
      o += "\"inputSchema\":{\"json\":"; schema_json(o, t); o += "},\"name\":"; oj::enc_str(o, t.name); o += "}}";
    }
    o += "]}";
  }
  o.push_back('}');
  return true;
}

}  // namespace msgx

// schema = SCHEMA_MESSAGES | {SCHEMA_MSG_OPENAI, SCHEMA_MSG_AWS_BEDROCK}; prefix = the OpenAI backend's path prefix ("v1")
inline TranslateResult messages_translate_full(int schema, std::string_view body, const std::string& model_override, const std::string& prefix) {
  TranslateResult res;
  Value root; std::string perr;
  if (!oj::parse(body, root, perr)) { res.err = bad("malformed request: failed to parse JSON for /v1/messages: " + perr); return res; }
  msgx::Req r;
  const int c = msgx::parse_request(root, r);
  if (c == 2) { res.err = bad("malformed request: failed to parse JSON for /v1/messages"); return res; }
  if (c == 1) { res.err = Error{DECLINED, "outside the restated subset of anthropic.MessagesRequest"}; return res; }
  if (r.model.empty()) { res.err = invalid("invalid request body: model field is required"); return res; }
  res.model = r.model; res.stream = r.stream;
  res.request_model = model_override.empty() ? r.model : model_override;
  const int base = schema & 15;
  std::string nb;
  if (base == SCHEMA_MSG_OPENAI) {
    for (auto& s : r.stop) if (!msgx::plain_stop(s)) { res.err = Error{DECLINED, "stop sequence outside the spelling encoding/json and sonic share"}; return res; }
    msgx::to_openai(r, res.request_model, nb);
    std::string p = "/"; if (!prefix.empty()) p += prefix + "/"; p += "chat/completions";
    res.headers.push_back({":path", p});
  } else if (base == SCHEMA_MSG_AWS_BEDROCK) {
    std::string bad_role;
    if (!msgx::to_bedrock(r, nb, bad_role)) { res.err = invalid("invalid request body: unexpected role: " + bad_role); return res; }
    res.headers.push_back({":path", "/model/" + path_escape(res.request_model) + (r.stream ? "/converse-stream" : "/converse")});
  } else { res.err = Error{DECLINED, "schema not restated"}; return res; }
  res.body_kind = BYTES; res.body = nb;
  res.headers.push_back({"content-length", std::to_string(nb.size())});
  return res;
}

}  // namespace oracle
