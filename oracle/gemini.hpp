// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of OpenAI → GCP Vertex AI Gemini, request direction:
//   RequestBody / openAIMessageToGeminiMessage     internal/translator/openai_gcpvertexai.go:93-126,512-580
//   openAIMessagesToGeminiContents                 internal/translator/gemini_helper.go:52-122
//   developer / user / assistant parts             internal/translator/gemini_helper.go:144-165,168-251,257-339
//   openAIToolsToGeminiTools                       internal/translator/gemini_helper.go:401-474
//   openAIReqToGeminiGenerationConfig              internal/translator/gemini_helper.go:609-734
//   responseJSONSchemaAvailable                    internal/translator/gemini_helper.go:554-556
//   buildGCPModelPathSuffix                        internal/translator/gemini_helper.go:1004-1011
//   gcp.GenerateContentRequest                     internal/apischema/gcp/gcp.go:14-42
// The inner types are google.golang.org/genai v1.55.0 (not in tree).  Pinned byte for byte by
// tests/data-plane/testupstream_test.go:313,327,341,510 (contents / tools / generation_config / system_instruction order, `null`
// for empty contents and tools, Content = {parts, role}, Part = {text}, FunctionDeclaration = {description, name, parameters},
// Schema keys in alphabetical order).  Parity unpinned ⇒ DECLINED: tool calls and tool messages (FunctionCall / FunctionResponse
// part layouts), images / audio / files, tool_choice, thinking, reasoning_effort, response_format / guided_*, vendor fields, JSON
// schemas beyond {type, description, properties, required, items, enum}, empty text parts.  generation_config members are emitted
// in genai.GenerationConfig's field order as published for v1.55.0 (candidateCount, frequencyPenalty, logprobs, maxOutputTokens,
// presencePenalty, responseLogprobs, seed, stopSequences, temperature, topP); the reference pins them only with JSONEq
// (internal/translator/openai_gcpvertexai_test.go:69-813), so the oracle's golden test compares those semantically.
#pragma once
#include "translate.hpp"

namespace oracle {
namespace gemini {
using anthropic::unpinned;

inline bool schema_subset(const Value& v, int depth = 0) {
  if (depth > 16 || !v.is_obj()) return false;
  for (auto& kv : v.obj) {
    const std::string_view k = kv.first; const Value& x = kv.second;
    if (k == "type" || k == "description") { if (!x.is_str()) return false; }
    else if (k == "properties") { if (!x.is_obj()) return false; for (auto& p : x.obj) if (!schema_subset(p.second, depth + 1)) return false; }
    else if (k == "items") { if (!schema_subset(x, depth + 1)) return false; }
    else if (k == "required" || k == "enum") { if (!x.is_arr()) return false; for (auto& e : x.arr) if (!e.is_str()) return false; }
    else return false;
  }
  return true;
}

inline void text_part(std::string& o, bool& first, std::string_view t) { if (!first) o.push_back(','); first = false; o += "{\"text\":"; oj::enc_str(o, t); o += "}"; }

inline TranslateResult request_body(const ChatReq& r, const std::string& model_override) {
  TranslateResult res; res.stream = r.stream; res.model = r.model;
  res.request_model = model_override.empty() ? r.model : model_override;
  auto fail = [&](Error e) { res.err = e; return res; };
  const std::string& rm = res.request_model;
  std::string contents, sys, pending; bool cfirst = true, sfirst = true, pfirst = true;
  auto flush_user = [&] { if (!pending.empty()) { if (!cfirst) contents.push_back(','); cfirst = false; contents += "{\"parts\":[" + pending + "],\"role\":\"user\"}"; pending.clear(); pfirst = true; } };
  for (const Message& m : r.messages) {
    switch (m.role) {
      case Message::System: case Message::Developer:
        if (m.ck == Message::String) { if (!m.content_str.empty()) text_part(sys, sfirst, m.content_str); }
        else if (m.ck == Message::TextParts) { for (auto& p : m.text_parts) if (!p.text.empty()) text_part(sys, sfirst, p.text); }
        else return fail(invalid("message 'content' must be a string or an array"));
        break;
      case Message::User:
        if (m.ck == Message::String) { if (!m.content_str.empty()) text_part(pending, pfirst, m.content_str); }
        else if (m.ck == Message::UserParts) {
          for (auto& p : m.user_parts) {
            if (p.k != UserPart::Text) return fail(unpinned("non-text user content part"));
            if (p.text.text.empty()) return fail(unpinned("empty text part"));
            text_part(pending, pfirst, p.text.text);
          }
        } else return fail(invalid("message 'content' must be a string or an array"));
        break;
      case Message::Tool: return fail(unpinned("tool message (FunctionResponse part)"));
      case Message::Assistant: {
        flush_user();
        if (!m.tool_calls.empty()) return fail(unpinned("assistant tool calls (FunctionCall part)"));
        std::string parts; bool f = true;
        if (m.ck == Message::String) { if (!m.content_str.empty()) text_part(parts, f, m.content_str); }
        else if (m.ck == Message::AsstParts || m.ck == Message::AsstSingle) {
          if (m.ck == Message::AsstSingle) return fail(unpinned("assistant content object"));
          for (auto& p : m.asst_parts) {
            if (p.type == "text") { if (p.text && !p.text->empty()) text_part(parts, f, *p.text); }
            else if (p.type == "refusal") {}
            else if (p.type == "thinking") return fail(unpinned("thought part"));
            else return fail(invalid("unsupported content type in assistant message"));
          }
        } else if (m.ck != Message::None) return fail(invalid("message 'content' must be a string or an array"));
        if (!cfirst) contents.push_back(','); cfirst = false;
        contents += parts.empty() ? std::string("{\"role\":\"model\"}") : "{\"parts\":[" + parts + "],\"role\":\"model\"}";
        break;
      }
    }
  }
  flush_user();
  // tools
  std::string decls; bool dfirst = true;
  const bool json_schema = rm.find("gemini") != std::string::npos && (rm.find("2.5") != std::string::npos || rm.find("3") != std::string::npos);
  for (const ToolDef& t : r.tools) {
    if (t.type != "function") return fail(unpinned("non-function tool"));
    if (!t.has_function) continue;
    if (!dfirst) decls.push_back(','); dfirst = false;
    decls += "{";
    bool f = true;
    if (!t.description.empty()) { decls += "\"description\":"; oj::enc_str(decls, t.description); f = false; }
    if (!t.name.empty()) { if (!f) decls.push_back(','); decls += "\"name\":"; oj::enc_str(decls, t.name); f = false; }
    if (t.parameters && !t.parameters->is_null()) {
      if (json_schema) {
        if (!(t.parameters->is_obj() && t.parameters->obj.empty())) { if (!f) decls.push_back(','); decls += "\"parametersJsonSchema\":"; oj::enc_any(decls, *t.parameters); }
      } else {
        if (!t.parameters->is_obj()) return fail(invalid("tool parameters must be a JSON object"));
        if (!t.parameters->obj.empty()) {
          if (!schema_subset(*t.parameters)) return fail(unpinned("JSON schema outside the restated subset"));
          if (!f) decls.push_back(','); decls += "\"parameters\":"; oj::enc_any(decls, *t.parameters);
        }
      }
    }
    decls += "}";
  }
  if (r.tool_choice != ChatReq::TCNone) return fail(unpinned("tool_choice"));
  if (r.thinking != ChatReq::ThNone) return fail(unpinned("thinking"));
  if (!r.reasoning_effort.empty()) return fail(unpinned("reasoning_effort"));
  if (r.response_format || r.guided_json || r.guided_choice || !r.guided_regex.empty()) return fail(unpinned("response format"));
  if (r.generation_config || r.safety_settings) return fail(unpinned("vendor fields"));
  // generation config (genai.GenerationConfig field order)
  std::string gc; bool gf = true;
  auto sep = [&] { if (!gf) gc.push_back(','); gf = false; };
  auto i32 = [&](const char* k, int64_t v) { sep(); gc += std::string("\"") + k + "\":" + std::to_string((int32_t)v); };
  if (r.n && (int32_t)*r.n != 0) i32("candidateCount", *r.n);
  if (r.frequency_penalty) { sep(); gc += "\"frequencyPenalty\":"; oj::enc_f32(gc, *r.frequency_penalty); }
  if (r.top_logprobs) i32("logprobs", *r.top_logprobs);
  { std::optional<int64_t> mt = r.max_completion_tokens ? r.max_completion_tokens : r.max_tokens; if (mt && (int32_t)*mt != 0) i32("maxOutputTokens", *mt); }
  if (r.presence_penalty) { sep(); gc += "\"presencePenalty\":"; oj::enc_f32(gc, *r.presence_penalty); }
  if (r.logprobs && *r.logprobs) { sep(); gc += "\"responseLogprobs\":true"; }
  if (r.seed) i32("seed", *r.seed);
  if (r.stop_is_string) { sep(); gc += "\"stopSequences\":["; oj::enc_str(gc, r.stop_string); gc += "]"; }
  else if (r.stop_array && !r.stop_array->empty()) { sep(); gc += "\"stopSequences\":["; for (size_t k = 0; k < r.stop_array->size(); k++) { if (k) gc.push_back(','); oj::enc_str(gc, (*r.stop_array)[k]); } gc += "]"; }
  if (r.temperature) { sep(); gc += "\"temperature\":"; oj::enc_f32(gc, (float)*r.temperature); }
  if (r.top_p) { sep(); gc += "\"topP\":"; oj::enc_f32(gc, (float)*r.top_p); }
  std::string& o = res.body;
  o = "{\"contents\":"; o += contents.empty() ? "null" : "[" + contents + "]";
  o += ",\"tools\":"; o += decls.empty() ? "null" : "[{\"functionDeclarations\":[" + decls + "]}]";
  o += ",\"generation_config\":{" + gc + "}";
  if (!sys.empty()) o += ",\"system_instruction\":{\"parts\":[" + sys + "]}";
  o += "}";
  res.body_kind = BYTES;
  std::string path = "publishers/google/models/" + rm + (r.stream ? ":streamGenerateContent?alt=sse" : ":generateContent");
  res.headers.push_back({":path", path});
  res.headers.push_back({"content-length", std::to_string(o.size())});
  return res;
}

}  // namespace gemini
}  // namespace oracle
