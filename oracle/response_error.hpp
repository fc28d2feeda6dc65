// TEST INFRASTRUCTURE (oracle): CPU restatement of Translator.ResponseError for the chat-completion translators of three backends.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it; the product never does.
//
//   AWS Bedrock     openAIToAWSBedrockTranslatorV1ChatCompletion.ResponseError   internal/translator/openai_awsbedrock.go:643-689
//   GCP Anthropic   openAIToGCPAnthropicTranslatorV1ChatCompletion.ResponseError internal/translator/openai_gcpanthropic.go:103-153
//   GCP Vertex AI   convertGCPVertexAIErrorToOpenAI                              internal/translator/openai_gcpvertexai.go:602-648
//   openai.Error / openai.ErrorType (field order, omitempty)                      internal/apischema/openai/openai.go:1567-1587
//
// Output: {"type":"error","error":{"type":T,"code":"<:status>","message":M}} — "message" omitted when empty (omitempty), "code" always
// present (a non-nil *string).  Pinned byte for byte by the data-plane goldens "aws-bedrock / gcp-vertexai / gcp-anthropicai -
// /v1/chat/completions - error response" (tests/data-plane/testupstream_test.go).  Parity unpinned: Vertex `details` (json.RawMessage
// echoed inside the message: DECLINED here), messages with bytes the encoder would escape beyond enc_str's rules.
#pragma once
#include "ojson.hpp"
#include "chat.hpp"

namespace oracle {

enum { ERR_OPENAI = 2, ERR_MESSAGES_AWS_BEDROCK = 6, ERR_MESSAGES_OPENAI = 0, ERR_AWS_BEDROCK = 1, ERR_GCP_VERTEX = 3, ERR_GCP_ANTHROPIC = 4 };

inline void error_json(std::string& out, const std::string& type, const std::string& code, const std::string& message) {
  out = "{\"type\":\"error\",\"error\":{\"type\":"; oj::enc_str(out, type);
  out += ",\"code\":"; oj::enc_str(out, code);
  if (!message.empty()) { out += ",\"message\":"; oj::enc_str(out, message); }
  out += "}}";
}

inline bool str_field(const Value& o, const char* key, std::string& dst, bool& type_err) {   // last occurrence decides; null leaves the zero value
  if (const Value* v = o.get(key)) { if (v->is_null()) return true; if (!v->is_str()) { type_err = true; return false; } dst = v->s; }
  return true;
}

// json_content_type: the upstream's content-type contains "application/json" (Bedrock and GCP Anthropic look at it; Vertex does not)
// /v1/messages served by an OpenAI-schema backend: openai.Error -> anthropic.ErrorResponse
// (anthropicToOpenAIV1ChatCompletionTranslator.ResponseError, internal/translator/anthropic_openai.go:187-253; ErrorResponse /
// ErrorResponseMessage field order internal/apischema/anthropic/anthropic.go:1751-1763).  Pinned by the data-plane golden
// "anthropic-openai - … - OpenAI JSON error translated to Anthropic error" (exact text).
inline const char* anthropic_error_type_of_status(const std::string& code) {
  if (code == "400") return "invalid_request_error";
  if (code == "401") return "authentication_error";
  if (code == "403") return "permission_error";
  if (code == "404") return "not_found_error";
  if (code == "413") return "request_too_large";
  if (code == "429") return "rate_limit_error";
  if (code == "500") return "internal_server_error";
  if (code == "503") return "service_unavailable_error";
  if (code == "529") return "overloaded_error";
  return "internal_server_error";
}
inline void anthropic_error_json(std::string& out, const std::string& type, const std::string& message) {
  out = "{\"error\":{\"message\":"; oj::enc_str(out, message); out += ",\"type\":"; oj::enc_str(out, type); out += "},\"request_id\":\"\",\"type\":\"error\"}";
}
inline Status messages_openai_error(std::string_view body, const std::string& status_code, bool json_content_type, std::string& out) {
  if (!json_content_type) { anthropic_error_json(out, anthropic_error_type_of_status(status_code), std::string(body)); return OK; }
  Value v; { oj::Parser ps(body.data(), body.size()); ps.ws(); if (!ps.value(v)) return INTERNAL; }   // "failed to unmarshal OpenAI error body"
  bool te = false; std::string type, msg, dummy;
  if (v.is_obj()) {
    str_field(v, "event_id", dummy, te); str_field(v, "type", dummy, te);
    if (const Value* e = v.get("error")) {
      if (e->is_obj()) { str_field(*e, "type", type, te); str_field(*e, "code", dummy, te); str_field(*e, "message", msg, te); str_field(*e, "param", dummy, te); str_field(*e, "event_id", dummy, te); }
      else if (!e->is_null()) te = true;
    }
  } else if (!v.is_null()) te = true;
  if (te) return INTERNAL;
  anthropic_error_json(out, type, msg);
  return OK;
}

// /v1/messages served by AWS Bedrock Converse (anthropicToAWSBedrockTranslator.ResponseError, internal/translator/anthropic_awsbedrock.go:738-794):
// the message is BedrockException.message (JSON content type) or the raw body; the type follows the status through a table without
// 413 / 529.  Content pinned by anthropic_awsbedrock_test.go:563-638 (struct compares); the byte layout is the golden-pinned ErrorResponse's.
inline Status messages_bedrock_error(std::string_view body, const std::string& status_code, bool json_content_type, std::string& out) {
  std::string type = anthropic_error_type_of_status(status_code);
  if (status_code == "413" || status_code == "529") type = "internal_server_error";
  if (!json_content_type) { anthropic_error_json(out, type, std::string(body)); return OK; }
  Value v; { oj::Parser ps(body.data(), body.size()); ps.ws(); if (!ps.value(v)) return INTERNAL; }   // "failed to unmarshal error body"
  bool te = false; std::string msg, dummy;
  if (v.is_obj()) { str_field(v, "code", dummy, te); str_field(v, "type", dummy, te); str_field(v, "message", msg, te); }
  else if (!v.is_null()) te = true;
  if (te) return INTERNAL;
  anthropic_error_json(out, type, msg);
  return OK;
}

inline Status response_error(int kind, std::string_view body, const std::string& status_code, const std::string& aws_error_type, bool json_content_type, std::string& out) {
  out.clear();
  if (kind == ERR_MESSAGES_OPENAI) return messages_openai_error(body, status_code, json_content_type, out);
  if (kind == ERR_MESSAGES_AWS_BEDROCK) return messages_bedrock_error(body, status_code, json_content_type, out);
  if (kind == ERR_OPENAI) {   // convertErrorOpenAIToOpenAIError (openai_openai.go:94-120): only a non-JSON body is rewritten
    if (json_content_type) return DECLINED;   // forwarded untouched: the shim does not call
    error_json(out, "OpenAIBackendError", status_code, std::string(body));
    return OK;
  }
  if (kind == ERR_GCP_VERTEX) {
    Value v; std::string err; std::string status, msg; bool ok = oj::parse(body, v, err), te = false;
    if (ok) {
      if (v.is_obj()) {
        if (const Value* e = v.get("error")) {
          if (e->is_obj()) {
            if (e->get("details")) return DECLINED;                         // RawMessage echoed into the message: not restated
            if (const Value* c = e->get("code")) { if (!c->is_null()) { if (!c->is_num()) te = true; else for (char ch : c->s) if (ch == '.' || ch == 'e' || ch == 'E') te = true; } }
            str_field(*e, "message", msg, te); str_field(*e, "status", status, te);
          } else if (!e->is_null()) te = true;
        }
      } else if (!v.is_null()) te = true;
    }
    if (ok && !te) { error_json(out, status, status_code, msg); return OK; }
    error_json(out, "GCPVertexAIBackendError", status_code, std::string(body));   // not the expected JSON: the raw body is the message
    return OK;
  }
  if (!json_content_type) {
    error_json(out, kind == ERR_AWS_BEDROCK ? "AWSBedrockBackendError" : "GCPBackendError", status_code, std::string(body));
    return OK;
  }
  // json.NewDecoder(body).Decode: the first JSON value
  Value v; { oj::Parser ps(body.data(), body.size()); ps.ws(); if (!ps.value(v)) return INTERNAL; }
  bool te = false; std::string type, msg;
  if (kind == ERR_AWS_BEDROCK) {
    if (v.is_obj()) { std::string dummy; str_field(v, "code", dummy, te); str_field(v, "type", dummy, te); str_field(v, "message", msg, te); }
    else if (!v.is_null()) te = true;
    if (te) return INTERNAL;
    error_json(out, aws_error_type, status_code, msg);
    return OK;
  }
  if (v.is_obj()) {
    std::string dummy; str_field(v, "request_id", dummy, te); str_field(v, "type", dummy, te);
    if (const Value* e = v.get("error")) { if (e->is_obj()) { str_field(*e, "message", msg, te); str_field(*e, "type", type, te); } else if (!e->is_null()) te = true; }
  } else if (!v.is_null()) te = true;
  if (te) return INTERNAL;
  error_json(out, type, status_code, msg);
  return OK;
}

}  // namespace oracle
