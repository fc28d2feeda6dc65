// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the RESPONSE direction of /v1/messages served by an OpenAI-schema backend (T5):
//   ResponseBody, buffered: ChatCompletionResponse -> anthropic.MessagesResponse   internal/translator/anthropic_openai.go:112-152,
//                                                                                  openai_helper.go:263-338
//   ResponseBody, stream: OpenAI SSE chunks -> Anthropic SSE events                 internal/translator/anthropic_openai.go:154-185,
//                                                                                  openai_helper.go:340-766
//   ExtractTokenUsageFromExplicitCaching                                           internal/metrics/metrics.go:292-307
//   anthropic.MessagesResponse / MessagesContentBlock / Usage field order          internal/apischema/anthropic/anthropic.go:1413-1442
// Pinned by the data-plane goldens tests/data-plane/testupstream_test.go ("anthropic-openai - …": three buffered bodies, two streams),
// whose expected texts are compared byte for byte here.  Unpinned: the byte layout of a tool_use `input` with more than one key
// (Go map order; chosen: sorted, as encoding/json and sonic's SortMapKeys do) and float formatting of non-integral numbers in it.
#pragma once
#include <map>

#include "anthropic_stream.hpp"
#include "stream.hpp"

namespace oracle {

inline const char* openai_finish_to_anthropic(const std::string& r) {  // openai_helper.go:324-337
  if (r == "stop") return "end_turn";
  if (r == "length") return "max_tokens";
  if (r == "tool_calls") return "tool_use";
  if (r == "content_filter") return "refusal";
  return "end_turn";
}

// metrics.ExtractTokenUsageFromExplicitCaching (metrics.go:292-307); has_cache: both cache pointers non-nil (zero)
inline TokenUsage explicit_caching(int64_t in, int64_t out, bool has_cache) {
  TokenUsage u;
  if (has_cache) { u.cached = 0; u.cache_creation = 0; u.mask |= TokenUsage::CACHED | TokenUsage::CACHE_CREATION; }
  u.input = (uint32_t)in; u.output = (uint32_t)out; u.total = (uint32_t)(in + out);
  u.mask |= TokenUsage::IN | TokenUsage::OUT | TokenUsage::TOTAL;
  return u;
}

// ---------------------------------------------------------------- stream
struct OpenAIToAnthropicStream {  // openAIStreamToAnthropicState, openai_helper.go:436-457
  std::string buffer;
  bool started = false, open = false, closing = false;
  std::string id, model, stop_reason, request_model;
  int64_t in_tok = 0, out_tok = 0;
  TokenUsage usage;
  int64_t block_index = 0;
  std::map<int64_t, int64_t> tools;  // OpenAI tool_call index -> Anthropic block index
};

inline void an_sse(std::string& out, const char* type, const std::string& data) {  // appendAnthropicSSEEvent :759-766
  out += "event: "; out += type; out += "\ndata: "; out += data; out += "\n\n";
}

inline void o2a_closing(OpenAIToAnthropicStream& S, std::string& out) {  // emitClosingEvents :717-757
  if (S.closing) return;
  S.closing = true;
  if (S.open) { an_sse(out, "content_block_stop", "{\"type\":\"content_block_stop\",\"index\":" + std::to_string(S.block_index) + "}"); S.open = false; }
  std::string sr = S.stop_reason.empty() ? "end_turn" : S.stop_reason;
  an_sse(out, "message_delta", "{\"type\":\"message_delta\",\"delta\":{\"stop_reason\":\"" + sr + "\",\"stop_sequence\":null},\"usage\":{\"output_tokens\":" + std::to_string(S.out_tok) + "}}");
  an_sse(out, "message_stop", "{\"type\":\"message_stop\"}");
}

inline void o2a_block(OpenAIToAnthropicStream& S, std::string_view block, std::string& out) {  // processEventBlock :492-520 + handleChunk :522-589
  std::string data;
  size_t pos = 0;
  for (;;) {
    size_t nl = block.find('\n', pos);
    std::string_view line = block.substr(pos, nl == std::string_view::npos ? std::string_view::npos : nl - pos);
    if (line.substr(0, 6) == "data: ") { std::string_view t = trim_space(line.substr(6)); if (!t.empty()) data = std::string(t); }
    if (nl == std::string_view::npos) break;
    pos = nl + 1;
  }
  if (data.empty() || data == "[DONE]") return;
  std::string m; bool hu; TokenUsage tu;
  if (!decode_chunk(data, m, hu, tu)) return;   // malformed chunks are skipped silently
  Value v; std::string err; oj::parse(data, v, err);
  auto str = [](const Value* x) { return x && x->is_str() ? x->s : std::string(); };
  if (!v.is_obj()) return;   // `null` decodes into the zero chunk: no id, no choices, no usage
  const std::string id = str(v.get("id")), model = str(v.get("model"));
  if (!id.empty() && S.id.empty()) S.id = id;
  if (!model.empty() && S.model.empty()) S.model = model;
  const Value* ch = v.get("choices"); const Value* us = v.get("usage");
  const bool no_choices = !ch || !ch->is_arr() || ch->arr.empty();
  if (no_choices && us && us->is_obj()) {
    int64_t p = 0, c = 0; int_field(us->get("prompt_tokens"), p); int_field(us->get("completion_tokens"), c);
    S.in_tok = p; S.out_tok = c;
    S.usage = explicit_caching(p, c, true);
    o2a_closing(S, out);
    return;
  }
  if (no_choices) return;
  const Value& c0 = ch->arr[0];
  const Value* delta = c0.is_obj() ? c0.get("delta") : nullptr;
  if (delta && !delta->is_obj()) delta = nullptr;
  if (!S.started && delta) {  // emitMessageStart :591-613
    S.started = true;
    std::string d = "{\"type\":\"message_start\",\"message\":{\"id\":"; oj::enc_str(d, S.id);
    d += ",\"type\":\"message\",\"role\":\"assistant\",\"content\":[],\"model\":"; oj::enc_str(d, S.model.empty() ? S.request_model : S.model);
    d += ",\"stop_reason\":null,\"stop_sequence\":null,\"usage\":{\"input_tokens\":0,\"output_tokens\":0}}}";
    an_sse(out, "message_start", d);
  }
  if (delta) {
    const Value* content = delta->get("content");
    if (content && content->is_str() && !content->s.empty()) {
      if (!S.open) {
        S.open = true;
        an_sse(out, "content_block_start", "{\"type\":\"content_block_start\",\"index\":" + std::to_string(S.block_index) + ",\"content_block\":{\"type\":\"text\",\"text\":\"\"}}");
      }
      std::string d = "{\"type\":\"content_block_delta\",\"index\":" + std::to_string(S.block_index) + ",\"delta\":{\"type\":\"text_delta\",\"text\":"; oj::enc_str(d, content->s); d += "}}";
      an_sse(out, "content_block_delta", d);
    }
    const Value* tcs = delta->get("tool_calls");
    if (tcs && tcs->is_arr()) for (const Value& tc : tcs->arr) {  // handleToolCallDelta :647-704; a null element is the zero tool call
      int64_t idx = 0; std::string tid, name, args;
      if (tc.is_obj()) {
        int_field(tc.get("index"), idx); tid = str(tc.get("id"));
        if (const Value* f = tc.get("function"); f && f->is_obj()) { name = str(f->get("name")); args = str(f->get("arguments")); }
      }
      auto it = S.tools.find(idx);
      int64_t blk;
      if (it == S.tools.end()) {
        if (S.open) { an_sse(out, "content_block_stop", "{\"type\":\"content_block_stop\",\"index\":" + std::to_string(S.block_index) + "}"); S.block_index++; }
        blk = S.block_index; S.tools[idx] = blk; S.open = true;
        std::string d = "{\"type\":\"content_block_start\",\"index\":" + std::to_string(blk) + ",\"content_block\":{\"type\":\"tool_use\",\"id\":"; oj::enc_str(d, tid);
        d += ",\"name\":"; oj::enc_str(d, name); d += ",\"input\":{}}}";
        an_sse(out, "content_block_start", d);
      } else blk = it->second;
      if (!args.empty()) {
        std::string d = "{\"type\":\"content_block_delta\",\"index\":" + std::to_string(blk) + ",\"delta\":{\"type\":\"input_json_delta\",\"partial_json\":"; oj::enc_str(d, args); d += "}}";
        an_sse(out, "content_block_delta", d);
      }
    }
  }
  const std::string fr = c0.is_obj() ? str(c0.get("finish_reason")) : std::string();
  if (!fr.empty()) S.stop_reason = openai_finish_to_anthropic(fr);
}

// One ResponseBody(stream) call (anthropic_openai.go:154-185, processBuffer openai_helper.go:460-490): `out` is the call's body (always
// non-nil, possibly empty), `usage` the state's TokenUsage, `response_model` = cmp.Or(state.model, requestModel).
inline Status messages_openai_stream_feed(OpenAIToAnthropicStream& S, std::string_view chunk, bool eos, std::string& out, TokenUsage& usage, std::string& response_model) {
  S.buffer.append(chunk);
  out.clear();
  for (;;) {
    size_t cut = S.buffer.find("\n\n");
    if (cut == std::string::npos) break;
    std::string block = S.buffer.substr(0, cut);
    S.buffer.erase(0, cut + 2);
    o2a_block(S, block, out);
  }
  if (eos) {
    if (!S.buffer.empty()) { std::string rem; rem.swap(S.buffer); o2a_block(S, rem, out); }
    if (!S.closing) o2a_closing(S, out);
  }
  usage = S.usage;
  response_model = S.model.empty() ? S.request_model : S.model;
  return OK;
}

// ---------------------------------------------------------------- buffered
// responseBodyNonStreaming (anthropic_openai.go:112-152): json.NewDecoder(body).Decode(&ChatCompletionResponse{}) — trailing bytes are
// not an error — then openAIResponseToAnthropic.  Returns false for "failed to unmarshal OpenAI response".
inline bool messages_openai_response(std::string_view body, const std::string& request_model, std::string& out, TokenUsage& usage, std::string& response_model) {
  TokenUsage tu0; std::string rm0;
  response_model = request_model; usage = TokenUsage{}; out.clear();
  if (!response_openai(body, request_model, tu0, rm0)) return false;
  Value v; oj::Parser ps(body.data(), body.size()); ps.value(v);
  auto str = [](const Value* x) { return x && x->is_str() ? x->s : std::string(); };
  std::string id, model; int64_t p = 0, c = 0; const Value* ch = nullptr;
  if (v.is_obj()) {
    id = str(v.get("id")); model = str(v.get("model")); ch = v.get("choices");
    if (const Value* us = v.get("usage"); us && us->is_obj()) { int_field(us->get("prompt_tokens"), p); int_field(us->get("completion_tokens"), c); }
  }
  response_model = model.empty() ? request_model : model;
  usage = explicit_caching(p, c, false);
  std::string blocks; bool any = false, have_choice = false; std::string stop;
  if (ch && ch->is_arr() && !ch->arr.empty()) {
    have_choice = true;
    const Value& c0 = ch->arr[0];
    const Value* msg = c0.is_obj() ? c0.get("message") : nullptr;
    if (msg && msg->is_obj()) {
      const Value* content = msg->get("content");
      if (content && content->is_str() && !content->s.empty()) { blocks += "{\"type\":\"text\",\"text\":"; oj::enc_str(blocks, content->s); blocks += "}"; any = true; }
      const Value* tcs = msg->get("tool_calls");
      if (tcs && tcs->is_arr()) for (const Value& tc : tcs->arr) {
        std::string tid, name, args;
        if (tc.is_obj()) { tid = str(tc.get("id")); if (const Value* f = tc.get("function"); f && f->is_obj()) { name = str(f->get("name")); args = str(f->get("arguments")); } }
        std::string input = "{}";
        if (!args.empty()) {  // json.Unmarshal into map[string]any; anything else (error, null, non-object) leaves the empty map
          Value a; std::string e2;
          if (oj::parse(args, a, e2) && a.is_obj()) { input.clear(); oj::enc_any(input, a); }
        }
        if (any) blocks.push_back(',');
        any = true;
        blocks += "{\"type\":\"tool_use\",\"id\":"; oj::enc_str(blocks, tid); blocks += ",\"name\":"; oj::enc_str(blocks, name); blocks += ",\"input\":" + input + "}";
      }
    }
    stop = openai_finish_to_anthropic(c0.is_obj() ? str(c0.get("finish_reason")) : std::string());
  }
  out = "{\"id\":"; oj::enc_str(out, id); out += ",\"type\":\"message\",\"role\":\"assistant\",\"content\":";
  if (any) out += "[" + blocks + "]"; else out += "null";   // a nil []MessagesContentBlock
  out += ",\"model\":"; oj::enc_str(out, response_model);
  if (have_choice) out += ",\"stop_reason\":\"" + stop + "\"";
  out += ",\"usage\":{\"cache_creation_input_tokens\":0,\"cache_read_input_tokens\":0,\"input_tokens\":" + std::to_string(p) + ",\"output_tokens\":" + std::to_string(c) + "}}";
  return true;
}

}  // namespace oracle
