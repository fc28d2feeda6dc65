// TEST INFRASTRUCTURE (oracle): CPU restatement of the streamed response path of OpenAI chat completions to Anthropic on AWS Bedrock.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it; the product never does.
//
//   openAIToAWSAnthropicTranslatorV1ChatCompletion.ResponseBody (stream branch)   internal/translator/openai_awsanthropic.go:162-184
//   extractAnthropicSSEFromEventStream                                           internal/translator/openai_awsanthropic.go:216-261
//   eventstream.Decoder (aws-sdk-go-v2 eventstream v1.7.10): frame format as in bedrock_stream.hpp
//
// Every frame's payload is {"bytes":"<base64 of one Anthropic event JSON>"}.  The reference decodes it, takes gjson "type" of the
// decoded text and hands "event: T\ndata: J\n\n" to the Anthropic stream parser (anthropic_stream.hpp).  DECLINED (sticky) where the
// outcome depends on library behaviour that is not restated: escapes in the base64 string or in member names, case-folded field
// names, a decoded event that is not valid JSON (gjson on malformed text), a non-string "type", a line break inside the event.
// Pinned by the Anthropic parser's own goldens (the unwrapped text is fed to the same parser); parity unpinned for nothing new.
#pragma once
#include "anthropic_stream.hpp"
#include "bedrock_stream.hpp"

namespace oracle {

struct AwsAnthropicStreamState { std::string buffered; AnthropicStreamState an; bool dead = false; };

inline Status aws_anthropic_stream_feed(AwsAnthropicStreamState& st, const AnthropicStreamCfg& cfg, std::string_view chunk, bool eos, std::string& out, TokenUsage& usage) {
  usage = TokenUsage{};
  if (st.dead) return DECLINED;
  auto decline = [&]() { st.dead = true; return DECLINED; };
  st.buffered.append(chunk);
  std::string sse;
  const uint8_t* p = (const uint8_t*)st.buffered.data(); size_t n = st.buffered.size(), off = 0;
  for (;;) {
    if (n - off < 12) break;
    const uint32_t total = be32(p + off), hlen = be32(p + off + 4), pcrc = be32(p + off + 8);
    if (crc32_ieee(p + off, 8) != pcrc) break;
    if (hlen > 128u * 1024 || total < 16u || hlen > total - 16u || total - hlen - 16u > 16u * 1024 * 1024) break;
    if (n - off < total) break;
    if (crc32_ieee(p + off, total - 4) != be32(p + off + total - 4)) break;
    bool bad_headers = false;
    size_t h = off + 12, hend = h + hlen;
    while (h < hend) {
      const size_t nl = p[h]; h++;
      if (h + nl + 1 > hend) { bad_headers = true; break; }
      h += nl;
      const int type = p[h]; h++;
      int vs = header_value_size(type);
      if (vs == -2) { bad_headers = true; break; }
      if (vs == -1) { if (h + 2 > hend) { bad_headers = true; break; } vs = (p[h] << 8) | p[h + 1]; h += 2; }
      if (h + vs > hend) { bad_headers = true; break; }
      h += vs;
    }
    if (bad_headers) break;
    const std::string_view payload((const char*)p + off + 12 + hlen, total - hlen - 16);
    off += total;
    Value v; std::string err;
    if (!oj::parse(payload, v, err)) continue;               // json.Unmarshal error: the frame is skipped
    if (!v.is_obj()) continue;
    const Value* bytes = nullptr; int nb = 0; bool fold = false;
    for (auto& kv : v.obj) {
      if (kv.first == "bytes") { bytes = &kv.second; nb++; }
      else if (kv.first.size() == 5) { std::string l = kv.first; for (auto& c : l) c = (char)tolower((unsigned char)c); if (l == "bytes") fold = true; }
    }
    // member names spelled with escapes
    {
      bool in_str = false, esc_key = false; size_t start = 0; bool esc = false;
      for (size_t i = 0; i < payload.size(); i++) {
        const char c = payload[i];
        if (!in_str) { if (c == '"') { in_str = true; start = i; esc = false; } continue; }
        if (c == '\\') { esc = true; i++; continue; }
        if (c == '"') { in_str = false; size_t k = i + 1; while (k < payload.size() && (unsigned char)payload[k] <= ' ') k++; if (esc && k < payload.size() && payload[k] == ':') esc_key = true; }
      }
      (void)start;
      if (esc_key) return decline();
    }
    if (nb > 1 || fold) return decline();
    if (!bytes || !bytes->is_str()) continue;
    if (payload.substr(bytes->b, bytes->e - bytes->b).find('\\') != std::string_view::npos) return decline();
    if (bytes->s.empty()) continue;
    for (char c : bytes->s) if (c == '\r' || c == '\n') return decline();
    std::string J;
    if (!oj::b64dec(bytes->s, J)) continue;                   // base64 error: skipped
    if (J.find('\n') != std::string::npos) return decline();
    Value ev; std::string e2;
    if (!oj::parse(J, ev, e2)) return decline();
    std::string type;
    if (ev.is_obj()) {
      bool in_str = false, esc = false, esc_key = false;
      for (size_t i = 0; i < J.size(); i++) {
        const char c = J[i];
        if (!in_str) { if (c == '"') { in_str = true; esc = false; } continue; }
        if (c == '\\') { esc = true; i++; continue; }
        if (c == '"') { in_str = false; size_t k = i + 1; while (k < J.size() && (unsigned char)J[k] <= ' ') k++; if (esc && k < J.size() && J[k] == ':') esc_key = true; }
      }
      if (esc_key) return decline();
      if (const Value* t = ev.get_first("type")) {
        if (!t->is_str()) return decline();
        if (std::string_view(J).substr(t->b, t->e - t->b).find('\\') != std::string_view::npos) return decline();
        type = t->s;
      }
    }
    sse += "event: " + type + "\ndata: " + J + "\n\n";
  }
  st.buffered.erase(0, off);
  const Status s = anthropic_stream_feed(st.an, cfg, sse, eos, out, usage);
  if (s != OK) st.dead = true;
  return s;
}

}  // namespace oracle
