// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the route/backend body mutation:
//   BodyMutator.Mutate / isJSONValue     internal/bodymutator/body_mutator.go:33-119
//   applyBodyMutation                    internal/extproc/util.go:107-131 (applied to the translated body, else the original)
//   config type                          internal/filterapi/filterconfig.go:258-277 ("only top-level fields are supported")
// The edits themselves are github.com/tidwall/sjson v1.2.6-0.20251103175603 (DeleteBytes, SetRawBytesOptions, SetBytesOptions
// with ReplaceInPlace and Optimistic=false) on github.com/tidwall/gjson v1.18.0 lookups; neither is in the tree.  Restated
// from their published algorithm: gjson.Get returns the FIRST member with that key; set on an existing key splices the value
// bytes; set on a missing key rebuilds the root as raw[first non-space .. last '}') + [","] + "key":value + "}" (bytes outside
// the braces are dropped); delete removes the member and ONE adjacent comma — the one in front of it, or, for the first
// member, the one behind it (deleteTailItem) — and a missing key is a no-op.  Pinned by the byte-exact goldens at
// tests/data-plane/testupstream_test.go:1160-1162 (delete last member, in-place set, two appends) and :1172-1174.
// Only plain top-level key paths are restated (no '.', '*', '?', '#', '|', ':', '\\', no all-digit keys): anything else is
// "parity unpinned" and reported as unsupported.
#pragma once
#include "translate.hpp"

namespace oracle {

struct BodySet { std::string path, value; };

inline bool mutate_is_json_value(std::string_view v) {  // body_mutator.go:33-75
  size_t b = 0, e = v.size();
  auto sp = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r' || c == 0x85 || c == 0xA0; };
  while (b < e && sp((unsigned char)v[b])) b++;
  while (e > b && sp((unsigned char)v[e - 1])) e--;
  v = v.substr(b, e - b);
  if (v.size() >= 1 && v.front() == '"' && v.back() == '"') return true;
  if (v == "0" || v == "true" || v == "false" || v == "null") return true;
  if (!v.empty()) {
    const char f = v[0];
    if ((f >= '0' && f <= '9') || f == '-' || f == '+') {
      bool num = true;
      for (char c : v) if ((c < '0' || c > '9') && c != '.' && c != '-' && c != '+' && c != 'e' && c != 'E') { num = false; break; }
      if (num) return true;
    }
  }
  if (v.size() >= 1 && v.front() == '{' && v.back() == '}') return true;
  if (v.size() >= 1 && v.front() == '[' && v.back() == ']') return true;
  return false;
}

inline bool mutate_path_supported(std::string_view p) {
  if (p.empty()) return true;  // skipped by the reference
  bool digits = true;
  for (unsigned char c : p) {
    if (c == '.' || c == '*' || c == '?' || c == '#' || c == '|' || c == ':' || c == '\\' || c == '@' || c < 0x21 || c > 0x7e || c == '"') return false;
    if (c < '0' || c > '9') digits = false;
  }
  return !digits && p != "-1";
}

// sjson.DeleteBytes(body, key) for a top-level key
inline std::string sjson_delete_top(std::string_view body, const Value& root, std::string_view key) {
  if (!root.is_obj()) return std::string(body);
  // first occurrence: need the member's position in source order
  size_t idx = 0; const Value* val = nullptr;
  for (; idx < root.obj.size(); idx++) if (root.obj[idx].first == key) { val = &root.obj[idx].second; break; }
  if (!val) return std::string(body);
  // deleteTailItem on body[:val.b]: back over ':' and the key string to the ',' or '{' in front
  size_t i = val->b;
  while (i > 0 && body[i - 1] != ':') i--;          // now body[i-1] == ':'
  i--;                                              // index of ':'
  size_t q = i; while (q > 0 && body[q - 1] != '"') q--;  // closing quote at q-1
  size_t k = q - 1;                                 // closing quote
  // opening quote: scan back for an unescaped '"'
  size_t o = k;
  for (;;) { o--; if (body[o] == '"' && !(o > 0 && body[o - 1] == '\\')) break; }
  size_t c = o;                                     // scan back to ',' or '{'
  bool del_next_comma = false; size_t cut;
  for (;;) { c--; if (body[c] == '{') { cut = c + 1; del_next_comma = true; break; } if (body[c] == ',') { cut = c; break; } }
  std::string out(body.substr(0, cut));
  size_t r = val->e;
  if (del_next_comma) { while (r < body.size()) { if ((unsigned char)body[r] <= ' ') { r++; continue; } if (body[r] == ',') r++; break; } }
  out += body.substr(r);
  return out;
}

// Mutate(): 0 ok, 1 unsupported path/value (not restated)
inline int body_mutate(std::string_view body, const std::vector<std::string>& removes, const std::vector<BodySet>& sets, std::string& out) {
  out.assign(body);
  for (auto& r : removes) if (!mutate_path_supported(r)) return 1;
  for (auto& s : sets) if (!mutate_path_supported(s.path)) return 1;
  for (auto& r : removes) {
    if (r.empty()) continue;
    Value root; std::string err;
    if (!oj::parse(out, root, err)) return 1;
    out = sjson_delete_top(out, root, r);
  }
  for (auto& s : sets) {
    if (s.path.empty()) continue;
    Value root; std::string err;
    if (!oj::parse(out, root, err)) return 1;
    std::string raw;
    if (mutate_is_json_value(s.value)) raw = s.value; else sjson_stringify(raw, s.value);
    out = sjson_set_raw(out, root, s.path, "", raw);
  }
  return 0;
}

}  // namespace oracle
