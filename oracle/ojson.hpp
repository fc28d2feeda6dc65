// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under aigw_b200/ may include, link or
// execute this file.  It is the CPU restatement of the reference's JSON arithmetic
// and exists to check the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs).
//
// What it restates
//   The reference funnels all JSON through github.com/bytedance/sonic v1.15.1
//   (wrapper: internal/json/json.go:16-35, `CaseSensitive: true`, everything else
//   default) and github.com/tidwall/{gjson v1.18.0, sjson v1.2.6}.  Neither module
//   is vendored under /root/reference, so this file restates their *published*
//   behaviour; every byte-level choice that no in-tree golden pins is listed in
//   DESIGN.md ("parity unpinned") and summarised here:
//     * decode: RFC 8259 grammar; raw control characters inside strings are accepted
//       (sonic default ValidateString=false); UTF-8 is not validated; \uXXXX escapes
//       are decoded, lone surrogates become U+FFFD; duplicate keys: last one wins;
//       unknown struct fields ignored; field names matched case-sensitively.
//     * encode: compact; '"' '\\' '\n' '\r' '\t' use short escapes, other bytes < 0x20
//       use \u00XX (lower-case hex); no HTML escaping (<,>,& and U+2028/9 pass
//       through); non-ASCII bytes pass through untouched; map keys are emitted in
//       sorted byte order (internal/json/json.go:26-28 says order is unspecified;
//       sorted satisfies every golden).
//     * float64 formatting follows Go strconv ('f' unless exp < -6 || exp >= 21,
//       shortest round-trip digits), as encoding/json and sonic both do.
#pragma once
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace oj {

enum class T : uint8_t { Null, False, True, Number, String, Array, Object };

struct Value {
  T t = T::Null;
  uint32_t b = 0, e = 0;  // raw span [b,e) in the source text
  std::string s;          // String: decoded bytes.  Number: raw literal.
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;  // decoded keys, source order

  bool is_null() const { return t == T::Null; }
  bool is_bool() const { return t == T::True || t == T::False; }
  bool is_str() const { return t == T::String; }
  bool is_num() const { return t == T::Number; }
  bool is_arr() const { return t == T::Array; }
  bool is_obj() const { return t == T::Object; }
  // struct-field lookup as encoding/json/sonic do it: every occurrence is decoded
  // into the same field, so the LAST occurrence decides.
  const Value* get(std::string_view k) const {
    const Value* r = nullptr;
    for (auto& kv : obj)
      if (kv.first == k) r = &kv.second;
    return r;
  }
  // gjson.Get semantics: FIRST occurrence.
  const Value* get_first(std::string_view k) const {
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

inline void append_utf8(std::string& o, uint32_t cp) {
  if (cp < 0x80) o.push_back((char)cp);
  else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) {
    o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F)));
  }
}

struct Parser {
  const char* p; size_t n; size_t i = 0; std::string err; int depth = 0;
  Parser(const char* p_, size_t n_) : p(p_), n(n_) {}
  void ws() { while (i < n && (p[i] == ' ' || p[i] == '\t' || p[i] == '\n' || p[i] == '\r')) i++; }
  bool fail(const char* m) { if (err.empty()) { err = m; err += " at offset " + std::to_string(i); } return false; }
  static int hex(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
  }
  bool hex4(uint32_t& out) {
    if (i + 4 > n) return fail("truncated \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) { int h = hex(p[i + k]); if (h < 0) return fail("invalid \\u escape"); v = v * 16 + h; }
    i += 4; out = v; return true;
  }
  bool str(std::string& out) {  // at opening quote
    i++;
    for (;;) {
      if (i >= n) return fail("unterminated string");
      unsigned char c = (unsigned char)p[i];
      if (c == '"') { i++; return true; }
      if (c == '\\') {
        i++;
        if (i >= n) return fail("unterminated escape");
        char ecc = p[i++];
        switch (ecc) {
          case '"': out.push_back('"'); break;
          case '\\': out.push_back('\\'); break;
          case '/': out.push_back('/'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'n': out.push_back('\n'); break;
          case 'r': out.push_back('\r'); break;
          case 't': out.push_back('\t'); break;
          case 'u': {
            uint32_t cp; if (!hex4(cp)) return false;
            if (cp >= 0xD800 && cp < 0xDC00) {
              if (i + 6 <= n && p[i] == '\\' && p[i + 1] == 'u') {
                size_t save = i; i += 2; uint32_t lo; if (!hex4(lo)) return false;
                if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                else { i = save; cp = 0xFFFD; }
              } else cp = 0xFFFD;
            } else if (cp >= 0xDC00 && cp < 0xE000) cp = 0xFFFD;
            append_utf8(out, cp); break;
          }
          default: i--; return fail("invalid escape character");
        }
        continue;
      }
      out.push_back((char)c); i++;  // raw byte, control characters included (sonic default)
    }
  }
  bool number(Value& v) {
    size_t s = i;
    if (i < n && p[i] == '-') i++;
    if (i >= n) return fail("invalid number");
    if (p[i] == '0') i++;
    else if (p[i] >= '1' && p[i] <= '9') { while (i < n && p[i] >= '0' && p[i] <= '9') i++; }
    else return fail("invalid number");
    if (i < n && p[i] == '.') { i++; if (i >= n || p[i] < '0' || p[i] > '9') return fail("invalid number"); while (i < n && p[i] >= '0' && p[i] <= '9') i++; }
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
      i++; if (i < n && (p[i] == '+' || p[i] == '-')) i++;
      if (i >= n || p[i] < '0' || p[i] > '9') return fail("invalid number"); while (i < n && p[i] >= '0' && p[i] <= '9') i++;
    }
    v.t = T::Number; v.s.assign(p + s, i - s); return true;
  }
  bool lit(const char* w, size_t l) { if (i + l <= n && memcmp(p + i, w, l) == 0) { i += l; return true; } return fail("invalid literal"); }
  bool value(Value& v) {
    ws();
    if (i >= n) return fail("unexpected end of input");
    if (++depth > 512) return fail("nesting too deep");
    v.b = (uint32_t)i;
    bool ok = true;
    char c = p[i];
    if (c == '{') {
      v.t = T::Object; i++; ws();
      if (i < n && p[i] == '}') i++;
      else for (;;) {
        ws();
        if (i >= n || p[i] != '"') { ok = fail("expected object key"); break; }
        std::string k; if (!str(k)) { ok = false; break; }
        ws(); if (i >= n || p[i] != ':') { ok = fail("expected ':'"); break; }
        i++;
        v.obj.emplace_back(std::move(k), Value{});
        if (!value(v.obj.back().second)) { ok = false; break; }
        ws();
        if (i < n && p[i] == ',') { i++; continue; }
        if (i < n && p[i] == '}') { i++; break; }
        ok = fail("expected ',' or '}'"); break;
      }
    } else if (c == '[') {
      v.t = T::Array; i++; ws();
      if (i < n && p[i] == ']') i++;
      else for (;;) {
        v.arr.emplace_back();
        if (!value(v.arr.back())) { ok = false; break; }
        ws();
        if (i < n && p[i] == ',') { i++; continue; }
        if (i < n && p[i] == ']') { i++; break; }
        ok = fail("expected ',' or ']'"); break;
      }
    } else if (c == '"') { v.t = T::String; ok = str(v.s); }
    else if (c == 't') { ok = lit("true", 4); v.t = T::True; }
    else if (c == 'f') { ok = lit("false", 5); v.t = T::False; }
    else if (c == 'n') { ok = lit("null", 4); v.t = T::Null; }
    else if (c == '-' || (c >= '0' && c <= '9')) ok = number(v);
    else ok = fail("invalid character");
    depth--;
    v.e = (uint32_t)i;
    return ok;
  }
};

// Whole-document parse (json.Unmarshal): one value, only whitespace after it.
inline bool parse(const char* p, size_t n, Value& out, std::string& err) {
  Parser ps(p, n);
  if (!ps.value(out)) { err = ps.err; return false; }
  ps.ws();
  if (ps.i != n) { ps.fail("invalid character after top-level value"); err = ps.err; return false; }
  return true;
}
inline bool parse(std::string_view s, Value& out, std::string& err) { return parse(s.data(), s.size(), out, err); }

// ---------------------------------------------------------------- encoding
inline void enc_str(std::string& o, std::string_view s) {
  static const char* hexd = "0123456789abcdef";
  o.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { o += "\\u00"; o.push_back(hexd[c >> 4]); o.push_back(hexd[c & 15]); }
        else o.push_back((char)c);
    }
  }
  o.push_back('"');
}

// Go strconv.AppendFloat(b, f, fmt, -1, bits) as encoding/json uses it
// (encoding/json/encode.go floatEncoder): fmt 'f' unless abs<1e-6 || abs>=1e21 → 'e',
// with the "e-09" → "e-9" clean-up.
inline void fmt_digits(std::string& o, bool neg, const char* digs, int nd, int dexp /*value = 0.d1d2.. * 10^dexp*/, double absv, bool f32) {
  bool use_e = absv != 0 && (f32 ? ((float)absv < 1e-6f || (float)absv >= 1e21f) : (absv < 1e-6 || absv >= 1e21));
  if (neg) o.push_back('-');
  if (!use_e) {
    if (dexp <= 0) { o += "0"; if (nd > 0 && !(nd == 1 && digs[0] == '0')) { o.push_back('.'); o.append((size_t)(-dexp), '0'); o.append(digs, nd); } }
    else if (dexp >= nd) { o.append(digs, nd); o.append((size_t)(dexp - nd), '0'); }
    else { o.append(digs, dexp); o.push_back('.'); o.append(digs + dexp, nd - dexp); }
  } else {
    o.push_back(digs[0]);
    if (nd > 1) { o.push_back('.'); o.append(digs + 1, nd - 1); }
    int ex = dexp - 1;
    o.push_back('e'); o.push_back(ex < 0 ? '-' : '+');
    int ax = ex < 0 ? -ex : ex;
    // strconv always writes at least two exponent digits; encoding/json strips the
    // leading zero of a two-digit negative exponent only ("e-09"→"e-9").
    std::string es = std::to_string(ax);
    if (es.size() < 2) es = "0" + es;
    if (ex < 0 && es.size() == 2 && es[0] == '0') es = es.substr(1);
    o += es;
  }
}
template <class F> inline void enc_float_t(std::string& o, F f, bool f32) {
  if (f == 0) { o += (std::signbit(f) ? "-0" : "0"); return; }
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf - 1, f, std::chars_format::scientific);  // shortest round-trip
  *r.ptr = 0;
  // d.ddddde[+-]XX
  bool neg = buf[0] == '-'; char* q = buf + (neg ? 1 : 0);
  char digs[32]; int nd = 0; char* e = q;
  while (e < r.ptr && *e != 'e') { if (*e != '.') digs[nd++] = *e; e++; }
  int ex = atoi(e + 1);
  fmt_digits(o, neg, digs, nd, ex + 1, std::fabs((double)f), f32);
}
inline void enc_f64(std::string& o, double f) { enc_float_t<double>(o, f, false); }
inline void enc_f32(std::string& o, float f) { enc_float_t<float>(o, f, true); }

// Re-encode a decoded `any` (Go interface{}): numbers go through float64, map keys sorted.
inline void enc_any(std::string& o, const Value& v) {
  switch (v.t) {
    case T::Null: o += "null"; break;
    case T::True: o += "true"; break;
    case T::False: o += "false"; break;
    case T::Number: enc_f64(o, strtod(v.s.c_str(), nullptr)); break;
    case T::String: enc_str(o, v.s); break;
    case T::Array:
      o.push_back('[');
      for (size_t k = 0; k < v.arr.size(); k++) { if (k) o.push_back(','); enc_any(o, v.arr[k]); }
      o.push_back(']'); break;
    case T::Object: {
      // decode into map[string]any: last duplicate wins; marshal: keys sorted
      std::vector<std::pair<std::string_view, const Value*>> m;
      for (auto& kv : v.obj) {
        bool dup = false;
        for (auto& x : m) if (x.first == kv.first) { x.second = &kv.second; dup = true; break; }
        if (!dup) m.emplace_back(kv.first, &kv.second);
      }
      std::sort(m.begin(), m.end(), [](auto& a, auto& b) { return a.first < b.first; });
      o.push_back('{');
      for (size_t k = 0; k < m.size(); k++) { if (k) o.push_back(','); enc_str(o, m[k].first); o.push_back(':'); enc_any(o, *m[k].second); }
      o.push_back('}'); break;
    }
  }
}

// ---------------------------------------------------------------- typed decode helpers
// Integer kinds: encoding/json / sonic reject fraction or exponent forms and overflow.
inline bool num_to_i64(const std::string& lit, int64_t& out) {
  if (lit.empty()) return false;
  for (size_t k = (lit[0] == '-' ? 1 : 0); k < lit.size(); k++) if (lit[k] < '0' || lit[k] > '9') return false;
  errno = 0; char* end = nullptr; long long v = strtoll(lit.c_str(), &end, 10);
  if (errno == ERANGE || *end) return false;
  out = v; return true;
}

// base64.StdEncoding
inline std::string b64enc(std::string_view s) {
  static const char* tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string o; size_t i = 0;
  for (; i + 3 <= s.size(); i += 3) {
    uint32_t v = ((unsigned char)s[i] << 16) | ((unsigned char)s[i + 1] << 8) | (unsigned char)s[i + 2];
    o.push_back(tb[v >> 18]); o.push_back(tb[(v >> 12) & 63]); o.push_back(tb[(v >> 6) & 63]); o.push_back(tb[v & 63]);
  }
  if (s.size() - i == 1) { uint32_t v = (unsigned char)s[i] << 16; o.push_back(tb[v >> 18]); o.push_back(tb[(v >> 12) & 63]); o += "=="; }
  else if (s.size() - i == 2) { uint32_t v = ((unsigned char)s[i] << 16) | ((unsigned char)s[i + 1] << 8); o.push_back(tb[v >> 18]); o.push_back(tb[(v >> 12) & 63]); o.push_back(tb[(v >> 6) & 63]); o.push_back('='); }
  return o;
}
// base64.StdEncoding.DecodeString: padding required, '\r' and '\n' ignored, trailing bits not checked (non-Strict).
inline bool b64dec(std::string_view s, std::string& out) {
  auto val = [](unsigned char c) -> int {
    if (c >= 'A' && c <= 'Z') return c - 'A'; if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52; if (c == '+') return 62; if (c == '/') return 63; return -1; };
  std::string t; t.reserve(s.size());
  for (char c : s) if (c != '\r' && c != '\n') t.push_back(c);
  if (t.size() % 4) return false;
  for (size_t i = 0; i < t.size(); i += 4) {
    int a = val(t[i]), b = val(t[i + 1]); if (a < 0 || b < 0) return false;
    bool last = i + 4 == t.size();
    if (t[i + 2] == '=') { if (!last || t[i + 3] != '=') return false; out.push_back((char)((a << 2) | (b >> 4))); continue; }
    int c = val(t[i + 2]); if (c < 0) return false;
    if (t[i + 3] == '=') { if (!last) return false; out.push_back((char)((a << 2) | (b >> 4))); out.push_back((char)((b << 4) | (c >> 2))); continue; }
    int d = val(t[i + 3]); if (d < 0) return false;
    out.push_back((char)((a << 2) | (b >> 4))); out.push_back((char)((b << 4) | (c >> 2))); out.push_back((char)((c << 6) | d));
  }
  return true;
}

}  // namespace oj
