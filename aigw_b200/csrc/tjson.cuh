// Per-thread, table-driven typed JSON walk (device).  One thread validates one JSON text against a
// schema table the way Go's json.Unmarshal validates it against a struct type (unknown keys
// skipped after a syntax check, `null` accepted for every field, duplicate keys decode again) and
// captures a few scalar fields.  Schemas are data: a new response type is a new table.
//
// Decoder behaviour restated (github.com/bytedance/sonic v1.15.1 defaults, listed in DESIGN.md):
// control characters inside strings are accepted, UTF-8 is not validated, escapes must be
// well-formed, numbers follow RFC 8259, integer fields reject fraction/exponent/overflow.
#pragma once
#include <stdint.h>

namespace aigw { namespace tj {

enum Kind : uint8_t { K_ANY = 0, K_STR, K_INT, K_FLOAT, K_BOOL, K_OBJ, K_ARR, K_CREATED, K_B64, K_STROBJ /* string, or the object node in `elem` */,
                      K_MAP /* map[string]T: an object with arbitrary keys whose values are all of node `elem` */ };

struct Field { uint16_t koff; uint8_t klen; uint8_t node; };  // key bytes at keys[koff..koff+klen)
struct Node { uint8_t kind; uint8_t cap; uint8_t f0; uint8_t nf; uint8_t elem; };

// Capture sink: up to 8 integer captures (uint32 wrap), 2 span captures, "object present" bits.
template <int NSPANS>
struct CaptureT {
  uint32_t ints[8];
  uint32_t int_set;        // bit i: ints[i] was written
  uint32_t obj_seen;       // bit cap: a non-null object with that capture id was decoded
  uint32_t span_off[NSPANS], span_len[NSPANS];
  uint32_t span_set;       // bit i
  uint32_t span_esc;       // bit i: the captured string contains a backslash
  uint32_t weird;          // escape in a key, duplicate known key, nesting too deep: caller must not trust the result
  uint32_t big;            // a captured integer does not fit in 31 bits
};
using Capture = CaptureT<2>;

__device__ __forceinline__ bool ws(uint32_t c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }
__device__ __forceinline__ bool dig(uint32_t c) { return c - '0' < 10u; }
__device__ __forceinline__ bool delim(uint32_t c) { return ws(c) || c == ',' || c == '}' || c == ']'; }
__device__ __forceinline__ int hexv(uint32_t c) { if (c - '0' < 10u) return c - '0'; c |= 0x20; if (c - 'a' < 6u) return c - 'a' + 10; return -1; }

// string at p[i] == '"'.  Returns index after the closing quote, or -1.  esc: saw a backslash.
__device__ inline int scan_string(const uint8_t* p, int i, int n, bool& esc) {
  i++;
  while (i < n) {
    uint32_t c = p[i];
    if (c == '"') return i + 1;
    if (c == '\\') {
      esc = true;
      if (i + 1 >= n) return -1;
      uint32_t e = p[i + 1];
      if (e == 'u') { if (i + 6 > n) return -1; for (int k = 2; k < 6; k++) if (hexv(p[i + k]) < 0) return -1; i += 6; continue; }
      if (!(e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't')) return -1;
      i += 2; continue;
    }
    i++;
  }
  return -1;
}
// RFC 8259 number at p[i]; returns end index or -1; is_int: no fraction/exponent
__device__ inline int scan_number(const uint8_t* p, int i, int n, bool& is_int, int& int_end) {
  if (i < n && p[i] == '-') i++;
  if (i >= n) return -1;
  if (p[i] == '0') i++;
  else if (p[i] >= '1' && p[i] <= '9') { while (i < n && dig(p[i])) i++; }
  else return -1;
  int_end = i; is_int = true;
  if (i < n && p[i] == '.') { is_int = false; i++; if (i >= n || !dig(p[i])) return -1; while (i < n && dig(p[i])) i++; }
  if (i < n && (p[i] == 'e' || p[i] == 'E')) { is_int = false; i++; if (i < n && (p[i] == '+' || p[i] == '-')) i++; if (i >= n || !dig(p[i])) return -1; while (i < n && dig(p[i])) i++; }
  return i;
}
__device__ inline bool lit(const uint8_t* p, int i, int n, const char* w, int l) { if (i + l > n) return false; for (int k = 0; k < l; k++) if (p[i + k] != (uint8_t)w[k]) return false; return true; }
// strconv.ParseInt(s, 10, 64) on -?digits: false on overflow.  out = low 32 bits (Go uint32(int) conversion).
__device__ inline bool parse_i64(const uint8_t* p, int b, int e, uint32_t& low32, bool* big = nullptr) {
  bool neg = p[b] == '-'; if (neg) b++;
  unsigned long long v = 0;
  for (int i = b; i < e; i++) {
    unsigned d = p[i] - '0';
    if (v > (0xFFFFFFFFFFFFFFFFull - d) / 10ull) return false;
    v = v * 10ull + d;
  }
  if (neg) { if (v > 0x8000000000000000ull) return false; low32 = (uint32_t)(0ull - v); }
  else { if (v > 0x7FFFFFFFFFFFFFFFull) return false; low32 = (uint32_t)v; }
  if (big) *big = neg ? v != 0 : (v >> 31) != 0;
  return true;
}

// Untyped value: syntax only.  Returns end index or -1.
__device__ inline int skip_any(const uint8_t* p, int i, int n) {
  uint64_t isobj = 0; int depth = 0;
  // state machine identical in spirit to validate_tokens, but byte driven
  for (;;) {
    while (i < n && ws(p[i])) i++;
    if (i >= n) return -1;
    uint32_t c = p[i];
    bool value_done = false;
    if (c == '"') { bool e = false; i = scan_string(p, i, n, e); if (i < 0) return -1; value_done = true; }
    else if (c == '{') {
      if (depth >= 64) return -1;
      i++; while (i < n && ws(p[i])) i++;
      if (i < n && p[i] == '}') { i++; value_done = true; }
      else {
        isobj |= 1ull << depth; depth++;
        // key
        if (i >= n || p[i] != '"') return -1;
        bool e = false; i = scan_string(p, i, n, e); if (i < 0) return -1;
        while (i < n && ws(p[i])) i++;
        if (i >= n || p[i] != ':') return -1;
        i++; continue;
      }
    } else if (c == '[') {
      if (depth >= 64) return -1;
      i++; while (i < n && ws(p[i])) i++;
      if (i < n && p[i] == ']') { i++; value_done = true; }
      else { isobj &= ~(1ull << depth); depth++; continue; }
    } else if (c == 't') { if (!lit(p, i, n, "true", 4)) return -1; i += 4; value_done = true; }
    else if (c == 'f') { if (!lit(p, i, n, "false", 5)) return -1; i += 5; value_done = true; }
    else if (c == 'n') { if (!lit(p, i, n, "null", 4)) return -1; i += 4; value_done = true; }
    else { bool ii; int ie; i = scan_number(p, i, n, ii, ie); if (i < 0) return -1; value_done = true; }
    // after a value: close containers / next member
    while (value_done) {
      if (depth == 0) return i;
      while (i < n && ws(p[i])) i++;
      if (i >= n) return -1;
      uint32_t d = p[i];
      bool obj = (isobj >> (depth - 1)) & 1;
      if (d == ',') {
        i++;
        if (obj) {
          while (i < n && ws(p[i])) i++;
          if (i >= n || p[i] != '"') return -1;
          bool e = false; i = scan_string(p, i, n, e); if (i < 0) return -1;
          while (i < n && ws(p[i])) i++;
          if (i >= n || p[i] != ':') return -1;
          i++;
        }
        value_done = false;
      } else if ((d == '}' && obj) || (d == ']' && !obj)) { i++; depth--; }
      else return -1;
    }
  }
}

// Typed walk of p[0..n).  Returns true when json.Unmarshal into the schema's root type would succeed.
template <class CAP>
__device__ inline bool walk(const uint8_t* p, int n, const Node* nodes, const Field* fields, const char* keys, int root, CAP& cap, bool allow_trailing = false) {
  const int MAXD = 12;
  uint8_t st_node[MAXD]; uint32_t st_seen[MAXD];
  int sp = 0;
  int node = root;
  int i = 0;
  for (;;) {
    // ---- a value of type `node` starts here
    while (i < n && ws(p[i])) i++;
    if (i >= n) return false;
    Node nd = nodes[node];
    uint32_t c = p[i];
    if (nd.kind == K_STROBJ) { if (c == '"' || c == 'n') nd.kind = K_STR; else { node = nd.elem; nd = nodes[node]; } }
    bool descend = false;
    if (nd.kind == K_ANY) {
      const int j = skip_any(p, i, n); if (j < 0) return false;
      if (nd.cap != 0xff) { const uint32_t s = nd.cap; cap.span_off[s] = i; cap.span_len[s] = j - i; cap.span_set |= 1u << s; }   // raw value (quotes included for strings)
      i = j;
    }
    else if (c == 'n') {
      if (!lit(p, i, n, "null", 4)) return false;
      i += 4;
      if (nd.kind == K_CREATED) return false;  // JSONUNIXTime.UnmarshalJSON is called with "null" and fails
    } else switch (nd.kind) {
      case K_STR: case K_B64: {
        if (c != '"') return false;
        bool e = false; int j = scan_string(p, i, n, e); if (j < 0) return false;
        if (nd.kind == K_B64) {  // []byte: base64.StdEncoding of the decoded string
          if (e) cap.weird = 1;
          int L = j - i - 2; const uint8_t* q = p + i + 1;
          if (L % 4) return false;
          for (int k = 0; k < L; k++) { uint32_t ch = q[k]; if (ch < 0x20) cap.weird = 1; bool ok = (ch - 'A' < 26u) || (ch - 'a' < 26u) || (ch - '0' < 10u) || ch == '+' || ch == '/' || (ch == '=' && k >= L - 2); if (!ok) return false; }
          if (L >= 2 && q[L - 2] == '=' && q[L - 1] != '=') return false;
        }
        if (nd.cap != 0xff) { const uint32_t s = nd.cap; cap.span_off[s] = i + 1; cap.span_len[s] = j - i - 2; cap.span_set |= 1u << s; if (e) cap.span_esc |= 1u << s; else cap.span_esc &= ~(1u << s); }
        i = j; break;
      }
      case K_INT: case K_FLOAT: case K_CREATED: {
        if (!(c == '-' || dig(c))) return false;
        bool ii; int ie; int j = scan_number(p, i, n, ii, ie); if (j < 0) return false;
        if (j < n && !delim(p[j])) return false;
        if (nd.kind == K_INT) { if (!ii) return false; uint32_t v; bool bg = false; if (!parse_i64(p, i, j, v, &bg)) return false; if (nd.cap != 0xff) { cap.ints[nd.cap] = v; cap.int_set |= 1u << nd.cap; if (bg) cap.big = 1; } }
        else if (nd.kind == K_CREATED) {  // integer part up to '.', exponent without '.' fails ParseInt
          if (!ii && (ie >= j || p[ie] != '.')) return false;
          uint32_t v; if (!parse_i64(p, i, ie, v)) return false;
        }
        i = j; break;
      }
      case K_BOOL: {
        if (c == 't') { if (!lit(p, i, n, "true", 4)) return false; i += 4; }
        else if (c == 'f') { if (!lit(p, i, n, "false", 5)) return false; i += 5; }
        else return false;
        break;
      }
      case K_OBJ: case K_MAP: {
        if (c != '{') return false;
        if (nd.cap != 0xff) cap.obj_seen |= 1u << nd.cap;
        i++; while (i < n && ws(p[i])) i++;
        if (i < n && p[i] == '}') { i++; break; }
        if (sp >= MAXD) { cap.weird = 1; return false; }
        st_node[sp] = (uint8_t)node; st_seen[sp] = 0; sp++;
        descend = true; break;
      }
      case K_ARR: {
        if (c != '[') return false;
        i++; while (i < n && ws(p[i])) i++;
        if (i < n && p[i] == ']') { i++; break; }
        if (sp >= MAXD) { cap.weird = 1; return false; }
        st_node[sp] = (uint8_t)node; st_seen[sp] = 0; sp++;
        node = nd.elem; continue;
      }
      default: return false;
    }
    // ---- after a value (or at the first member of an object)
    for (;;) {
      if (!descend) {
        if (sp == 0) { if (allow_trailing) return true; while (i < n && ws(p[i])) i++; return i == n; }
        while (i < n && ws(p[i])) i++;
        if (i >= n) return false;
        const Node par = nodes[st_node[sp - 1]];
        uint32_t d = p[i];
        if (par.kind == K_ARR) {
          if (d == ',') { i++; node = par.elem; break; }
          if (d == ']') { i++; sp--; continue; }
          return false;
        }
        if (d == '}') { i++; sp--; continue; }
        if (d != ',') return false;
        i++; while (i < n && ws(p[i])) i++;
      }
      descend = false;
      // object member: key ':'
      if (i >= n || p[i] != '"') return false;
      bool e = false; int j = scan_string(p, i, n, e); if (j < 0) return false;
      const uint8_t* k = p + i + 1; int kl = j - i - 2;
      if (e) cap.weird = 1;
      i = j; while (i < n && ws(p[i])) i++;
      if (i >= n || p[i] != ':') return false;
      i++;
      const Node par = nodes[st_node[sp - 1]];
      int child = -1;
      for (int f = 0; f < par.nf; f++) {
        const Field fd = fields[par.f0 + f];
        if (fd.klen == kl) { const char* fk = keys + fd.koff; int t = 0; while (t < kl && k[t] == (uint8_t)fk[t]) t++; if (t == kl) { child = fd.node; if (st_seen[sp - 1] & (1u << f)) cap.weird = 1; st_seen[sp - 1] |= 1u << f; break; } }
      }
      if (par.kind == K_MAP) child = par.elem;
      node = child < 0 ? 0 /* node 0 is K_ANY by convention */ : child;
      break;
    }
  }
}

}}  // namespace aigw::tj
