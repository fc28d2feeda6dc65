// See chat_kernel.cuh for the design.  sm_100a only.
#include "chat_kernel.cuh"

namespace aigw {

__device__ __constant__ LitTable c_lits = make_lit_table();
static constexpr LitTable h_lits = make_lit_table();

#define FULL 0xffffffffu

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint32_t nib_from_ff(uint32_t m) {  // m: 0xFF per selected byte → 4-bit mask
  return ((m & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ bool is_ws(uint32_t c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }
__device__ __forceinline__ bool is_op(uint32_t c) { return c == '{' || c == '}' || c == '[' || c == ']' || c == ':' || c == ','; }
__device__ __forceinline__ bool is_digit(uint32_t c) { return c - '0' < 10u; }

// ------------------------------------------------------------------ token view of one JSON text
struct Doc {
  const uint8_t* s;   // bytes
  uint32_t len;       // valid bytes
  const uint16_t* tok;
  uint16_t* jmp;
  int nt;
  uint32_t kind;      // op source kind: 0 input, 2 scratch

  __device__ __forceinline__ uint32_t tb(int i) const { return s[tok[i]]; }
  __device__ __forceinline__ int next(int i) const {
    uint32_t b = tb(i);
    if (b == '{' || b == '[') return jmp[i] + 1;
    if (b == '"') return i + 2;
    return i + 1;
  }
  __device__ __forceinline__ uint32_t str_off(int i) const { return tok[i] + 1u; }
  __device__ __forceinline__ uint32_t str_len(int i) const { return (uint32_t)tok[i + 1] - tok[i] - 1u; }
  __device__ bool str_eq(int i, const char* lit, uint32_t n) const {
    if (str_len(i) != n) return false;
    const uint8_t* p = s + str_off(i);
    for (uint32_t k = 0; k < n; k++) if (p[k] != (uint8_t)lit[k]) return false;
    return true;
  }
  __device__ bool str_has_backslash(int i) const {
    const uint8_t* p = s + str_off(i); uint32_t n = str_len(i);
    for (uint32_t k = 0; k < n; k++) if (p[k] == '\\') return true;
    return false;
  }
  __device__ uint32_t scalar_end(int i) const {
    uint32_t p = tok[i];
    while (p < len) { uint32_t c = s[p]; if (is_ws(c) || is_op(c) || c == '"') break; p++; }
    return p;
  }
  __device__ __forceinline__ bool is_null(int i) const { return tb(i) == 'n'; }
};

// JSON scalar grammar: number | true | false | null
__device__ bool scalar_valid(const uint8_t* p, uint32_t n) {
  if (n == 0) return false;
  uint32_t c = p[0];
  if (c == 't') return n == 4 && p[1] == 'r' && p[2] == 'u' && p[3] == 'e';
  if (c == 'f') return n == 5 && p[1] == 'a' && p[2] == 'l' && p[3] == 's' && p[4] == 'e';
  if (c == 'n') return n == 4 && p[1] == 'u' && p[2] == 'l' && p[3] == 'l';
  uint32_t i = 0;
  if (p[i] == '-') { i++; if (i >= n) return false; }
  if (p[i] == '0') i++;
  else if (p[i] >= '1' && p[i] <= '9') { while (i < n && is_digit(p[i])) i++; }
  else return false;
  if (i < n && p[i] == '.') { i++; if (i >= n || !is_digit(p[i])) return false; while (i < n && is_digit(p[i])) i++; }
  if (i < n && (p[i] == 'e' || p[i] == 'E')) {
    i++; if (i < n && (p[i] == '+' || p[i] == '-')) i++;
    if (i >= n || !is_digit(p[i])) return false;
    while (i < n && is_digit(p[i])) i++;
  }
  return i == n;
}

// Stage 3: grammar check + bracket matching over the token list (single lane).
// Returns 0 ok, else an aigw_reason.
__device__ int validate_tokens(Doc& d) {
  const int MAXDEPTH = 48;
  uint16_t open_idx[MAXDEPTH];
  uint64_t isobj = 0;
  int depth = 0;
  // state: 0 value expected, 1 key or '}' , 2 key, 3 ':', 4 ',' or close, 5 value or ']', 6 done
  int st = 0;
  int i = 0;
  const int nt = d.nt;
  while (i < nt) {
    uint32_t b = d.tb(i);
    if (st == 6) return AIGW_R_SYNTAX;
    if (b == '"') {
      if (i + 1 >= nt || d.tb(i + 1) != '"') return AIGW_R_SYNTAX;
      if (st == 1 || st == 2) st = 3;
      else if (st == 0 || st == 5) st = depth ? 4 : 6;
      else return AIGW_R_SYNTAX;
      i += 2; continue;
    }
    if (b == '{' || b == '[') {
      if (!(st == 0 || st == 5)) return AIGW_R_SYNTAX;
      if (depth >= MAXDEPTH) return AIGW_R_DEPTH;
      open_idx[depth] = (uint16_t)i;
      if (b == '{') { isobj |= (1ull << depth); st = 1; } else { isobj &= ~(1ull << depth); st = 5; }
      depth++; i++; continue;
    }
    if (b == '}' || b == ']') {
      if (depth == 0) return AIGW_R_SYNTAX;
      bool obj = (isobj >> (depth - 1)) & 1;
      if (b == '}') { if (!obj || !(st == 1 || st == 4)) return AIGW_R_SYNTAX; }
      else { if (obj || !(st == 5 || st == 4)) return AIGW_R_SYNTAX; }
      depth--;
      d.jmp[open_idx[depth]] = (uint16_t)i;
      st = depth ? 4 : 6; i++; continue;
    }
    if (b == ':') { if (st != 3) return AIGW_R_SYNTAX; st = 0; i++; continue; }
    if (b == ',') {
      if (st != 4) return AIGW_R_SYNTAX;
      st = ((isobj >> (depth - 1)) & 1) ? 2 : 0; i++; continue;
    }
    // scalar
    if (!(st == 0 || st == 5)) return AIGW_R_SYNTAX;
    uint32_t e = d.scalar_end(i);
    if (!scalar_valid(d.s + d.tok[i], e - d.tok[i])) return AIGW_R_SYNTAX;
    st = depth ? 4 : 6; i++;
  }
  return st == 6 ? 0 : AIGW_R_SYNTAX;
}

// ------------------------------------------------------------------ copy-op plan
// op = kind(2) | len(14) | off(16); kind 0 input bytes, 1 literal table, 2 scratch.
// Main ops occupy [0, cap); "system" ops (emitted after the messages array) are parked in
// [cap, cap + kSysCap) and appended once the messages are done.
static constexpr int kSysCap = 96;
struct Plan {
  uint32_t* ops;
  int nops, nsys, cap;
  uint32_t olen;
  int err;               // aigw_reason, 0 = fine

  __device__ void push(uint32_t kind, uint32_t off, uint32_t len) {
    while (len) {
      uint32_t l = len < 16383u ? len : 16383u;
      if (nops > 0) {
        uint32_t p = ops[nops - 1];
        uint32_t pl = (p >> 16) & 0x3fffu, po = p & 0xffffu;
        if ((p >> 30) == kind && po + pl == off && pl + l <= 16383u) {
          ops[nops - 1] = (kind << 30) | ((pl + l) << 16) | po;
          olen += l; off += l; len -= l; continue;
        }
      }
      if (nops >= cap) { err = AIGW_R_OPS; return; }
      ops[nops++] = (kind << 30) | (l << 16) | off;
      olen += l; off += l; len -= l;
    }
  }
  __device__ void push_sys(uint32_t kind, uint32_t off, uint32_t len) {
    while (len) {
      uint32_t l = len < 16383u ? len : 16383u;
      if (nsys >= kSysCap) { err = AIGW_R_OPS; return; }
      ops[cap + nsys++] = (kind << 30) | (l << 16) | off;
      off += l; len -= l;
    }
  }
  __device__ __forceinline__ void lit(int id, bool sys = false) {
    uint32_t o = c_lits.off[id], l = c_lits.off[id + 1] - o;
    if (sys) push_sys(1, o, l); else push(1, o, l);
  }
  __device__ __forceinline__ void src(const Doc& d, uint32_t off, uint32_t len, bool sys = false) {
    if (sys) push_sys(d.kind, off, len); else push(d.kind, off, len);
  }
  __device__ void flush_sys() {
    for (int k = 0; k < nsys; k++) { uint32_t p = ops[cap + k]; push(p >> 30, p & 0xffffu, (p >> 16) & 0x3fffu); }
    nsys = 0;
  }
};

// Scratch bump allocator (per warp)
struct Scratch {
  uint8_t* p; uint32_t n, cap;
};

// ------------------------------------------------------------------ canonical scalars
// A JSON number literal that strconv would print back digit-for-digit once trailing fractional
// zeros are dropped: -?(0|[1-9]\d*)(\.\d+)? , ≤ 15 significant digits, |v| ≥ 1e-6 unless zero.
// Returns the emitted length (prefix of the literal), 0 when not canonical.
__device__ uint32_t canon_number(const uint8_t* p, uint32_t n, bool integer_only) {
  uint32_t i = 0;
  bool neg = false;
  if (p[0] == '-') { neg = true; i = 1; }
  uint32_t int_start = i;
  while (i < n && is_digit(p[i])) i++;
  uint32_t int_digits = i - int_start;
  if (int_digits == 0) return 0;
  bool int_zero = (int_digits == 1 && p[int_start] == '0');
  if (i == n) {
    if (integer_only) { if (int_digits > 18) return 0; if (neg && int_zero) return 0; return n; }
    if (int_digits > 15) return 0;
    return n;  // "-0" prints "-0" for a float64
  }
  if (integer_only || p[i] != '.') return 0;  // exponent form, or fraction where an int is required
  uint32_t dot = i; i++;
  uint32_t frac_start = i;
  while (i < n && is_digit(p[i])) i++;
  if (i != n) return 0;  // exponent
  uint32_t fe = n;
  while (fe > frac_start && p[fe - 1] == '0') fe--;
  uint32_t frac_digits = fe - frac_start;
  if (frac_digits == 0) return dot;  // "1.0" → "1", "-0.0" → "-0"
  uint32_t sig;
  if (int_zero) {
    uint32_t z = 0; while (frac_start + z < fe && p[frac_start + z] == '0') z++;
    if (z > 5) return 0;  // < 1e-6 switches strconv to exponent form
    sig = frac_digits - z;
  } else sig = int_digits + frac_digits;
  if (sig > 15) return 0;
  return fe;
}

// ------------------------------------------------------------------ generic canonical re-serialisation
// Go: decode into interface{} / map[string]any, then marshal ⇒ compact, keys sorted, numbers via float64.
__device__ int cmp_keys(const Doc& d, int a, int b) {  // key tokens (opening quotes)
  const uint8_t* pa = d.s + d.str_off(a); uint32_t la = d.str_len(a);
  const uint8_t* pb = d.s + d.str_off(b); uint32_t lb = d.str_len(b);
  uint32_t m = la < lb ? la : lb;
  for (uint32_t k = 0; k < m; k++) { if (pa[k] != pb[k]) return pa[k] < pb[k] ? -1 : 1; }
  return la == lb ? 0 : (la < lb ? -1 : 1);
}

__device__ void emit_scalar_any(const Doc& d, Plan& pl, int vi, bool sys) {
  uint32_t c = d.tb(vi);
  uint32_t e = d.scalar_end(vi), o = d.tok[vi];
  if (c == 't' || c == 'f' || c == 'n') { pl.src(d, o, e - o, sys); return; }
  uint32_t l = canon_number(d.s + o, e - o, false);
  if (!l) { pl.err = AIGW_R_NUMBER; return; }
  pl.src(d, o, l, sys);
}

// Emit value at token `root` canonically (compact, keys sorted, numbers as float64 prints them).
// Iterative with an explicit frame stack; object members are selected in key order.
__device__ void emit_any(const Doc& d, Plan& pl, int root, bool sys = false) {
  const int MAXF = 24;
  int f_open[MAXF]; int f_state[MAXF];  // array: current element token; object: last emitted key token (-1 none)
  int sp = 0;
  int vi = root;
  bool need_value = true;
  for (;;) {
    if (pl.err) return;
    if (need_value) {
      need_value = false;
      const uint32_t c = d.tb(vi);
      if (c == '"') pl.src(d, d.tok[vi], (uint32_t)d.tok[vi + 1] - d.tok[vi] + 1u, sys);
      else if (c == '[') {
        pl.src(d, d.tok[vi], 1, sys);
        if (d.tb(vi + 1) == ']') pl.src(d, d.tok[vi + 1], 1, sys);
        else {
          if (sp >= MAXF) { pl.err = AIGW_R_DEPTH; return; }
          f_open[sp] = vi; f_state[sp] = vi + 1; sp++;
          vi = vi + 1; need_value = true; continue;
        }
      } else if (c == '{') {
        pl.src(d, d.tok[vi], 1, sys);
        if (sp >= MAXF) { pl.err = AIGW_R_DEPTH; return; }
        f_open[sp] = vi; f_state[sp] = -1; sp++;
      } else emit_scalar_any(d, pl, vi, sys);
    }
    if (sp == 0) return;
    const int open = f_open[sp - 1];
    if (d.tb(open) == '[') {
      const int nx = d.next(f_state[sp - 1]);
      pl.src(d, d.tok[nx], 1, sys);  // ',' or ']'
      if (d.tb(nx) == ',') { f_state[sp - 1] = nx + 1; vi = nx + 1; need_value = true; }
      else sp--;
    } else {
      const int last = f_state[sp - 1];
      int best = -1, cnt = 0;
      for (int m = open + 1; d.tb(m) != '}';) {
        if (d.str_has_backslash(m)) { pl.err = AIGW_R_ESCAPE; return; }
        const int c1 = last < 0 ? 1 : cmp_keys(d, m, last);
        if (c1 == 0 && m != last) { pl.err = AIGW_R_DUP_KEY; return; }
        if (c1 > 0 && (best < 0 || cmp_keys(d, m, best) < 0)) best = m;
        int nx = d.next(m + 3);
        if (d.tb(nx) == ',') nx++;
        m = nx;
        if (++cnt > 64) { pl.err = AIGW_R_UNSUPPORTED_FIELD; return; }
      }
      if (best < 0) { pl.src(d, d.tok[d.jmp[open]], 1, sys); sp--; continue; }
      if (last >= 0) pl.lit(L_COMMA, sys);
      const uint32_t ko = d.tok[best], kc = d.tok[best + 1], colon = d.tok[best + 2];
      if (colon == kc + 1u) pl.src(d, ko, colon + 1u - ko, sys);
      else { pl.src(d, ko, kc + 1u - ko, sys); pl.lit(L_COLON, sys); }
      f_state[sp - 1] = best;
      vi = best + 3; need_value = true;
    }
  }
}

// Sequential tokenizer for a scratch-resident JSON text (tool-call arguments after unescaping).
// Returns number of tokens, or -reason.
__device__ int tokenize_seq(const uint8_t* s, uint32_t len, uint16_t* tok, int cap) {
  int nt = 0; uint32_t i = 0;
  while (i < len) {
    uint32_t c = s[i];
    if (is_ws(c)) { i++; continue; }
    if (nt + 2 > cap) return -AIGW_R_TOKENS;
    if (c == '"') {
      tok[nt++] = (uint16_t)i; i++;
      for (;;) {
        if (i >= len) return -AIGW_R_SYNTAX;
        uint32_t ch = s[i];
        if (ch == '"') break;
        if (ch < 0x20) return -AIGW_R_CTRL_IN_STRING;
        if (ch == '\\') {
          if (i + 1 >= len) return -AIGW_R_SYNTAX;
          uint32_t e = s[i + 1];
          if (!(e == '"' || e == '\\' || e == 'n' || e == 'r' || e == 't')) return -AIGW_R_ESCAPE;
          i += 2; continue;
        }
        i++;
      }
      tok[nt++] = (uint16_t)i; i++; continue;
    }
    tok[nt++] = (uint16_t)i;
    if (is_op(c)) { i++; continue; }
    while (i < len && !is_ws(s[i]) && !is_op(s[i]) && s[i] != '"') i++;
  }
  return nt;
}

// ------------------------------------------------------------------ schema walk
struct Walker {
  Doc d;
  Plan pl;
  Scratch sc;
  uint16_t* tok_tail; int tok_tail_cap;  // free token/jump space for nested documents
  uint16_t* jmp_tail;
  const ChatParams* P;
  int reason;  // decline reason

  __device__ __forceinline__ void decline(int r) { if (!reason) reason = r; }
  __device__ __forceinline__ bool bad() const { return reason != 0 || pl.err != 0; }

  // value type helpers (null counts as "absent": zero value)
  __device__ __forceinline__ bool is_str(int v) const { return d.tb(v) == '"'; }
  __device__ __forceinline__ bool is_obj(int v) const { return d.tb(v) == '{'; }
  __device__ __forceinline__ bool is_arr(int v) const { return d.tb(v) == '['; }
  __device__ __forceinline__ bool is_bool(int v) const { uint32_t c = d.tb(v); return c == 't' || c == 'f'; }
  __device__ __forceinline__ bool is_num(int v) const { uint32_t c = d.tb(v); return c == '-' || is_digit(c); }
  __device__ __forceinline__ bool is_null(int v) const { return d.tb(v) == 'n'; }

  __device__ void emit_str(int v, bool sys = false) { pl.src(d, d.tok[v], (uint32_t)d.tok[v + 1] - d.tok[v] + 1u, sys); }

  // returns value token or -1; flags duplicates of that key
  __device__ int find(int obj, const char* key, uint32_t n) {
    int r = -1;
    for (int m = obj + 1; d.tb(m) != '}';) {
      if (d.str_eq(m, key, n)) { if (r >= 0) { decline(AIGW_R_DUP_KEY); return -1; } r = m + 3; }
      int nx = d.next(m + 3);
      if (d.tb(nx) == ',') nx++;
      m = nx;
    }
    if (r >= 0 && is_null(r)) return -1;
    return r;
  }
  // cache_control: object whose "type" == "ephemeral" (anthropic_helper.go:261-263)
  __device__ bool cache_enabled(int obj) {
    int cc = find(obj, "cache_control", 13);
    if (cc < 0) return false;
    if (!is_obj(cc)) { decline(AIGW_R_TYPE); return false; }
    int t = find(cc, "type", 4);
    int ttl = find(cc, "ttl", 3);
    if (ttl >= 0 && !is_str(ttl)) { decline(AIGW_R_TYPE); return false; }
    if (t < 0) return false;
    if (!is_str(t)) { decline(AIGW_R_TYPE); return false; }
    return d.str_eq(t, "ephemeral", 9);
  }
  __device__ void emit_int_field(int v) {
    uint32_t e = d.scalar_end(v), o = d.tok[v];
    if (!is_num(v)) { decline(AIGW_R_TYPE); return; }
    uint32_t l = canon_number(d.s + o, e - o, true);
    if (!l) { decline(AIGW_R_NUMBER); return; }
    pl.src(d, o, l);
  }
  __device__ void emit_float_field(int v) {
    uint32_t e = d.scalar_end(v), o = d.tok[v];
    if (!is_num(v)) { decline(AIGW_R_TYPE); return; }
    uint32_t l = canon_number(d.s + o, e - o, false);
    if (!l) { decline(AIGW_R_NUMBER); return; }
    pl.src(d, o, l);
  }
  __device__ bool check_scalar_type(int v, int kind /*0 str,1 bool,2 int,3 float*/) {
    if (v < 0) return true;
    bool ok;
    if (kind == 0) ok = is_str(v);
    else if (kind == 1) ok = is_bool(v);
    else {
      ok = is_num(v);
      if (ok && kind == 2) { uint32_t e = d.scalar_end(v), o = d.tok[v]; ok = canon_number(d.s + o, e - o, true) != 0; if (!ok) { decline(AIGW_R_NUMBER); return false; } }
    }
    if (!ok) decline(AIGW_R_TYPE);
    return ok;
  }

  // text part list for system / developer / tool messages: [{"text":…,"type":…,"cache_control":…}]
  // emits {"text":S}[,{"cachePoint":…}] per element (cache only when with_cache)
  __device__ void emit_text_parts(int arr, bool with_cache, bool sys, bool& first) {
    for (int e = arr + 1; d.tb(e) != ']';) {
      if (!is_obj(e)) { decline(AIGW_R_CONTENT); return; }
      int t = find(e, "text", 4), ty = find(e, "type", 4);
      if ((t >= 0 && !is_str(t)) || (ty >= 0 && !is_str(ty))) { decline(AIGW_R_TYPE); return; }
      bool cache = cache_enabled(e);
      if (bad()) return;
      if (!first) pl.lit(L_COMMA, sys); first = false;
      pl.lit(L_TEXT_OPEN, sys);
      if (t >= 0) emit_str(t, sys); else pl.lit(L_EMPTY_STR, sys);
      pl.lit(L_RBRACE, sys);
      if (with_cache && cache) { pl.lit(L_COMMA, sys); pl.lit(L_CACHEPOINT, sys); }
      int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx;
    }
  }

  struct Msg { int role_v, content, name, tool_calls, tool_call_id, refusal, audio; };
  // role: 0 user 1 assistant 2 system 3 developer 4 tool
  __device__ int scan_message(int m, Msg& g) {
    g.role_v = g.content = g.name = g.tool_calls = g.tool_call_id = g.refusal = g.audio = -1;
    if (!is_obj(m)) { decline(AIGW_R_ROLE); return -1; }
    uint32_t seen = 0;
    for (int k = m + 1; d.tb(k) != '}';) {
      int v = k + 3;
      uint32_t n = d.str_len(k);
      int which = -1;
      if (n == 4) { if (d.str_eq(k, "role", 4)) which = 0; else if (d.str_eq(k, "name", 4)) which = 2; }
      else if (n == 7) { if (d.str_eq(k, "content", 7)) which = 1; else if (d.str_eq(k, "refusal", 7)) which = 5; }
      else if (n == 10) { if (d.str_eq(k, "tool_calls", 10)) which = 3; }
      else if (n == 12) { if (d.str_eq(k, "tool_call_id", 12)) which = 4; }
      else if (n == 5) { if (d.str_eq(k, "audio", 5)) which = 6; }
      if (which >= 0) {
        if (seen & (1u << which)) { decline(AIGW_R_DUP_KEY); return -1; }
        seen |= 1u << which;
        switch (which) { case 0: g.role_v = v; break; case 1: g.content = v; break; case 2: g.name = v; break; case 3: g.tool_calls = v; break;
          case 4: g.tool_call_id = v; break; case 5: g.refusal = v; break; case 6: g.audio = v; break; }
      }
      int nx = d.next(v); if (d.tb(nx) == ',') nx++; k = nx;
    }
    if (g.role_v < 0 || !is_str(g.role_v)) { decline(AIGW_R_ROLE); return -1; }
    int r = g.role_v;
    if (d.str_eq(r, "user", 4)) return 0;
    if (d.str_eq(r, "assistant", 9)) return 1;
    if (d.str_eq(r, "system", 6)) return 2;
    if (d.str_eq(r, "developer", 9)) return 3;
    if (d.str_eq(r, "tool", 4)) return 4;
    decline(AIGW_R_ROLE); return -1;
  }

  // ---- Bedrock: tool result block for one tool message (openai_awsbedrock.go:451-486)
  __device__ void bedrock_tool_result(const Msg& g) {
    pl.lit(L_TOOLRESULT_OPEN);
    if (g.content < 0) { decline(AIGW_R_CONTENT); return; }   // absent ⇒ 422, null ⇒ 400 in the reference
    if (is_str(g.content)) { pl.lit(L_TEXT_OPEN); emit_str(g.content); pl.lit(L_RBRACE); }
    else if (is_arr(g.content)) { bool first = true; emit_text_parts(g.content, false, false, first); }
    else { decline(AIGW_R_CONTENT); return; }
    pl.lit(L_TOOLRESULT_MID);
    int id = g.tool_call_id;
    if (id >= 0 && is_null(id)) id = -1;
    if (id >= 0) { if (!is_str(id)) { decline(AIGW_R_TYPE); return; } emit_str(id); } else pl.lit(L_EMPTY_STR);
    pl.lit(L_TOOLRESULT_CLOSE);
  }

  // ---- tool call arguments: JSON text inside a JSON string → map[string]any → marshal
  __device__ void emit_arguments(int v) {
    // unescape into scratch
    uint32_t off = d.str_off(v), n = d.str_len(v);
    if (sc.n + n + 1 > sc.cap) { decline(AIGW_R_SCRATCH); return; }
    uint8_t* dst = sc.p + sc.n; uint32_t w = 0;
    const uint8_t* p = d.s + off;
    for (uint32_t i = 0; i < n; i++) {
      uint32_t c = p[i];
      if (c == '\\') { uint32_t e = p[++i]; c = e == 'n' ? '\n' : e == 'r' ? '\r' : e == 't' ? '\t' : e; }
      dst[w++] = (uint8_t)c;
    }
    uint32_t base = sc.n; sc.n += (w + 1u) & ~1u;
    int nt = tokenize_seq(dst, w, tok_tail, tok_tail_cap);
    if (nt <= 0) { decline(nt == 0 ? AIGW_R_ARGS : -nt); return; }
    Doc a; a.s = sc.p; a.len = base + w; a.tok = tok_tail; a.jmp = jmp_tail; a.nt = nt; a.kind = 2;
    // token positions are relative to dst: rebase to scratch origin
    for (int i = 0; i < nt; i++) tok_tail[i] = (uint16_t)(tok_tail[i] + base);
    int r = validate_tokens(a);
    if (r) { decline(AIGW_R_ARGS); return; }
    uint32_t c0 = a.tb(0);
    if (c0 == 'n') { pl.lit(L_NULL); return; }
    if (c0 != '{') { decline(AIGW_R_ARGS); return; }
    emit_any(a, pl, 0);
  }

  // ---- Bedrock assistant content blocks (openai_awsbedrock.go:309-419)
  __device__ void bedrock_asst_part(int e, bool& first) {
    if (!is_obj(e)) { decline(AIGW_R_CONTENT); return; }
    int ty = find(e, "type", 4), text = find(e, "text", 4), refusal = find(e, "refusal", 7), sig = find(e, "signature", 9), red = find(e, "redactedContent", 15);
    if ((ty >= 0 && !is_str(ty)) || (text >= 0 && !is_str(text)) || (refusal >= 0 && !is_str(refusal)) || (sig >= 0 && !is_str(sig))) { decline(AIGW_R_TYPE); return; }
    if (red >= 0) { decline(AIGW_R_UNSUPPORTED_FIELD); return; }
    bool cache = cache_enabled(e);
    if (bad()) return;
    if (ty < 0) return;  // type "" matches no case
    bool emitted = false;
    if (d.str_eq(ty, "text", 4)) { if (text >= 0) { if (!first) pl.lit(L_COMMA); first = false; pl.lit(L_TEXT_OPEN); emit_str(text); pl.lit(L_RBRACE); emitted = true; } }
    else if (d.str_eq(ty, "refusal", 7)) { if (refusal >= 0) { if (!first) pl.lit(L_COMMA); first = false; pl.lit(L_TEXT_OPEN); emit_str(refusal); pl.lit(L_RBRACE); emitted = true; } }
    else if (d.str_eq(ty, "thinking", 8)) {
      if (text >= 0) {
        if (!first) pl.lit(L_COMMA); first = false;
        pl.lit(L_REASON_OPEN); emit_str(text);
        if (sig >= 0 && d.str_len(sig) > 0) { pl.lit(L_REASON_SIG); emit_str(sig); }
        pl.lit(L_REASON_CLOSE); emitted = true;
      }
    } else if (d.str_eq(ty, "redacted_thinking", 17)) { /* RedactedContent nil ⇒ nothing */ }
    if (emitted && cache) { pl.lit(L_COMMA); pl.lit(L_CACHEPOINT); }
  }

  __device__ void bedrock_assistant(const Msg& g) {
    pl.lit(L_MSG_CONTENT_OPEN);
    bool first = true;
    int c = g.content;
    if (c >= 0 && !is_null(c)) {
      if (is_str(c)) { if (d.str_len(c) > 0) { pl.lit(L_TEXT_OPEN); emit_str(c); pl.lit(L_RBRACE); first = false; } }
      else if (is_arr(c)) { for (int e = c + 1; d.tb(e) != ']';) { bedrock_asst_part(e, first); if (bad()) return; int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx; } }
      else if (is_obj(c)) bedrock_asst_part(c, first);
      else { decline(AIGW_R_CONTENT); return; }
    }
    if (g.audio >= 0 && !is_null(g.audio)) { decline(AIGW_R_UNSUPPORTED_FIELD); return; }
    if (g.name >= 0 && !is_null(g.name) && !is_str(g.name)) { decline(AIGW_R_TYPE); return; }
    if (g.refusal >= 0 && !is_null(g.refusal) && !is_str(g.refusal)) { decline(AIGW_R_TYPE); return; }
    int tcs = g.tool_calls;
    if (tcs >= 0 && !is_null(tcs)) {
      if (!is_arr(tcs)) { decline(AIGW_R_TYPE); return; }
      for (int e = tcs + 1; d.tb(e) != ']';) {
        if (!is_obj(e)) { decline(AIGW_R_TOOL); return; }
        int id = find(e, "id", 2), fn = find(e, "function", 8), ty = find(e, "type", 4);
        (void)cache_enabled(e);
        if (bad()) return;
        if (id < 0 || !is_str(id)) { decline(AIGW_R_TOOL); return; }  // nil id panics in the reference
        if (ty >= 0 && !is_str(ty)) { decline(AIGW_R_TYPE); return; }
        int name = -1, args = -1;
        if (fn >= 0) { if (!is_obj(fn)) { decline(AIGW_R_TYPE); return; } name = find(fn, "name", 4); args = find(fn, "arguments", 9); }
        if (bad()) return;
        if ((name >= 0 && !is_str(name)) || (args >= 0 && !is_str(args))) { decline(AIGW_R_TYPE); return; }
        if (args < 0) { decline(AIGW_R_ARGS); return; }  // "" fails to unmarshal in the reference
        if (!first) pl.lit(L_COMMA); first = false;
        pl.lit(L_TOOLUSE_OPEN);
        if (name >= 0) emit_str(name); else pl.lit(L_EMPTY_STR);
        pl.lit(L_TOOLUSE_INPUT);
        emit_arguments(args);
        if (bad()) return;
        pl.lit(L_TOOLUSE_ID); emit_str(id);
        pl.lit(L_TOOLRESULT_CLOSE);
        int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx;
      }
    }
    pl.lit(L_ASST_CLOSE);
  }

  __device__ void bedrock_user(const Msg& g) {
    int c = g.content;
    if (g.name >= 0 && !is_null(g.name) && !is_str(g.name)) { decline(AIGW_R_TYPE); return; }
    if (c < 0) { decline(AIGW_R_CONTENT); return; }  // absent ⇒ 422
    if (is_null(c)) { pl.lit(L_MSG_TEXT_OPEN); pl.lit(L_EMPTY_STR); pl.lit(L_USER_CLOSE1); return; }
    if (is_str(c)) { pl.lit(L_MSG_TEXT_OPEN); emit_str(c); pl.lit(L_USER_CLOSE1); return; }
    if (!is_arr(c)) { decline(AIGW_R_CONTENT); return; }
    pl.lit(L_MSG_CONTENT_OPEN);
    bool first = true;
    for (int e = c + 1; d.tb(e) != ']';) {
      if (!is_obj(e)) { decline(AIGW_R_CONTENT); return; }
      int ty = find(e, "type", 4);
      if (ty < 0 || !is_str(ty) || !d.str_eq(ty, "text", 4)) { decline(AIGW_R_CONTENT); return; }  // images/audio/files: stock path
      int text = find(e, "text", 4);
      if (text >= 0 && !is_str(text)) { decline(AIGW_R_TYPE); return; }
      bool cache = cache_enabled(e);
      if (bad()) return;
      if (!first) pl.lit(L_COMMA); first = false;
      pl.lit(L_TEXT_OPEN); if (text >= 0) emit_str(text); else pl.lit(L_EMPTY_STR); pl.lit(L_RBRACE);
      if (cache) { pl.lit(L_COMMA); pl.lit(L_CACHEPOINT); }
      int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx;
    }
    pl.lit(L_USER_CLOSE);
  }

  __device__ void bedrock_system(const Msg& g, bool& sys_first) {
    int c = g.content;
    if (g.name >= 0 && !is_null(g.name) && !is_str(g.name)) { decline(AIGW_R_TYPE); return; }
    if (c < 0) { decline(AIGW_R_CONTENT); return; }
    if (is_str(c)) { if (!sys_first) pl.lit(L_COMMA, true); sys_first = false; pl.lit(L_TEXT_OPEN, true); emit_str(c, true); pl.lit(L_RBRACE, true); }
    else if (is_arr(c)) emit_text_parts(c, true, true, sys_first);
    else decline(AIGW_R_CONTENT);
  }

  struct Top { int model, messages, temperature, top_p, max_tokens, mct, stop, stream, stream_options, tools, tool_choice, thinking, service_tier; };

  // top-level member scan with type checks for every known field (endpointspec.go:102-105)
  __device__ bool scan_top(Top& t) {
    t.model = t.messages = t.temperature = t.top_p = t.max_tokens = t.mct = t.stop = t.stream = t.stream_options = t.tools = t.tool_choice = t.thinking = t.service_tier = -1;
    if (d.nt == 0 || !is_obj(0)) { decline(AIGW_R_ROOT); return false; }
    uint64_t seen = 0;
    for (int k = 1; d.tb(k) != '}';) {
      int v = k + 3;
      uint32_t n = d.str_len(k);
      const uint8_t* kp = d.s + d.str_off(k);
      int id = -1;  // index into the table below
      // kinds: 0 str, 1 bool, 2 int, 3 float, 8 handled, 9 unsupported-when-present, 10 any
      #define KEY(lit, ident, kindv) if (id < 0 && n == sizeof(lit) - 1 && d.str_eq(k, lit, sizeof(lit) - 1)) { id = ident; kind = kindv; }
      int kind = -1;
      uint32_t c0 = n ? kp[0] : 0;
      switch (c0) {
        case 'm': KEY("model", 0, 8) KEY("messages", 1, 8) KEY("max_tokens", 2, 8) KEY("max_completion_tokens", 3, 8) KEY("modalities", 20, 9) break;
        case 't': KEY("temperature", 4, 8) KEY("top_p", 5, 8) KEY("tools", 6, 8) KEY("tool_choice", 7, 8) KEY("thinking", 8, 8) KEY("top_logprobs", 21, 2) break;
        case 's': KEY("stop", 9, 8) KEY("stream", 10, 8) KEY("stream_options", 11, 8) KEY("service_tier", 12, 8) KEY("seed", 22, 2) KEY("safetySettings", 23, 9) break;
        case 'f': KEY("frequency_penalty", 24, 3) break;
        case 'l': KEY("logit_bias", 25, 9) KEY("logprobs", 26, 1) break;
        case 'n': KEY("n", 27, 2) break;
        case 'p': KEY("presence_penalty", 28, 3) KEY("parallel_tool_calls", 29, 1) KEY("prediction", 30, 9) break;
        case 'r': KEY("response_format", 31, 9) KEY("reasoning_effort", 32, 0) break;
        case 'v': KEY("verbosity", 33, 0) break;
        case 'u': KEY("user", 34, 0) break;
        case 'a': KEY("audio", 35, 9) break;
        case 'w': KEY("web_search_options", 36, 9) break;
        case 'g': KEY("generationConfig", 37, 9) KEY("guided_choice", 38, 9) KEY("guided_regex", 39, 0) KEY("guided_json", 40, 10) break;
        default: break;
      }
      #undef KEY
      if (id >= 0) {
        if (seen & (1ull << id)) { decline(AIGW_R_DUP_KEY); return false; }
        seen |= 1ull << id;
        if (!is_null(v)) {
          if (kind == 9) { decline(AIGW_R_UNSUPPORTED_FIELD); return false; }
          if (kind >= 0 && kind <= 3) { if (!check_scalar_type(v, kind)) return false; }
          if (kind == 8) switch (id) {
            case 0: t.model = v; break; case 1: t.messages = v; break; case 2: t.max_tokens = v; break; case 3: t.mct = v; break;
            case 4: t.temperature = v; break; case 5: t.top_p = v; break; case 6: t.tools = v; break; case 7: t.tool_choice = v; break;
            case 8: t.thinking = v; break; case 9: t.stop = v; break; case 10: t.stream = v; break; case 11: t.stream_options = v; break; case 12: t.service_tier = v; break;
          }
        }
      }
      int nx = d.next(v); if (d.tb(nx) == ',') nx++; k = nx;
    }
    // types of the handled fields
    if (t.model >= 0 && (!is_str(t.model) || d.str_has_backslash(t.model))) { decline(is_str(t.model) ? AIGW_R_ESCAPE : AIGW_R_TYPE); return false; }
    if (t.messages >= 0 && !is_arr(t.messages)) { decline(AIGW_R_TYPE); return false; }
    if (t.stream >= 0 && !is_bool(t.stream)) { decline(AIGW_R_TYPE); return false; }
    if (t.service_tier >= 0 && !is_str(t.service_tier)) { decline(AIGW_R_TYPE); return false; }
    if (t.tools >= 0 && !is_arr(t.tools)) { decline(AIGW_R_TYPE); return false; }
    if (t.stream_options >= 0) {
      if (!is_obj(t.stream_options)) { decline(AIGW_R_TYPE); return false; }
      int iu = find(t.stream_options, "include_usage", 13);
      if (iu >= 0 && !is_bool(iu)) { decline(AIGW_R_TYPE); return false; }
    }
    if (t.temperature >= 0 && !is_num(t.temperature)) { decline(AIGW_R_TYPE); return false; }
    if (t.top_p >= 0 && !is_num(t.top_p)) { decline(AIGW_R_TYPE); return false; }
    if (t.max_tokens >= 0 && !check_scalar_type(t.max_tokens, 2)) return false;
    if (t.mct >= 0 && !check_scalar_type(t.mct, 2)) return false;
    return !bad();
  }

  __device__ bool model_contains(int mv, const char* needle, uint32_t n) {
    if (mv < 0) return false;
    const uint8_t* p = d.s + d.str_off(mv); uint32_t L = d.str_len(mv);
    for (uint32_t i = 0; i + n <= L; i++) { uint32_t k = 0; while (k < n && p[i + k] == (uint8_t)needle[k]) k++; if (k == n) return true; }
    return false;
  }

  // ":path" = /model/{url.PathEscape(model)}/converse[-stream]  (openai_awsbedrock.go:94-109,155)
  __device__ void emit_bedrock_path(const Top& t, bool stream) {
    pl.lit(L_PATH_MODEL);
    const uint8_t* mp; uint32_t ml;
    if (P->override_len) { mp = (const uint8_t*)P->override_model; ml = P->override_len; }
    else if (t.model >= 0) { mp = d.s + d.str_off(t.model); ml = d.str_len(t.model); }
    else { mp = nullptr; ml = 0; }
    bool clean = true;
    for (uint32_t i = 0; i < ml; i++) {
      uint32_t c = mp[i];
      bool keep = (c - 'a' < 26u) || (c - 'A' < 26u) || (c - '0' < 10u) || c == '-' || c == '_' || c == '.' || c == '~' || c == '$' || c == '&' || c == '+' || c == ':' || c == '=' || c == '@';
      if (!keep) { clean = false; break; }
    }
    if (clean && !P->override_len) { if (ml) pl.src(d, d.str_off(t.model), ml); }
    else {
      if (sc.n + 3 * ml > sc.cap) { decline(AIGW_R_SCRATCH); return; }
      uint8_t* o = sc.p + sc.n; uint32_t w = 0;
      for (uint32_t i = 0; i < ml; i++) {
        uint32_t c = mp[i];
        bool keep = (c - 'a' < 26u) || (c - 'A' < 26u) || (c - '0' < 10u) || c == '-' || c == '_' || c == '.' || c == '~' || c == '$' || c == '&' || c == '+' || c == ':' || c == '=' || c == '@';
        if (keep) o[w++] = (uint8_t)c;
        else { const char* hx = "0123456789ABCDEF"; o[w++] = '%'; o[w++] = hx[c >> 4]; o[w++] = hx[c & 15]; }
      }
      pl.push(2, sc.n, w); sc.n += (w + 1u) & ~1u;
    }
    pl.lit(L_PATH_CONVERSE);
    if (stream) pl.lit(L_PATH_STREAM);
  }

  // ---- OpenAI → AWS Bedrock Converse (openai_awsbedrock.go:91-159)
  __device__ void plan_bedrock(const Top& t, bool stream, uint32_t& path_len) {
    emit_bedrock_path(t, stream);
    path_len = pl.olen;
    if (bad()) return;
    pl.lit(L_LBRACE);
    if (t.thinking >= 0) {  // openai.go:911-945, openai_awsbedrock.go:57-78
      int th = t.thinking;
      if (!is_obj(th)) { decline(AIGW_R_TYPE); return; }
      int ty = find(th, "type", 4);
      if (ty < 0 || !is_str(ty)) { decline(AIGW_R_TYPE); return; }
      if (d.str_eq(ty, "enabled", 7)) {
        int bt = find(th, "budget_tokens", 13), it = find(th, "includeThoughts", 15);
        if (it >= 0 && !is_bool(it)) { decline(AIGW_R_TYPE); return; }
        pl.lit(L_ADDL_EN_PRE);
        if (bt >= 0) emit_int_field(bt); else pl.lit(L_ZERO);
        pl.lit(L_ADDL_EN_POST);
      } else if (d.str_eq(ty, "disabled", 8)) pl.lit(L_ADDL_DIS);
      else if (d.str_eq(ty, "adaptive", 8)) {}
      else { decline(AIGW_R_TYPE); return; }
      if (bad()) return;
    }
    pl.lit(L_INF_OPEN);
    bool f = true;
    int mt = t.mct >= 0 ? t.mct : t.max_tokens;  // cmp.Or(MaxCompletionTokens, MaxTokens)
    if (mt >= 0) { pl.lit(L_MAXTOK); emit_int_field(mt); f = false; }
    if (t.stop >= 0) {
      int s = t.stop;
      if (is_str(s)) { if (!f) pl.lit(L_COMMA); f = false; pl.lit(L_STOPSEQ); emit_str(s); pl.lit(L_RBRACK); }
      else if (is_arr(s)) {
        if (d.tb(s + 1) != ']') {
          if (!f) pl.lit(L_COMMA); f = false; pl.lit(L_STOPSEQ);
          bool sf = true;
          for (int e = s + 1; d.tb(e) != ']';) { if (!is_str(e)) { decline(AIGW_R_TYPE); return; } if (!sf) pl.lit(L_COMMA); sf = false; emit_str(e); int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx; }
          pl.lit(L_RBRACK);
        }
      } else { decline(AIGW_R_TYPE); return; }
    }
    if (t.temperature >= 0) { if (!f) pl.lit(L_COMMA); f = false; pl.lit(L_TEMP); emit_float_field(t.temperature); }
    if (t.top_p >= 0) { if (!f) pl.lit(L_COMMA); f = false; pl.lit(L_TOPP); emit_float_field(t.top_p); }
    pl.lit(L_INF_CLOSE_MSGS);
    if (bad()) return;
    // messages (openai_awsbedrock.go:489-585)
    bool mfirst = true, sys_first = true;
    if (t.messages >= 0) {
      int e = t.messages + 1;
      while (d.tb(e) != ']') {
        Msg g; int role = scan_message(e, g);
        if (bad()) return;
        int nx = d.next(e); if (d.tb(nx) == ',') nx++;
        if (role == 2 || role == 3) { bedrock_system(g, sys_first); e = nx; if (bad()) return; continue; }
        if (!mfirst) pl.lit(L_COMMA); mfirst = false;
        if (role == 0) bedrock_user(g);
        else if (role == 1) bedrock_assistant(g);
        else {
          pl.lit(L_MSG_CONTENT_OPEN);
          bedrock_tool_result(g);
          if (bad()) return;
          // coalesce the following tool messages (openai_awsbedrock.go:559-575)
          while (d.tb(nx) != ']') {
            Msg g2; int r2 = scan_message(nx, g2);
            if (bad()) return;
            if (r2 != 4) break;
            pl.lit(L_COMMA); bedrock_tool_result(g2);
            if (bad()) return;
            int n2 = d.next(nx); if (d.tb(n2) == ',') n2++; nx = n2;
          }
          pl.lit(L_USER_CLOSE);
        }
        if (bad()) return;
        e = nx;
      }
    }
    pl.lit(L_RBRACK);
    if (pl.nsys) { pl.lit(L_SYSTEM_OPEN); pl.flush_sys(); pl.lit(L_RBRACK); }
    if (t.service_tier >= 0 && d.str_len(t.service_tier) > 0) { pl.lit(L_SERVICE_TIER); emit_str(t.service_tier); pl.lit(L_RBRACE); }
    // tool_choice decodes (and can fail) whether or not tools are present (openai.go:1194-1210)
    int tc_kind = 0, tc_name = -1;  // 1 auto, 2 any, 3 tool{name=tc_name or ""}
    if (t.tool_choice >= 0) {
      const int tc = t.tool_choice;
      if (is_str(tc)) {
        if (d.str_eq(tc, "auto", 4)) tc_kind = 1;
        else if (d.str_eq(tc, "required", 8)) tc_kind = 2;
        else if (model_contains(t.model, "anthropic", 9) && model_contains(t.model, "claude", 6)) { tc_kind = 3; tc_name = tc; }
      } else if (is_obj(tc)) {
        const int ty = find(tc, "type", 4), fn = find(tc, "function", 8);
        if (ty >= 0 && !is_str(ty)) { decline(AIGW_R_TYPE); return; }
        if (fn >= 0) { if (!is_obj(fn)) { decline(AIGW_R_TYPE); return; } tc_name = find(fn, "name", 4); if (tc_name >= 0 && !is_str(tc_name)) { decline(AIGW_R_TYPE); return; } }
        if (bad()) return;
        tc_kind = 3;
      } else { decline(AIGW_R_TYPE); return; }
    }
    // tools (openai_awsbedrock.go:162-226)
    if (t.tools >= 0 && d.tb(t.tools + 1) != ']') {
      pl.lit(L_TOOLCFG_OPEN);
      if (tc_kind == 1) pl.lit(L_TOOLCHOICE_AUTO);
      else if (tc_kind == 2) pl.lit(L_TOOLCHOICE_ANY);
      else if (tc_kind == 3) { pl.lit(L_TOOLCHOICE_TOOL); if (tc_name >= 0) emit_str(tc_name); else pl.lit(L_EMPTY_STR); pl.lit(L_TOOLCHOICE_TOOL_END); }
      pl.lit(L_TOOLS_OPEN);
      bool tf = true;
      for (int e = t.tools + 1; d.tb(e) != ']';) {
        if (!is_obj(e)) { decline(AIGW_R_TOOL); return; }
        int ty = find(e, "type", 4), fn = find(e, "function", 8), gs = find(e, "google_search", 13);
        if (ty >= 0 && !is_str(ty)) { decline(AIGW_R_TYPE); return; }
        if (gs >= 0) { decline(AIGW_R_UNSUPPORTED_FIELD); return; }
        if (bad()) return;
        if (fn >= 0) {
          if (!is_obj(fn)) { decline(AIGW_R_TYPE); return; }
          int nm = find(fn, "name", 4), ds = find(fn, "description", 11), st = find(fn, "strict", 6), pr = find(fn, "parameters", 10);
          bool cache = cache_enabled(fn);
          if (bad()) return;
          if ((nm >= 0 && !is_str(nm)) || (ds >= 0 && !is_str(ds)) || (st >= 0 && !is_bool(st))) { decline(AIGW_R_TYPE); return; }
          if (!tf) pl.lit(L_COMMA); tf = false;
          pl.lit(L_TOOLSPEC_OPEN);
          if (ds >= 0 && d.str_len(ds) > 0) { pl.lit(L_DESC); emit_str(ds); pl.lit(L_COMMA); }
          pl.lit(L_INPUTSCHEMA);
          if (pr >= 0) emit_any(d, pl, pr); else pl.lit(L_NULL);
          if (bad()) return;
          pl.lit(L_NAME); if (nm >= 0) emit_str(nm); else pl.lit(L_EMPTY_STR); pl.lit(L_RBRACE);
          if (cache) pl.lit(L_TOOL_CACHE);
          pl.lit(L_RBRACE);
        }
        int nx = d.next(e); if (d.tb(nx) == ',') nx++; e = nx;
      }
      pl.lit(L_TOOLS_CLOSE);
    }
    pl.lit(L_RBRACE);
  }
};

// ------------------------------------------------------------------ the kernel
template <int MAXD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) chat_translate_kernel(const __grid_constant__ ChatParams P) {
  using C = Cls<MAXD>;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* s_lits = smem;  // literal bytes, shared by the CTA
  for (uint32_t i = threadIdx.x; i < sizeof(c_lits.bytes) / 4; i += blockDim.x) ((uint32_t*)s_lits)[i] = ((const uint32_t*)c_lits.bytes)[i];
  __syncthreads();
  uint8_t* wb = smem + sizeof(c_lits.bytes) + (size_t)warp * C::kWarpBytes;
  uint8_t* s_in = wb;
  uint8_t* s_out = wb + C::kIn;
  uint16_t* s_tok = (uint16_t*)s_out;
  uint16_t* s_jmp = s_tok + C::kTokCap;
  uint32_t* s_ops = (uint32_t*)(wb + C::kIn + C::kOut);
  uint8_t* s_scr = wb + C::kIn + C::kOut + C::kOpCap * 4;

  for (;;) {
    uint32_t doc = 0;
    if (lane == 0) doc = atomicAdd(P.next_doc, 1u);
    doc = __shfl_sync(FULL, doc, 0);
    if (doc >= P.n) break;
    const uint32_t len = P.lens[doc];
    const uint8_t* g = P.bodies + P.offsets[doc];
    aigw_doc_result res;
    res.out_off = 0; res.body_len = 0; res.path_len = 0; res.status = AIGW_DECLINED; res.reason = AIGW_R_NONE;
    res.model_off = 0; res.model_len = 0; res.body_kind = AIGW_BODY_UNCHANGED; res.flags = 0; res.in_len = len; res._pad = 0;
    if (len > (uint32_t)MAXD || len == 0) {
      res.reason = len ? AIGW_R_TOO_LARGE : AIGW_R_SYNTAX;
      if (lane == 0) P.results[doc] = res;
      continue;
    }
    // ---- stage 1: load (16-byte coalesced), pad the last round with spaces
    const uint32_t rounds = (len + 1023u) >> 10;
    {
      const uint4* g4 = (const uint4*)g;
      uint4* s4 = (uint4*)s_in;
      const uint32_t n16 = (len + 15u) >> 4;
      for (uint32_t i = lane; i < n16; i += 32) s4[i] = __ldg(g4 + i);
      __syncwarp();
      const uint32_t padded = rounds << 10;
      for (uint32_t i = len + lane; i < padded; i += 32) s_in[i] = ' ';
      __syncwarp();
    }
    // ---- stage 2: structural index
    uint32_t carry_esc = 0, carry_str = 0, carry_sc = 0;
    uint32_t ntok = 0;
    uint32_t flags = 0;  // bit0 ctrl in string, bit1 bad escape, bit2 token overflow
    for (uint32_t r = 0; r < rounds; r++) {
      const uint32_t base = (r << 10) + (lane << 5);
      const uint4 a = *(const uint4*)(s_in + base), b = *(const uint4*)(s_in + base + 16);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t mq = 0, mb = 0, mctl = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        mq |= nib_from_ff(__vcmpeq4(w[j], 0x22222222u)) << (4 * j);
        mb |= nib_from_ff(__vcmpeq4(w[j], 0x5c5c5c5cu)) << (4 * j);
        mctl |= nib_from_ff(__vcmpltu4(w[j], 0x20202020u)) << (4 * j);
      }
      // escaped characters: odd-length backslash runs, carry across lanes
      const uint32_t tail = __clz(~mb);  // backslash run length at the top of this lane's 32 bytes
      const uint32_t odd = __ballot_sync(FULL, tail < 32u && (tail & 1u));
      const uint32_t full = __ballot_sync(FULL, tail == 32u);
      uint32_t cin;
      {
        const uint32_t below = ~full & ((1u << lane) - 1u);
        if (below == 0) cin = carry_esc; else cin = (odd >> (31 - __clz(below))) & 1u;
      }
      {
        // carry out of the round = carry into a virtual lane 32
        const uint32_t below = ~full;
        carry_esc = below == 0 ? carry_esc : (odd >> (31 - __clz(below))) & 1u;
      }
      uint32_t esc;
      {
        const uint32_t bs = mb & ~cin;  // a leading backslash that is itself escaped starts no run
        const uint32_t follows = (bs << 1) | cin;
        const uint32_t even = 0x55555555u;
        const uint32_t odd_starts = bs & ~even & ~follows;
        const uint32_t seq_even = odd_starts + bs;  // carry-out handled through `tail`
        const uint32_t invert = seq_even << 1;
        esc = (even ^ invert) & follows;
      }
      const uint32_t uq = mq & ~esc;
      uint32_t ps = uq;
      ps ^= ps << 1; ps ^= ps << 2; ps ^= ps << 4; ps ^= ps << 8; ps ^= ps << 16;
      const uint32_t par = __ballot_sync(FULL, __popc(uq) & 1);
      const uint32_t sin = (__popc(par & ((1u << lane) - 1u)) & 1u) ^ carry_str;
      if (sin) ps = ~ps;
      carry_str ^= __popc(par) & 1u;
      // ps: bit set from an opening quote up to the byte before its closing quote
      if (mctl & ps) flags |= 1u;
      {
        uint32_t e = esc & ps;
        while (e) { const int j = __ffs(e) - 1; e &= e - 1; const uint32_t c = s_in[base + j]; if (!(c == '"' || c == '\\' || c == 'n' || c == 'r' || c == 't')) flags |= 2u; }
      }
      // bytes outside strings: classify sparsely
      uint32_t mtok = uq, msc = 0;
      {
        uint32_t o = ~ps & ~uq;
        while (o) {
          const int j = __ffs(o) - 1; o &= o - 1;
          const uint32_t c = s_in[base + j];
          if (is_ws(c)) continue;
          if (is_op(c)) mtok |= 1u << j; else msc |= 1u << j;
        }
      }
      {
        uint32_t prev = __shfl_up_sync(FULL, msc >> 31, 1);
        if (lane == 0) prev = carry_sc;
        carry_sc = __shfl_sync(FULL, msc >> 31, 31);
        mtok |= msc & ~((msc << 1) | prev);
      }
      // compact token positions
      const uint32_t cnt = __popc(mtok);
      uint32_t incl = cnt;
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, sft); if (lane >= sft) incl += v; }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      uint32_t wpos = ntok + incl - cnt;
      if (ntok + total > (uint32_t)C::kTokCap) flags |= 4u;
      else { uint32_t m = mtok; while (m) { const int j = __ffs(m) - 1; m &= m - 1; s_tok[wpos++] = (uint16_t)(base + j); } }
      ntok += total;
    }
    flags = __reduce_or_sync(FULL, flags);
    __syncwarp();
    int reason = 0;
    if (flags & 4u) reason = AIGW_R_TOKENS;
    else if (flags & 1u) reason = AIGW_R_CTRL_IN_STRING;
    else if (flags & 2u) reason = AIGW_R_ESCAPE;
    else if (carry_str) reason = AIGW_R_SYNTAX;
    // ---- stages 3+4 on lane 0
    uint32_t nops = 0, olen = 0, path_len = 0, model_off = 0, model_len = 0, rflags = 0, body_kind = AIGW_BODY_BYTES;
    if (lane == 0 && !reason) {
      Walker W;
      W.d.s = s_in; W.d.len = len; W.d.tok = s_tok; W.d.jmp = s_jmp; W.d.nt = (int)ntok; W.d.kind = 0;
      W.pl.ops = s_ops; W.pl.nops = 0; W.pl.nsys = 0; W.pl.cap = C::kOpCap - kSysCap; W.pl.olen = 0; W.pl.err = 0;
      W.sc.p = s_scr; W.sc.n = 0; W.sc.cap = C::kScr;
      W.tok_tail = s_tok + ntok; W.jmp_tail = s_jmp + ntok; W.tok_tail_cap = C::kTokCap - (int)ntok;
      W.P = &P; W.reason = 0;
      reason = validate_tokens(W.d);
      if (!reason) {
        Walker::Top t;
        if (W.scan_top(t)) {
          const bool stream = t.stream >= 0 && W.d.tb(t.stream) == 't';
          if (t.model >= 0) { model_off = W.d.str_off(t.model); model_len = W.d.str_len(t.model); }
          rflags = stream ? 1u : 0u;
          if (P.schema == AIGW_SCHEMA_AWS_BEDROCK) W.plan_bedrock(t, stream, path_len);
          else W.decline(AIGW_R_SCHEMA);
        }
        reason = W.reason ? W.reason : W.pl.err;
        nops = W.pl.nops; olen = W.pl.olen;
        if (!reason && olen > (uint32_t)C::kOut) reason = AIGW_R_OUT_SPACE;
      }
    }
    reason = __shfl_sync(FULL, reason, 0);
    if (reason) {
      res.reason = (uint8_t)reason;
      if (lane == 0) P.results[doc] = res;
      __syncwarp();
      continue;
    }
    nops = __shfl_sync(FULL, nops, 0); olen = __shfl_sync(FULL, olen, 0);
    __syncwarp();
    // ---- stage 5: emit.  tok/jmp (aliasing s_out) are dead from here on.
    {
      uint32_t dst = 0;
      const uint8_t* lits = s_lits;
      for (uint32_t k = 0; k < nops; k++) {
        const uint32_t op = s_ops[k];
        const uint32_t kind = op >> 30, l = (op >> 16) & 0x3fffu, off = op & 0xffffu;
        if (kind == 1) { for (uint32_t i = lane; i < l; i += 32) s_out[dst + i] = lits[off + i]; }
        else { const uint8_t* sp = (kind == 0 ? s_in : s_scr) + off; for (uint32_t i = lane; i < l; i += 32) s_out[dst + i] = sp[i]; }
        dst += l;
      }
    }
    __syncwarp();
    const uint32_t rec = (olen + 15u) & ~15u;
    unsigned long long obase = 0;
    if (lane == 0) obase = atomicAdd(P.out_used, (unsigned long long)rec);
    obase = __shfl_sync(FULL, obase, 0);
    if (obase + rec > P.out_capacity) {
      res.reason = AIGW_R_ARENA_FULL;
      if (lane == 0) P.results[doc] = res;
      continue;
    }
    {
      uint4* o4 = (uint4*)(P.out + obase);
      const uint4* s4 = (const uint4*)s_out;
      for (uint32_t i = lane; i < (rec >> 4); i += 32) o4[i] = s4[i];
    }
    if (lane == 0) {
      res.out_off = obase + P.out_bias; res.body_len = olen - path_len; res.path_len = (uint16_t)path_len; res.status = AIGW_OK; res.reason = 0;
      res.model_off = model_off; res.model_len = (uint16_t)model_len; res.body_kind = (uint8_t)body_kind; res.flags = (uint8_t)rflags;
      P.results[doc] = res;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ host-side launcher
template <int MAXD, int WARPS>
static cudaError_t launch_cls(const ChatParams& P, int sm_count, cudaStream_t st, int* blocks_out) {
  using C = Cls<MAXD>;
  const size_t smem = sizeof(LitTable::bytes) + (size_t)C::kWarpBytes * WARPS;
  static bool configured = false;
  static int blocks_per_sm = 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(chat_translate_kernel<MAXD, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, chat_translate_kernel<MAXD, WARPS>, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  long long want = ((long long)P.n + WARPS - 1) / WARPS;
  long long grid = (long long)sm_count * blocks_per_sm;  // persistent: a multiple of the SM count
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  if (blocks_out) *blocks_out = (int)grid;
  chat_translate_kernel<MAXD, WARPS><<<(unsigned)grid, WARPS * 32, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_chat_translate(const ChatParams& P, uint32_t max_len, int sm_count, cudaStream_t st) {
  if (max_len <= 2048) return launch_cls<2048, 2>(P, sm_count, st, nullptr);
  if (max_len <= 5120) return launch_cls<5120, 2>(P, sm_count, st, nullptr);
  if (max_len <= 9216) return launch_cls<9216, 2>(P, sm_count, st, nullptr);
  if (max_len <= 17408) return launch_cls<17408, 1>(P, sm_count, st, nullptr);
  if (max_len <= 33792) return launch_cls<33792, 1>(P, sm_count, st, nullptr);
  return launch_cls<65536, 1>(P, sm_count, st, nullptr);
}

}  // namespace aigw
