// See cel_kernel.cuh.  Host: recursive-descent compiler with type checking; device: one thread per (record, program).
#include <cctype>
#include <cstring>

#include "cel_kernel.cuh"

namespace aigw {
namespace {

enum Ty { T_INT, T_UINT, T_BOOL, T_STR };
struct R { bool ok; Ty t; uint8_t skind; uint32_t soff, slen; };   // skind for strings: 0 literal, 1 backend, 2 route, 3 model

struct Compiler {
  const char* s; size_t n, i = 0; std::string err; CelProgramHost* P; int depth = 0, max_depth = 0;
  void ws() { while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
  bool at(const char* w) { const size_t l = strlen(w); return i + l <= n && memcmp(s + i, w, l) == 0; }
  R fail(const std::string& m) { if (err.empty()) err = m; return R{false, T_INT, 0, 0, 0}; }
  static bool num(Ty t) { return t == T_INT || t == T_UINT; }
  void emit(uint8_t op, uint8_t a = 0, uint64_t imm = 0, uint16_t tgt = 0) { CelInstr in{}; in.op = op; in.a = a; in.tgt = tgt; in.imm = imm; P->code.push_back(in); }
  void push() { if (++depth > max_depth) max_depth = depth; }
  void pop(int k = 1) { depth -= k; }

  R expr() {
    R c = lor(); if (!c.ok) return c;
    ws();
    if (i < n && s[i] == '?') {
      i++;
      if (c.t != T_BOOL) return fail("no matching overload for _?_:_");
      const size_t jf = P->code.size(); emit(C_JMP_FALSE); pop();
      R a = lor(); if (!a.ok) return a;
      if (a.t == T_STR) return fail("unsupported: string-valued conditional");
      const size_t je = P->code.size(); emit(C_JMP); pop();
      ws(); if (i >= n || s[i] != ':') return fail("expected ':'"); i++;
      P->code[jf].tgt = (uint16_t)P->code.size();
      R b = expr(); if (!b.ok) return b;
      if (a.t != b.t) return fail("no matching overload for _?_:_");
      P->code[je].tgt = (uint16_t)P->code.size(); P->code[jf].tgt2 = (uint16_t)P->code.size();
      return R{true, a.t, 0, 0, 0};
    }
    return c;
  }
  R lor() { R l = land(); if (!l.ok) return l; for (;;) { ws(); if (!at("||")) return l; i += 2; R r = land(); if (!r.ok) return r; if (l.t != T_BOOL || r.t != T_BOOL) return fail("no matching overload for _||_"); emit(C_OR); pop(); } }
  R land() { R l = rel(); if (!l.ok) return l; for (;;) { ws(); if (!at("&&")) return l; i += 2; R r = rel(); if (!r.ok) return r; if (l.t != T_BOOL || r.t != T_BOOL) return fail("no matching overload for _&&_"); emit(C_AND); pop(); } }
  R rel() {
    R l = add(); if (!l.ok) return l;
    for (;;) {
      ws(); int op = 0;
      if (at("==")) { op = 'e'; i += 2; } else if (at("!=")) { op = 'n'; i += 2; } else if (at("<=")) { op = 'l'; i += 2; } else if (at(">=")) { op = 'g'; i += 2; }
      else if (i < n && s[i] == '<') { op = '<'; i++; } else if (i < n && s[i] == '>') { op = '>'; i++; } else return l;
      R r = add(); if (!r.ok) return r;
      if (l.t != r.t) return fail("no matching overload for comparison");
      if (l.t == T_STR) {
        if (op != 'e' && op != 'n') return fail("unsupported: ordering of strings");
        const bool neg = op == 'n';
        if (l.skind == 3 && r.skind == 3) emit(C_PUSH_IMM, 0, neg ? 0 : 1);
        else if (l.skind == 3 || r.skind == 3) { const R& o = l.skind == 3 ? r : l; emit(C_STREQ_MODEL, o.skind, ((uint64_t)o.soff << 32) | o.slen, neg ? 1 : 0); }
        else {
          if (P->sites.size() >= 32) return fail("unsupported: more than 32 string comparisons");
          emit(C_PUSH_HOSTBIT, (uint8_t)P->sites.size());
          P->sites.push_back(CelHostSite{l.skind, r.skind, l.soff, l.slen, r.soff, r.slen, neg});
        }
        push();
      } else if (l.t == T_BOOL) { if (op != 'e' && op != 'n') return fail("unsupported: ordering of bools"); emit(op == 'e' ? C_EQ : C_NE); pop(); }
      else {
        const bool I = l.t == T_INT;
        emit(op == 'e' ? C_EQ : op == 'n' ? C_NE : op == '<' ? (I ? C_LT_I : C_LT_U) : op == 'l' ? (I ? C_LE_I : C_LE_U) : op == '>' ? (I ? C_GT_I : C_GT_U) : (I ? C_GE_I : C_GE_U)); pop();
      }
      l = R{true, T_BOOL, 0, 0, 0};
    }
  }
  R add() { R l = mul(); if (!l.ok) return l; for (;;) { ws(); if (i >= n || (s[i] != '+' && s[i] != '-')) return l; const char op = s[i++]; R r = mul(); if (!r.ok) return r;
      if (l.t != r.t || !num(l.t)) return fail("no matching overload for arithmetic"); emit(op == '+' ? (l.t == T_INT ? C_ADD_I : C_ADD_U) : (l.t == T_INT ? C_SUB_I : C_SUB_U)); pop(); } }
  R mul() { R l = unary(); if (!l.ok) return l; for (;;) { ws(); if (i >= n || (s[i] != '*' && s[i] != '/' && s[i] != '%')) return l; const char op = s[i++]; R r = unary(); if (!r.ok) return r;
      if (l.t != r.t || !num(l.t)) return fail("no matching overload for arithmetic");
      emit(op == '*' ? (l.t == T_INT ? C_MUL_I : C_MUL_U) : op == '/' ? (l.t == T_INT ? C_DIV_I : C_DIV_U) : (l.t == T_INT ? C_MOD_I : C_MOD_U)); pop(); } }
  R unary() {
    ws();
    if (i < n && s[i] == '!') { i++; R a = unary(); if (!a.ok) return a; if (a.t != T_BOOL) return fail("no matching overload for !_"); emit(C_NOT); return a; }
    if (i < n && s[i] == '-') { i++; R a = unary(); if (!a.ok) return a; if (a.t != T_INT) return fail("no matching overload for -_"); emit(C_NEG_I); return a; }
    return primary();
  }
  R primary() {
    ws();
    if (i >= n) return fail("unexpected end of expression");
    const char c = s[i];
    if (c == '(') { i++; R e = expr(); if (!e.ok) return e; ws(); if (i >= n || s[i] != ')') return fail("expected ')'"); i++; return e; }
    if (c == '\'' || c == '"') {
      size_t j = i + 1; while (j < n && s[j] != c) { if (s[j] == '\\' || s[j] == '\n') return fail("unsupported: escapes in string literal"); j++; }
      if (j >= n) return fail("unterminated string");
      if (j == i + 1 && j + 1 < n && s[j + 1] == c) return fail("unsupported: triple-quoted string");
      const uint32_t off = (uint32_t)P->strings.size(), len = (uint32_t)(j - i - 1);
      P->strings.append(s + i + 1, len); i = j + 1;
      return R{true, T_STR, 0, off, len};
    }
    if (c >= '0' && c <= '9') {
      uint64_t v = 0; bool ovf = false; size_t j = i;
      if (at("0x") || at("0X")) { j = i + 2; const size_t d0 = j; while (j < n && isxdigit((unsigned char)s[j])) { const int d = isdigit((unsigned char)s[j]) ? s[j] - '0' : (tolower(s[j]) - 'a' + 10); if (v >> 60) ovf = true; v = v * 16 + (uint64_t)d; j++; } if (j == d0) return fail("bad hex literal"); }
      else while (j < n && s[j] >= '0' && s[j] <= '9') { const uint64_t d = (uint64_t)(s[j] - '0'); if (v > (~0ull - d) / 10) ovf = true; v = v * 10 + d; j++; }
      if (j < n && (s[j] == '.' || s[j] == 'e' || s[j] == 'E')) return fail("unsupported: double literal");
      Ty t;
      if (j < n && (s[j] == 'u' || s[j] == 'U')) { if (ovf) return fail("uint literal out of range"); t = T_UINT; j++; }
      else { if (ovf || v > 0x7fffffffffffffffull) return fail("int literal out of range"); t = T_INT; }
      i = j; emit(C_PUSH_IMM, 0, v); push();
      return R{true, t, 0, 0, 0};
    }
    if (isalpha((unsigned char)c) || c == '_') {
      size_t j = i; while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_')) j++;
      const std::string id(s + i, j - i); i = j; ws();
      if (id == "true" || id == "false") { emit(C_PUSH_IMM, 0, id == "true"); push(); return R{true, T_BOOL, 0, 0, 0}; }
      if (i < n && s[i] == '(') {
        if (id != "uint" && id != "int") return fail("unsupported: function " + id);
        i++; R a = expr(); if (!a.ok) return a; ws(); if (i >= n || s[i] != ')') return fail("expected ')'"); i++;
        if (!num(a.t)) return fail("unsupported: conversion from non-integer");
        const Ty to = id == "uint" ? T_UINT : T_INT;
        if (to != a.t) emit(to == T_UINT ? C_TO_UINT : C_TO_INT);
        return R{true, to, 0, 0, 0};
      }
      if (i < n && (s[i] == '.' || s[i] == '[')) return fail("unsupported: member access / index");
      static const char* toks[6] = {"input_tokens", "cached_input_tokens", "cache_creation_input_tokens", "output_tokens", "total_tokens", "reasoning_tokens"};
      for (int k = 0; k < 6; k++) if (id == toks[k]) { emit(C_PUSH_TOK, (uint8_t)k); push(); return R{true, T_UINT, 0, 0, 0}; }
      if (id == "model") return R{true, T_STR, 3, 0, 0};
      if (id == "backend") return R{true, T_STR, 1, 0, 0};
      if (id == "route_name") return R{true, T_STR, 2, 0, 0};
      return fail("undeclared reference to '" + id + "'");
    }
    return fail("unexpected character");
  }
};

__global__ void cel_kernel(const __grid_constant__ CelLaunch L) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.n * L.n_progs) return;
  const uint32_t rec = t / L.n_progs, p = t % L.n_progs;
  const aigw_usage& u = L.results[rec].usage;
  const uint32_t tok[6] = {u.input, u.cached, u.cache_creation, u.output, u.total, u.reasoning};
  const uint8_t* model; uint32_t mlen;
  if (L.model_off) { model = L.model_bytes + L.model_off[rec]; mlen = L.model_len[rec]; } else { model = (const uint8_t*)L.cs.model; mlen = L.cs.model_len; }
  uint64_t v = 0;
  const uint8_t e = cel_run(L.code[p], L.strings[p], tok, model, mlen, &L.cs, L.host_bits[p], L.result_is_int[p] != 0, &v);
  L.costs[t] = v; L.errs[t] = e;
}

}  // namespace

int cel_compile(const char* expr, CelProgramHost& out, std::string& err) {
  Compiler c; c.s = expr; c.n = strlen(expr); c.P = &out;
  out.code.clear(); out.strings.clear(); out.sites.clear();
  R r = c.expr();
  if (r.ok) { c.ws(); if (c.i != c.n) r = c.fail("unexpected trailing input"); }
  if (!r.ok) { err = c.err; return 1; }
  if (r.t == T_STR) { err = "CEL expression result is not an integer"; return 2; }
  if (r.t == T_BOOL) { err = "CEL expression result is not an integer"; return 2; }
  c.emit(C_END);
  if (out.code.size() > (size_t)kCelMaxInstr || c.max_depth > kCelMaxStack) { err = "unsupported: expression too large"; return 1; }
  out.result_is_int = r.t == T_INT;
  // NewProgram's sanity evaluation: "dummy" strings, zero counters (internal/llmcostcel/cel.go:66-70)
  CelStrings cs; memset(&cs, 0, sizeof cs); memcpy(cs.backend, "dummy", 5); memcpy(cs.route, "dummy", 5); memcpy(cs.model, "dummy", 5); cs.backend_len = cs.route_len = cs.model_len = 5;
  uint32_t bits = 0;
  for (size_t k = 0; k < out.sites.size(); k++) {
    const CelHostSite& st = out.sites[k];
    auto str = [&](uint8_t kind, uint32_t off, uint32_t len) { return kind == 0 ? out.strings.substr(off, len) : std::string("dummy"); };
    const bool eq = str(st.lk, st.loff, st.llen) == str(st.rk, st.roff, st.rlen);
    if (eq != st.negate) bits |= 1u << k;
  }
  const uint32_t zero[6] = {0, 0, 0, 0, 0, 0}; uint64_t v = 0;
  std::vector<CelInstr> padded(out.code); padded.resize(kCelMaxInstr, CelInstr{});
  if (cel_run(padded.data(), (const uint8_t*)out.strings.data(), zero, (const uint8_t*)cs.model, cs.model_len, &cs, bits, out.result_is_int, &v) != CE_OK) { err = "failed to evaluate CEL expression"; return 2; }
  return 0;
}

cudaError_t launch_cel(const CelLaunch& L, cudaStream_t st) {
  const uint64_t tot = (uint64_t)L.n * L.n_progs;
  if (!tot) return cudaSuccess;
  cel_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L);
  return cudaGetLastError();
}

}  // namespace aigw
