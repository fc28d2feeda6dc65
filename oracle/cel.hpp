// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the CEL cost expressions:
//   llmcostcel.NewProgram / EvaluateProgram     internal/llmcostcel/cel.go:17-99
//   evalCost (CEL branch)                       internal/extproc/processor_impl.go:728-751
// The evaluator is github.com/google/cel-go v0.28.1 (not in the tree).  Restated from the CEL language definition for the
// subset the cost expressions use: variables model / backend / route_name (string) and the six *_tokens (uint); int, uint,
// bool, string literals; + - * / % with checked overflow ("integer overflow", "unsigned integer overflow", "divide by zero",
// "modulus by zero"); comparisons and equality between operands of the SAME type (the default environment has no cross-type
// numeric comparison); ! && || with CEL's commutative error absorption; ?: ; int() / uint() conversions with range errors.
// Everything else (double, string functions, `in`, lists, maps, macros, has()) is reported as UNSUPPORTED, never guessed.
// NewProgram's sanity evaluation (all counters 0, strings "dummy") is restated too: an expression failing it never loads.
// Pinned by internal/llmcostcel/cel_test.go:15-90 and examples/token_ratelimit/token_ratelimit.yaml:49-65.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

namespace oracle { namespace cel {

enum Ty { T_INT, T_UINT, T_BOOL, T_STR };
enum Err { E_OK = 0, E_INT_OVERFLOW = 1, E_UINT_OVERFLOW = 2, E_DIV_ZERO = 3, E_MOD_ZERO = 4, E_NEGATIVE = 5 };
struct Val { Ty t; uint64_t u = 0; std::string s; int err = 0; };   // int stored in u as two's complement
struct Vars { std::string model, backend, route; uint64_t tok[6] = {0, 0, 0, 0, 0, 0}; };  // input, cached, cache_creation, output, total, reasoning

struct Node {
  enum K { LIT, VAR, UN, BIN, COND, CONV } k; Ty t; int op = 0; Val lit; int var = 0;
  std::unique_ptr<Node> a, b, c;
};

struct Parser {
  std::string_view s; size_t i = 0; std::string err;
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
  bool eat(const char* w) { ws(); size_t n = strlen(w); if (s.compare(i, n, w) == 0) { i += n; return true; } return false; }
  std::unique_ptr<Node> fail(const std::string& m) { if (err.empty()) err = m; return nullptr; }
  static bool num(Ty t) { return t == T_INT || t == T_UINT; }

  std::unique_ptr<Node> expr() {
    auto c = lor(); if (!c) return nullptr;
    ws();
    if (i < s.size() && s[i] == '?') {
      i++;
      auto a = lor(); if (!a) return nullptr;    // CEL: expr = conditionalOr ['?' conditionalOr ':' expr]
      if (!eat(":")) return fail("expected ':'");
      auto b = expr(); if (!b) return nullptr;
      if (c->t != T_BOOL || a->t != b->t) return fail("no matching overload for _?_:_");
      auto n = std::make_unique<Node>(); n->k = Node::COND; n->t = a->t; n->a = std::move(c); n->b = std::move(a); n->c = std::move(b); return n;
    }
    return c;
  }
  std::unique_ptr<Node> lor() {
    auto l = land(); if (!l) return nullptr;
    for (;;) { ws(); if (s.compare(i, 2, "||") != 0) return l; i += 2; auto r = land(); if (!r) return nullptr; if (l->t != T_BOOL || r->t != T_BOOL) return fail("no matching overload for _||_"); l = bin('O', T_BOOL, std::move(l), std::move(r)); }
  }
  std::unique_ptr<Node> land() {
    auto l = rel(); if (!l) return nullptr;
    for (;;) { ws(); if (s.compare(i, 2, "&&") != 0) return l; i += 2; auto r = rel(); if (!r) return nullptr; if (l->t != T_BOOL || r->t != T_BOOL) return fail("no matching overload for _&&_"); l = bin('A', T_BOOL, std::move(l), std::move(r)); }
  }
  std::unique_ptr<Node> rel() {
    auto l = add(); if (!l) return nullptr;
    for (;;) {
      ws(); int op = 0;
      if (s.compare(i, 2, "==") == 0) { op = 'e'; i += 2; } else if (s.compare(i, 2, "!=") == 0) { op = 'n'; i += 2; }
      else if (s.compare(i, 2, "<=") == 0) { op = 'l'; i += 2; } else if (s.compare(i, 2, ">=") == 0) { op = 'g'; i += 2; }
      else if (i < s.size() && s[i] == '<') { op = '<'; i++; } else if (i < s.size() && s[i] == '>') { op = '>'; i++; }
      else return l;
      auto r = add(); if (!r) return nullptr;
      if (l->t != r->t) return fail("no matching overload for comparison");
      if ((op != 'e' && op != 'n') && !num(l->t)) return fail("unsupported: ordering of non-numeric operands");
      l = bin(op, T_BOOL, std::move(l), std::move(r));
    }
  }
  std::unique_ptr<Node> add() {
    auto l = mul(); if (!l) return nullptr;
    for (;;) { ws(); if (i >= s.size() || (s[i] != '+' && s[i] != '-')) return l; const int op = s[i++]; auto r = mul(); if (!r) return nullptr;
      if (l->t != r->t || !num(l->t)) return fail(op == '+' && l->t == T_STR ? "unsupported: string concatenation" : "no matching overload for arithmetic"); { const Ty lt = l->t; l = bin(op, lt, std::move(l), std::move(r)); } }
  }
  std::unique_ptr<Node> mul() {
    auto l = unary(); if (!l) return nullptr;
    for (;;) { ws(); if (i >= s.size() || (s[i] != '*' && s[i] != '/' && s[i] != '%')) return l; const int op = s[i++]; auto r = unary(); if (!r) return nullptr;
      if (l->t != r->t || !num(l->t)) return fail("no matching overload for arithmetic"); { const Ty lt = l->t; l = bin(op, lt, std::move(l), std::move(r)); } }
  }
  std::unique_ptr<Node> unary() {
    ws();
    if (i < s.size() && s[i] == '!') { i++; auto a = unary(); if (!a) return nullptr; if (a->t != T_BOOL) return fail("no matching overload for !_"); auto n = std::make_unique<Node>(); n->k = Node::UN; n->op = '!'; n->t = T_BOOL; n->a = std::move(a); return n; }
    if (i < s.size() && s[i] == '-') { i++; auto a = unary(); if (!a) return nullptr; if (a->t != T_INT) return fail("no matching overload for -_"); auto n = std::make_unique<Node>(); n->k = Node::UN; n->op = '-'; n->t = T_INT; n->a = std::move(a); return n; }
    return primary();
  }
  std::unique_ptr<Node> primary() {
    ws();
    if (i >= s.size()) return fail("unexpected end of expression");
    const char c = s[i];
    if (c == '(') { i++; auto e = expr(); if (!e) return nullptr; if (!eat(")")) return fail("expected ')'"); return e; }
    if (c == '\'' || c == '"') {
      size_t j = i + 1; while (j < s.size() && s[j] != c) { if (s[j] == '\\' || s[j] == '\n') return fail("unsupported: escapes in string literal"); j++; }
      if (j >= s.size()) return fail("unterminated string");
      if (j + 1 < s.size() && j == i + 1 && s[j + 1] == c) return fail("unsupported: triple-quoted string");
      auto n = std::make_unique<Node>(); n->k = Node::LIT; n->t = T_STR; n->lit.t = T_STR; n->lit.s = std::string(s.substr(i + 1, j - i - 1)); i = j + 1; return n;
    }
    if (c >= '0' && c <= '9') {
      uint64_t v = 0; bool ovf = false; size_t j = i;
      if (s.compare(i, 2, "0x") == 0 || s.compare(i, 2, "0X") == 0) { j = i + 2; size_t d0 = j; while (j < s.size() && isxdigit((unsigned char)s[j])) { const int d = isdigit((unsigned char)s[j]) ? s[j] - '0' : (tolower(s[j]) - 'a' + 10); if (v >> 60) ovf = true; v = v * 16 + d; j++; } if (j == d0) return fail("bad hex literal"); }
      else while (j < s.size() && s[j] >= '0' && s[j] <= '9') { const uint64_t d = s[j] - '0'; if (v > (UINT64_MAX - d) / 10) ovf = true; v = v * 10 + d; j++; }
      if (j < s.size() && (s[j] == '.' || s[j] == 'e' || s[j] == 'E')) return fail("unsupported: double literal");
      auto n = std::make_unique<Node>(); n->k = Node::LIT;
      if (j < s.size() && (s[j] == 'u' || s[j] == 'U')) { if (ovf) return fail("uint literal out of range"); n->t = T_UINT; j++; }
      else { if (ovf || v > (uint64_t)INT64_MAX) return fail("int literal out of range"); n->t = T_INT; }
      n->lit.t = n->t; n->lit.u = v; i = j; return n;
    }
    if (isalpha((unsigned char)c) || c == '_') {
      size_t j = i; while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_')) j++;
      const std::string id(s.substr(i, j - i)); i = j;
      ws();
      if (id == "true" || id == "false") { auto n = std::make_unique<Node>(); n->k = Node::LIT; n->t = T_BOOL; n->lit.t = T_BOOL; n->lit.u = id == "true"; return n; }
      if (i < s.size() && s[i] == '(') {
        if (id != "uint" && id != "int") return fail("unsupported: function " + id);
        i++; auto a = expr(); if (!a) return nullptr; if (!eat(")")) return fail("expected ')'");
        if (!num(a->t)) return fail("unsupported: conversion from non-integer");
        auto n = std::make_unique<Node>(); n->k = Node::CONV; n->t = id == "uint" ? T_UINT : T_INT; n->a = std::move(a); return n;
      }
      if (i < s.size() && (s[i] == '.' || s[i] == '[')) return fail("unsupported: member access / index");
      static const char* names[9] = {"model", "backend", "route_name", "input_tokens", "cached_input_tokens", "cache_creation_input_tokens", "output_tokens", "total_tokens", "reasoning_tokens"};
      for (int k = 0; k < 9; k++) if (id == names[k]) { auto n = std::make_unique<Node>(); n->k = Node::VAR; n->var = k; n->t = k < 3 ? T_STR : T_UINT; return n; }
      if (id == "in") return fail("unsupported: in");
      return fail("undeclared reference to '" + id + "'");
    }
    return fail("unexpected character");
  }
  std::unique_ptr<Node> bin(int op, Ty t, std::unique_ptr<Node> l, std::unique_ptr<Node> r) { auto n = std::make_unique<Node>(); n->k = Node::BIN; n->op = op; n->t = t; n->a = std::move(l); n->b = std::move(r); return n; }
};

inline Val eval(const Node& n, const Vars& v) {
  Val r; r.t = n.t;
  switch (n.k) {
    case Node::LIT: return n.lit;
    case Node::VAR: if (n.var < 3) { r.s = n.var == 0 ? v.model : n.var == 1 ? v.backend : v.route; } else r.u = v.tok[n.var - 3]; return r;
    case Node::CONV: {
      Val a = eval(*n.a, v); if (a.err) { r.err = a.err; return r; }
      if (n.t == T_UINT && a.t == T_INT && (int64_t)a.u < 0) { r.err = E_UINT_OVERFLOW; return r; }
      if (n.t == T_INT && a.t == T_UINT && a.u > (uint64_t)INT64_MAX) { r.err = E_INT_OVERFLOW; return r; }
      r.u = a.u; return r;
    }
    case Node::UN: {
      Val a = eval(*n.a, v); if (a.err) { r.err = a.err; return r; }
      if (n.op == '!') r.u = !a.u; else { if ((int64_t)a.u == INT64_MIN) { r.err = E_INT_OVERFLOW; return r; } r.u = (uint64_t)(-(int64_t)a.u); }
      return r;
    }
    case Node::COND: { Val c = eval(*n.a, v); if (c.err) { r.err = c.err; return r; } return eval(c.u ? *n.b : *n.c, v); }
    case Node::BIN: {
      Val a = eval(*n.a, v), b = eval(*n.b, v);
      if (n.op == 'A') { if ((!a.err && !a.u) || (!b.err && !b.u)) { r.u = 0; return r; } if (a.err || b.err) { r.err = a.err ? a.err : b.err; return r; } r.u = 1; return r; }
      if (n.op == 'O') { if ((!a.err && a.u) || (!b.err && b.u)) { r.u = 1; return r; } if (a.err || b.err) { r.err = a.err ? a.err : b.err; return r; } r.u = 0; return r; }
      if (a.err || b.err) { r.err = a.err ? a.err : b.err; return r; }
      const Ty t = n.a->t;
      switch (n.op) {
        case 'e': r.u = t == T_STR ? a.s == b.s : a.u == b.u; return r;
        case 'n': r.u = t == T_STR ? a.s != b.s : a.u != b.u; return r;
        case '<': r.u = t == T_INT ? (int64_t)a.u < (int64_t)b.u : a.u < b.u; return r;
        case 'l': r.u = t == T_INT ? (int64_t)a.u <= (int64_t)b.u : a.u <= b.u; return r;
        case '>': r.u = t == T_INT ? (int64_t)a.u > (int64_t)b.u : a.u > b.u; return r;
        case 'g': r.u = t == T_INT ? (int64_t)a.u >= (int64_t)b.u : a.u >= b.u; return r;
      }
      if (t == T_UINT) {
        switch (n.op) {
          case '+': if (a.u > UINT64_MAX - b.u) r.err = E_UINT_OVERFLOW; else r.u = a.u + b.u; return r;
          case '-': if (a.u < b.u) r.err = E_UINT_OVERFLOW; else r.u = a.u - b.u; return r;
          case '*': if (b.u && a.u > UINT64_MAX / b.u) r.err = E_UINT_OVERFLOW; else r.u = a.u * b.u; return r;
          case '/': if (!b.u) r.err = E_DIV_ZERO; else r.u = a.u / b.u; return r;
          case '%': if (!b.u) r.err = E_MOD_ZERO; else r.u = a.u % b.u; return r;
        }
      } else {
        const int64_t x = (int64_t)a.u, y = (int64_t)b.u; int64_t z = 0;
        switch (n.op) {
          case '+': if (__builtin_add_overflow(x, y, &z)) r.err = E_INT_OVERFLOW; break;
          case '-': if (__builtin_sub_overflow(x, y, &z)) r.err = E_INT_OVERFLOW; break;
          case '*': if (__builtin_mul_overflow(x, y, &z)) r.err = E_INT_OVERFLOW; break;
          case '/': if (!y) r.err = E_DIV_ZERO; else if (x == INT64_MIN && y == -1) r.err = E_INT_OVERFLOW; else z = x / y; break;
          case '%': if (!y) r.err = E_MOD_ZERO; else if (x == INT64_MIN && y == -1) r.err = E_INT_OVERFLOW; else z = x % y; break;
        }
        r.u = (uint64_t)z; return r;
      }
    }
  }
  return r;
}

struct Program { std::unique_ptr<Node> root; };
// 0 ok; 1 unsupported (outside the restated subset) or a compile / type error; 2 rejected by NewProgram's sanity evaluation
inline int compile(std::string_view expr, Program& p, std::string& err) {
  Parser ps; ps.s = expr;
  auto n = ps.expr();
  if (n) { ps.ws(); if (ps.i != expr.size()) { n = nullptr; if (ps.err.empty()) ps.err = "unexpected trailing input"; } }
  if (!n) { err = ps.err; return 1; }
  if (n->t != T_INT && n->t != T_UINT) { err = "CEL expression result is not an integer"; return 2; }
  p.root = std::move(n);
  Vars d; d.model = d.backend = d.route = "dummy";
  Val v = eval(*p.root, d);
  if (v.err || (v.t == T_INT && (int64_t)v.u < 0)) { err = "failed to evaluate CEL expression"; return 2; }
  return 0;
}
// EvaluateProgram: 0 ok (cost in `out`), else Err
inline int evaluate(const Program& p, const Vars& v, uint64_t& out) {
  Val r = eval(*p.root, v);
  if (r.err) return r.err;
  if (r.t == T_INT && (int64_t)r.u < 0) return E_NEGATIVE;
  out = r.u; return 0;
}

}}  // namespace oracle::cel
