// AMPL .nl (ASCII "g" format) reader + tape compiler.
//
// Restates, for the engine, what the reference gets from ASL through Ipopt::AmplTNLP
// (src/Interfaces/Ampl/BonAmplTMINLP.cpp:176-178 construction; :473-564 variable typing from the header counts;
// :607-612 constraint linearity; fixture test/mytoy.nl:1-55).  The expression segments are prefix trees
// (`oN` operator, `nX` constant, `vI` variable); each function is lowered to
//     c + a.x + 1/2 x'Qx + sum of element terms
// by symbolic degree-<=2 polynomial extraction; what is not polynomial of degree <= 2 becomes a postfix element term
// over at most BNLP_MAX_TERM_VARS variables (partially separable structure) — up to BNLP_MAX_WIDE_TERM_VARS in a model
// that the wide build of the term machine evaluates (include/bnlp.h).
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include <sstream>

#include "../../include/bnlp.h"
#include "bnlp_internal.h"

namespace bnlp {
namespace {

struct Node {
  enum Kind { NUM, VAR, OP } kind = NUM;
  double num = 0;
  int var = 0, op = 0;
  std::vector<std::unique_ptr<Node>> kids;
};
using NodeP = std::unique_ptr<Node>;

struct Poly {  // degree <= 2 polynomial
  double c = 0;
  std::map<int, double> lin;
  std::map<std::pair<int, int>, double> quad;  // key (i>=j): coefficient of x_i x_j
  int degree() const { return !quad.empty() ? 2 : (!lin.empty() ? 1 : 0); }
  void add(const Poly &o, double s) {
    c += s * o.c;
    for (auto &kv : o.lin) lin[kv.first] += s * kv.second;
    for (auto &kv : o.quad) quad[kv.first] += s * kv.second;
  }
};

bool poly_mul(const Poly &a, const Poly &b, Poly &out) {
  if (a.degree() + b.degree() > 2) return false;
  out = Poly();
  out.c = a.c * b.c;
  for (auto &kv : a.lin) out.lin[kv.first] += kv.second * b.c;
  for (auto &kv : b.lin) out.lin[kv.first] += kv.second * a.c;
  for (auto &kv : a.quad) out.quad[kv.first] += kv.second * b.c;
  for (auto &kv : b.quad) out.quad[kv.first] += kv.second * a.c;
  for (auto &x : a.lin)
    for (auto &y : b.lin) {
      const int i = std::max(x.first, y.first), j = std::min(x.first, y.first);
      out.quad[{i, j}] += x.second * y.second;
    }
  return true;
}

// AMPL opcodes (public ASL opcode table; not part of the reference tree); only the subset below is supported.
// tanh, tan, sinh, log10, cosh are lowered at parse time to the operations the tape has (exp, sin, cos, log, /).
enum { O_PLUS = 0, O_MINUS = 1, O_MULT = 2, O_DIV = 3, O_POW = 5, O_ABS = 15, O_NEG = 16, O_TANH = 37, O_TAN = 38,
       O_SQRT = 39, O_SINH = 40, O_SIN = 41, O_LOG10 = 42, O_LOG = 43, O_EXP = 44, O_COSH = 45, O_COS = 46,
       O_ATAN = 49, O_ASIN = 51, O_ACOS = 53, O_SUMLIST = 54, O_1POW = 75, O_2POW = 76, O_CPOW = 77 };

int arity(int op) {
  switch (op) {
    case O_PLUS: case O_MINUS: case O_MULT: case O_DIV: case O_POW: case O_1POW: case O_CPOW: return 2;
    case O_ABS: case O_NEG: case O_SQRT: case O_SIN: case O_LOG: case O_EXP: case O_COS: case O_2POW:
    case O_TANH: case O_TAN: case O_SINH: case O_LOG10: case O_COSH: case O_ATAN: case O_ASIN: case O_ACOS: return 1;
    case O_SUMLIST: return -1;
    default: return -2;
  }
}

NodeP clone(const Node &e) {
  NodeP c(new Node());
  c->kind = e.kind; c->num = e.num; c->var = e.var; c->op = e.op;
  for (auto &k : e.kids) c->kids.push_back(clone(*k));
  return c;
}
NodeP num(double v) { NodeP c(new Node()); c->kind = Node::NUM; c->num = v; return c; }
NodeP op1(int op, NodeP a) { NodeP c(new Node()); c->kind = Node::OP; c->op = op; c->kids.push_back(std::move(a)); return c; }
NodeP op2(int op, NodeP a, NodeP b) {
  NodeP c(new Node()); c->kind = Node::OP; c->op = op; c->kids.push_back(std::move(a)); c->kids.push_back(std::move(b)); return c;
}
// functions the tape has no opcode for, in terms of the ones it has
NodeP lower_unary(int op, NodeP u) {
  switch (op) {
    case O_LOG10: return op2(O_MULT, num(0.43429448190325182765), op1(O_LOG, std::move(u)));
    case O_TAN: { NodeP v = clone(*u); return op2(O_DIV, op1(O_SIN, std::move(u)), op1(O_COS, std::move(v))); }
    case O_SINH: { NodeP v = clone(*u);
                   return op2(O_MULT, num(0.5), op2(O_MINUS, op1(O_EXP, std::move(u)), op1(O_EXP, op1(O_NEG, std::move(v))))); }
    case O_COSH: { NodeP v = clone(*u);
                   return op2(O_MULT, num(0.5), op2(O_PLUS, op1(O_EXP, std::move(u)), op1(O_EXP, op1(O_NEG, std::move(v))))); }
    case O_TANH:   // 1 - 2 / (exp(2u) + 1)
      return op2(O_MINUS, num(1.0), op2(O_DIV, num(2.0), op2(O_PLUS, op1(O_EXP, op2(O_MULT, num(2.0), std::move(u))), num(1.0))));
    default: return nullptr;
  }
}

struct Lines {
  std::vector<std::string> v;
  size_t pos = 0;
  bool next(std::string &out) {
    while (pos < v.size()) {
      std::string s = v[pos++];
      const size_t h = s.find('#');
      if (h != std::string::npos) s.erase(h);
      size_t a = s.find_first_not_of(" \t\r");
      if (a == std::string::npos) continue;
      size_t b = s.find_last_not_of(" \t\r");
      out = s.substr(a, b - a + 1);
      return true;
    }
    return false;
  }
};

// defs: the defined variables (V segments) read so far; a reference v(nvars + i) is replaced by a copy of definition i
bool parse_expr(Lines &L, NodeP &out, std::string &err, int nvars, const std::vector<NodeP> *defs = nullptr) {
  std::string s;
  if (!L.next(s)) { err = "unexpected end of file in expression"; return false; }
  out.reset(new Node());
  if (s[0] == 'n') { out->kind = Node::NUM; out->num = strtod(s.c_str() + 1, nullptr); return true; }
  if (s[0] == 'v') {
    out->kind = Node::VAR; out->var = atoi(s.c_str() + 1);
    if (out->var >= nvars && defs && out->var - nvars < (int)defs->size()) { out = clone(*(*defs)[out->var - nvars]); return true; }
    if (out->var < 0 || out->var >= nvars) { err = "variable index out of range: " + s; return false; }
    return true;
  }
  if (s[0] == 'o') {
    out->kind = Node::OP; out->op = atoi(s.c_str() + 1);
    int k = arity(out->op);
    if (k == -2) { err = "unsupported .nl opcode o" + std::to_string(out->op); return false; }
    if (k == -1) {
      std::string c;
      if (!L.next(c)) { err = "missing sumlist count"; return false; }
      k = atoi(c.c_str());
    }
    for (int i = 0; i < k; ++i) {
      NodeP kid;
      if (!parse_expr(L, kid, err, nvars, defs)) return false;
      out->kids.push_back(std::move(kid));
    }
    if (NodeP low = (k == 1) ? lower_unary(out->op, clone(*out->kids[0])) : nullptr) out = std::move(low);
    return true;
  }
  err = "bad expression line: " + s;
  return false;
}

bool is_const(const Node &e, double &v) {
  if (e.kind == Node::NUM) { v = e.num; return true; }
  return false;
}

bool to_poly(const Node &e, Poly &out) {
  out = Poly();
  if (e.kind == Node::NUM) { out.c = e.num; return true; }
  if (e.kind == Node::VAR) { out.lin[e.var] = 1.0; return true; }
  Poly a, b;
  switch (e.op) {
    case O_PLUS: case O_MINUS:
      if (!to_poly(*e.kids[0], a) || !to_poly(*e.kids[1], b)) return false;
      out = a; out.add(b, e.op == O_PLUS ? 1.0 : -1.0); return true;
    case O_SUMLIST:
      for (auto &k : e.kids) { if (!to_poly(*k, a)) return false; out.add(a, 1.0); }
      return true;
    case O_NEG:
      if (!to_poly(*e.kids[0], a)) return false;
      out.add(a, -1.0); return true;
    case O_MULT:
      if (!to_poly(*e.kids[0], a) || !to_poly(*e.kids[1], b)) return false;
      return poly_mul(a, b, out);
    case O_DIV:
      if (!to_poly(*e.kids[0], a) || !to_poly(*e.kids[1], b) || b.degree() != 0 || b.c == 0.0) return false;
      out.add(a, 1.0 / b.c); return true;
    case O_POW: case O_1POW: {
      double p;
      if (!is_const(*e.kids[1], p) || !to_poly(*e.kids[0], a)) return false;
      if (p == 0.0) { out.c = 1.0; return true; }
      if (p == 1.0) { out = a; return true; }
      if (p == 2.0) return poly_mul(a, a, out);
      if (a.degree() == 0) { out.c = std::pow(a.c, p); return true; }
      return false;
    }
    case O_2POW:
      if (!to_poly(*e.kids[0], a)) return false;
      return poly_mul(a, a, out);
    default: return false;
  }
}

void collect_vars(const Node &e, std::set<int> &vs) {
  if (e.kind == Node::VAR) vs.insert(e.var);
  for (auto &k : e.kids) collect_vars(*k, vs);
}

bool emit(const Node &e, const std::map<int, int> &slot, Term &T, std::string &err) {
  auto push = [&](int op, int ia, double da) { T.op.push_back(op); T.iarg.push_back(ia); T.darg.push_back(da); };
  if (e.kind == Node::NUM) { push(BNLP_OP_CONST, 0, e.num); return true; }
  if (e.kind == Node::VAR) { push(BNLP_OP_VAR, slot.at(e.var), 0.0); return true; }
  auto kids = [&](size_t upto) {
    for (size_t i = 0; i < upto; ++i)
      if (!emit(*e.kids[i], slot, T, err)) return false;
    return true;
  };
  double p;
  switch (e.op) {
    case O_PLUS: if (!kids(2)) return false; push(BNLP_OP_ADD, 0, 0); return true;
    case O_MINUS: if (!kids(2)) return false; push(BNLP_OP_SUB, 0, 0); return true;
    case O_MULT: if (!kids(2)) return false; push(BNLP_OP_MUL, 0, 0); return true;
    case O_DIV: if (!kids(2)) return false; push(BNLP_OP_DIV, 0, 0); return true;
    case O_POW: case O_1POW:
      if (is_const(*e.kids[1], p)) { if (!kids(1)) return false; push(BNLP_OP_POWC, 0, p); return true; }
      if (!kids(2)) return false; push(BNLP_OP_POW, 0, 0); return true;
    case O_CPOW: if (!kids(2)) return false; push(BNLP_OP_POW, 0, 0); return true;
    case O_2POW: if (!kids(1)) return false; push(BNLP_OP_SQR, 0, 0); return true;
    case O_NEG: if (!kids(1)) return false; push(BNLP_OP_NEG, 0, 0); return true;
    case O_ABS: if (!kids(1)) return false; push(BNLP_OP_ABS, 0, 0); return true;
    case O_SQRT: if (!kids(1)) return false; push(BNLP_OP_SQRT, 0, 0); return true;
    case O_SIN: if (!kids(1)) return false; push(BNLP_OP_SIN, 0, 0); return true;
    case O_COS: if (!kids(1)) return false; push(BNLP_OP_COS, 0, 0); return true;
    case O_ATAN: if (!kids(1)) return false; push(BNLP_OP_ATAN, 0, 0); return true;
    case O_ASIN: if (!kids(1)) return false; push(BNLP_OP_ASIN, 0, 0); return true;
    case O_ACOS: if (!kids(1)) return false; push(BNLP_OP_ACOS, 0, 0); return true;
    case O_LOG: if (!kids(1)) return false; push(BNLP_OP_LOG, 0, 0); return true;
    case O_EXP: if (!kids(1)) return false; push(BNLP_OP_EXP, 0, 0); return true;
    case O_SUMLIST:
      if (e.kids.empty()) { push(BNLP_OP_CONST, 0, 0.0); return true; }
      if (!emit(*e.kids[0], slot, T, err)) return false;
      for (size_t i = 1; i < e.kids.size(); ++i) {
        if (!emit(*e.kids[i], slot, T, err)) return false;
        push(BNLP_OP_ADD, 0, 0);
      }
      return true;
    default: err = "unsupported opcode in element term"; return false;
  }
}

// lower one function expression into (poly, element terms); sums are split so that each addend is lowered separately
bool lower(const Node &e, double sign, int fn, Poly &P, std::vector<Term> &terms, std::string &err) {
  if (e.kind == Node::OP) {
    if (e.op == O_PLUS || e.op == O_SUMLIST) {
      for (auto &k : e.kids) if (!lower(*k, sign, fn, P, terms, err)) return false;
      return true;
    }
    if (e.op == O_MINUS) return lower(*e.kids[0], sign, fn, P, terms, err) && lower(*e.kids[1], -sign, fn, P, terms, err);
    if (e.op == O_NEG) return lower(*e.kids[0], -sign, fn, P, terms, err);
  }
  Poly q;
  if (to_poly(e, q)) { P.add(q, sign); return true; }
  // c * g(x): pull a constant factor out so that the sign/scale does not hide a polynomial
  std::set<int> vs;
  collect_vars(e, vs);
  if ((int)vs.size() > BNLP_MAX_WIDE_TERM_VARS) {
    err = "a non-quadratic, non-separable expression couples more than " + std::to_string(BNLP_MAX_WIDE_TERM_VARS) + " variables";
    return false;
  }
  Term T;
  T.fn = fn;
  std::map<int, int> slot;
  for (int v : vs) { slot[v] = (int)T.vars.size(); T.vars.push_back(v); }
  if (!emit(e, slot, T, err)) return false;
  if (sign != 1.0) {
    if (sign == -1.0) { T.op.push_back(BNLP_OP_NEG); T.iarg.push_back(0); T.darg.push_back(0); }
    else { T.op.push_back(BNLP_OP_CONST); T.iarg.push_back(0); T.darg.push_back(sign);
           T.op.push_back(BNLP_OP_MUL); T.iarg.push_back(0); T.darg.push_back(0); }
  }
  terms.push_back(std::move(T));
  return true;
}

}  // namespace

bool parse_nl(const char *text, size_t len, HostModel &M, std::string &err) {
  Lines L;
  {
    std::string all(text, len);
    std::stringstream ss(all);
    std::string s;
    while (std::getline(ss, s)) L.v.push_back(s);
  }
  std::string s;
  if (!L.next(s) || s.empty() || s[0] != 'g') { err = "only ASCII ('g') .nl files are supported"; return false; }
  auto ints = [&](std::vector<long> &out) {
    out.clear();
    if (!L.next(s)) return false;
    std::stringstream ss(s);
    double v;
    while (ss >> v) out.push_back((long)v);
    return true;
  };
  std::vector<long> h;
  if (!ints(h) || h.size() < 3) { err = "bad header line 2"; return false; }
  const int n = (int)h[0], m = (int)h[1], nobj = (int)h[2];
  if (n <= 0 || m < 0 || nobj > 1) { err = "unsupported problem sizes (need n > 0 and at most one objective)"; return false; }
  if (!ints(h) || h.size() < 2) { err = "bad header line 3"; return false; }
  if (!ints(h)) { err = "bad header line 4"; return false; }
  if (h.size() >= 1 && h[0] + (h.size() > 1 ? h[1] : 0) > 0) { err = "network constraints are not supported"; return false; }
  if (!ints(h) || h.size() < 3) { err = "bad header line 5"; return false; }
  const int nlvc = (int)h[0], nlvo = (int)h[1], nlvb = (int)h[2];
  if (!ints(h)) { err = "bad header line 6"; return false; }
  if (h.size() > 1 && h[1] > 0) { err = "imported functions are not supported"; return false; }
  if (!ints(h) || h.size() < 5) { err = "bad header line 7"; return false; }
  const int nbv = (int)h[0], niv = (int)h[1], nlvbi = (int)h[2], nlvci = (int)h[3], nlvoi = (int)h[4];
  if (!ints(h)) { err = "bad header line 8"; return false; }
  if (!ints(h)) { err = "bad header line 9"; return false; }
  if (!ints(h)) { err = "bad header line 10"; return false; }
  // line 10: numbers of common expressions (defined variables); they are substituted into the expressions that use them

  M = HostModel();
  M.n = n; M.m = m;
  M.xl.assign(n, -kInf); M.xu.assign(n, kInf); M.gl.assign(m, -kInf); M.gu.assign(m, kInf);
  M.x0.assign(n, 0.0); M.c.assign(m + 1, 0.0);
  // variable types from the header counts (BonAmplTMINLP.cpp:473-564)
  M.vtype.assign(n, 0);
  {
    int start = 0, end = nlvb - nlvbi;
    start = end; end = start + nlvbi;
    for (int i = start; i < end && i < n; ++i) M.vtype[i] = 2;
    start = end; end = std::max(start, end + nlvc - nlvci - nlvb);
    start = end; end = start + nlvci;
    for (int i = start; i < end && i < n; ++i) M.vtype[i] = 2;
    start = end; end = std::max(start, end + nlvo - std::max(nlvb, nlvc) - nlvoi);
    start = end; end = start + nlvoi;
    for (int i = start; i < end && i < n; ++i) M.vtype[i] = 2;
    start = n - nbv - niv; end = start + nbv;
    for (int i = std::max(start, 0); i < end; ++i) M.vtype[i] = 1;
    start = end; end = start + niv;
    for (int i = std::max(start, 0); i < end && i < n; ++i) M.vtype[i] = 2;
  }
  std::vector<Poly> P(m + 1);
  std::vector<std::map<int, double>> lin(m + 1);
  double obj_sign = 1.0;
  std::vector<Term> obj_terms;
  std::vector<NodeP> defs;
  while (L.next(s)) {
    const char k = s[0];
    if (k == 'V') {   // defined variable n + i = sum of its linear terms + its expression
      std::stringstream ss(s.substr(1));
      int idx = 0, nlin = 0, usage = 0;
      ss >> idx >> nlin >> usage;
      if (idx != n + (int)defs.size()) { err = "defined variables must come in order: " + s; return false; }
      NodeP sum(new Node());
      sum->kind = Node::OP; sum->op = O_SUMLIST;
      for (int i = 0; i < nlin; ++i) {
        if (!L.next(s)) { err = "short V segment"; return false; }
        std::stringstream s2(s);
        int v; double a;
        s2 >> v >> a;
        NodeP ref(new Node());
        if (v >= n && v - n < (int)defs.size()) ref = clone(*defs[v - n]);
        else if (v >= 0 && v < n) { ref->kind = Node::VAR; ref->var = v; }
        else { err = "V segment variable out of range"; return false; }
        sum->kids.push_back(op2(O_MULT, num(a), std::move(ref)));
      }
      NodeP e;
      if (!parse_expr(L, e, err, n, &defs)) return false;
      sum->kids.push_back(std::move(e));
      defs.push_back(std::move(sum));
    } else if (k == 'C' || k == 'O') {
      std::stringstream ss(s.substr(1));
      int idx = 0, sense = 0;
      ss >> idx;
      if (k == 'O') ss >> sense;
      if ((k == 'C' && (idx < 0 || idx >= m)) || (k == 'O' && idx != 0)) { err = "segment index out of range: " + s; return false; }
      NodeP e;
      if (!parse_expr(L, e, err, n, &defs)) return false;
      const int fn = (k == 'C') ? idx + 1 : 0;
      if (k == 'O' && sense == 1) obj_sign = -1.0;   // maximisation is solved as min -f (as Ipopt::AmplTNLP does)
      if (!lower(*e, (k == 'O') ? obj_sign : 1.0, fn, P[fn], M.terms, err)) return false;
    } else if (k == 'r' || k == 'b') {
      const int cnt = (k == 'r') ? m : n;
      for (int i = 0; i < cnt; ++i) {
        if (!L.next(s)) { err = "short bounds segment"; return false; }
        std::stringstream ss(s);
        int code; double a = 0, b = 0;
        ss >> code;
        double lo = -kInf, up = kInf;
        switch (code) {
          case 0: ss >> a >> b; lo = a; up = b; break;
          case 1: ss >> a; up = a; break;
          case 2: ss >> a; lo = a; break;
          case 3: break;
          case 4: ss >> a; lo = up = a; break;
          default: err = "complementarity constraints are not supported"; return false;
        }
        if (k == 'r') { M.gl[i] = lo; M.gu[i] = up; } else { M.xl[i] = lo; M.xu[i] = up; }
      }
    } else if (k == 'k') {
      const int cnt = atoi(s.c_str() + 1);
      for (int i = 0; i < cnt; ++i) if (!L.next(s)) { err = "short k segment"; return false; }
    } else if (k == 'J' || k == 'G') {
      std::stringstream ss(s.substr(1));
      int idx = 0, cnt = 0;
      ss >> idx >> cnt;
      const int fn = (k == 'J') ? idx + 1 : 0;
      if (fn < 0 || fn > m) { err = "J/G index out of range"; return false; }
      for (int i = 0; i < cnt; ++i) {
        if (!L.next(s)) { err = "short J/G segment"; return false; }
        std::stringstream s2(s);
        int v; double a;
        s2 >> v >> a;
        if (v < 0 || v >= n) { err = "J/G variable out of range"; return false; }
        lin[fn][v] += a;   // structural entry even when the coefficient is 0 (test/mytoy.nl:41-42)
      }
    } else if (k == 'x') {
      const int cnt = atoi(s.c_str() + 1);
      for (int i = 0; i < cnt; ++i) {
        if (!L.next(s)) { err = "short x segment"; return false; }
        std::stringstream s2(s);
        int v; double a;
        s2 >> v >> a;
        if (v >= 0 && v < n) M.x0[v] = a;
      }
    } else if (k == 'd') {
      const int cnt = atoi(s.c_str() + 1);
      for (int i = 0; i < cnt; ++i) L.next(s);
    } else if (k == 'S') {
      // suffix "S<kind> <count> <name>": kind & 3 = 0 variables, 1 constraints, 2 objectives, 3 problem; kind & 4 = real values.
      // The variable suffixes AmplTMINLP::read_priorities (BonAmplTMINLP.cpp:215-252) and read_sos (:254-296: sosno, ref) use
      // are kept; the rest are skipped.
      std::stringstream ss(s.substr(1));
      int kind = 0, cnt = 0;
      std::string name;
      ss >> kind >> cnt >> name;
      std::vector<int> *ti = nullptr;
      std::vector<double> *td = nullptr;
      if ((kind & 3) == 0) {
        if (name == "priority") ti = &M.br_priority;
        else if (name == "direction") ti = &M.br_direction;
        else if (name == "upPseudocost") td = &M.br_up_ps;
        else if (name == "downPseudocost") td = &M.br_down_ps;
        else if (name == "sosno") ti = &M.sos_no;
        else if (name == "ref") td = &M.sos_ref;
      }
      if (ti) ti->assign(n, 0);
      if (td) td->assign(n, 0.0);
      for (int i = 0; i < cnt; ++i) {
        L.next(s);
        if (!ti && !td) continue;
        std::stringstream s2(s);
        int v; double a;
        s2 >> v >> a;
        if (v < 0 || v >= n) { err = "suffix index out of range"; return false; }
        if (ti) (*ti)[v] = (int)a; else (*td)[v] = a;
      }
    } else {
      err = std::string("unsupported .nl segment '") + k + "'";
      return false;
    }
  }
  // assemble: G/J coefficients of a maximised objective are negated as well
  M.lin_start.assign(m + 2, 0);
  for (int fn = 0; fn <= m; ++fn) {
    std::map<int, double> row;
    for (auto &kv : lin[fn]) row[kv.first] += (fn == 0 ? obj_sign : 1.0) * kv.second;
    for (auto &kv : P[fn].lin) row[kv.first] += kv.second;
    M.c[fn] = P[fn].c;
    M.lin_start[fn] = (int)M.lin_col.size();
    for (auto &kv : row) { M.lin_col.push_back(kv.first); M.lin_val.push_back(kv.second); }
    for (auto &kv : P[fn].quad) {
      if (kv.second == 0.0) continue;
      const int i = kv.first.first, j = kv.first.second;
      M.qs_fn.push_back(fn); M.qs_i.push_back(i); M.qs_j.push_back(j);
      // coefficient of x_i x_j is q; in the 1/2 x'Qx convention Q_ii = 2q, Q_ij = Q_ji = q
      M.qs_v.push_back(i == j ? 2.0 * kv.second : kv.second);
    }
  }
  M.lin_start[m + 1] = (int)M.lin_col.size();
  if (term_width_class(M, err) < 0) return false;
  M.finalize();
  return true;
}

}  // namespace bnlp
