// Host-side model compiler: structured description -> HostModel with derived sparsity structures.
// Restates the *data contract* of Bonmin::TMINLP (src/Interfaces/BonTMINLP.hpp:226-333): sizes, bounds, types,
// Jacobian triplet structure (row-major here), lower-triangular Lagrangian-Hessian structure.
#include "bnlp_internal.h"

#include <algorithm>
#include <map>
#include <set>
#include <tuple>

#include "../../include/bnlp.h"

namespace bnlp {

void HostModel::finalize() {
  const size_t P = (size_t)n * (n + 1) / 2;
  // which functions have which kind of nonlinearity
  std::vector<char> dense_q(m + 1, 0);
  for (int fn : qd_fn) dense_q[fn] = 1;
  std::vector<std::set<int>> cols(m + 1);
  std::set<std::pair<int, int>> hes;
  std::vector<char> nonlin(m + 1, 0);
  for (int k = 0; k <= m; ++k)
    for (int e = lin_start[k]; e < lin_start[k + 1]; ++e) cols[k].insert(lin_col[e]);
  for (size_t q = 0; q < qd_fn.size(); ++q) {
    const int fn = qd_fn[q];
    nonlin[fn] = 1;
    for (int j = 0; j < n; ++j) cols[fn].insert(j);
  }
  bool any_dense = !qd_fn.empty();
  if (any_dense) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) hes.insert({i, j});
  }
  for (size_t e = 0; e < qs_fn.size(); ++e) {
    nonlin[qs_fn[e]] = 1;
    cols[qs_fn[e]].insert(qs_i[e]);
    cols[qs_fn[e]].insert(qs_j[e]);
    hes.insert({std::max(qs_i[e], qs_j[e]), std::min(qs_i[e], qs_j[e])});
  }
  for (const Term &t : terms) {
    nonlin[t.fn] = 1;
    for (int a : t.vars) {
      cols[t.fn].insert(a);
      for (int b : t.vars) hes.insert({std::max(a, b), std::min(a, b)});
    }
  }
  (void)P;
  jac_row.clear(); jac_col.clear(); jac_start.assign(m + 1, 0);
  for (int i = 0; i < m; ++i) {
    jac_start[i] = (int)jac_row.size();
    for (int j : cols[i + 1]) { jac_row.push_back(i); jac_col.push_back(j); }
  }
  jac_start[m] = (int)jac_row.size();
  hes_row.clear(); hes_col.clear();
  for (auto &p : hes) { hes_row.push_back(p.first); hes_col.push_back(p.second); }
  grad_f_cols.assign(cols[0].begin(), cols[0].end());
  con_nonlinear.assign(m, 0);
  for (int i = 0; i < m; ++i) con_nonlinear[i] = nonlin[i + 1];
  obj_nonlinear = nonlin[0];
}

bool model_from_desc(const bnlp_tape_desc *d, HostModel &M, std::string &err) {
  if (!d || d->n <= 0 || d->m < 0) { err = "bad sizes"; return false; }
  const int n = d->n, m = d->m;
  M.n = n; M.m = m;
  M.xl.assign(d->xl, d->xl + n); M.xu.assign(d->xu, d->xu + n);
  M.gl.assign(d->gl, d->gl + m); M.gu.assign(d->gu, d->gu + m);
  if (d->x0) M.x0.assign(d->x0, d->x0 + n); else M.x0.assign(n, 0.0);
  if (d->var_type) M.vtype.assign(d->var_type, d->var_type + n); else M.vtype.assign(n, 0);
  M.c.assign(d->c, d->c + m + 1);
  M.lin_start.assign(d->lin_start, d->lin_start + m + 2);
  const int nl = M.lin_start[m + 1];
  M.lin_col.assign(d->lin_col, d->lin_col + nl);
  M.lin_val.assign(d->lin_val, d->lin_val + nl);
  for (int e = 0; e < nl; ++e)
    if (M.lin_col[e] < 0 || M.lin_col[e] >= n) { err = "linear column out of range"; return false; }
  const size_t P = (size_t)n * (n + 1) / 2;
  M.qd_fn.assign(d->qd_fn, d->qd_fn + d->n_qdense);
  M.qd.assign(d->qd, d->qd + (size_t)d->n_qdense * P);
  for (int fn : M.qd_fn)
    if (fn < 0 || fn > m) { err = "dense quadratic function index out of range"; return false; }
  M.qs_fn.assign(d->qs_fn, d->qs_fn + d->n_qsparse);
  M.qs_i.assign(d->qs_i, d->qs_i + d->n_qsparse);
  M.qs_j.assign(d->qs_j, d->qs_j + d->n_qsparse);
  M.qs_v.assign(d->qs_v, d->qs_v + d->n_qsparse);
  for (int e = 0; e < d->n_qsparse; ++e) {
    if (M.qs_fn[e] < 0 || M.qs_fn[e] > m || M.qs_i[e] < 0 || M.qs_i[e] >= n || M.qs_j[e] < 0 || M.qs_j[e] >= n) {
      err = "sparse quadratic index out of range"; return false;
    }
    if (M.qs_i[e] < M.qs_j[e]) std::swap(M.qs_i[e], M.qs_j[e]);
  }
  M.terms.clear();
  for (int t = 0; t < d->n_terms; ++t) {
    Term T;
    T.fn = d->term_fn[t];
    if (T.fn < 0 || T.fn > m) { err = "term function index out of range"; return false; }
    T.vars.assign(d->term_var + d->term_var_start[t], d->term_var + d->term_var_start[t + 1]);
    if ((int)T.vars.size() > BNLP_MAX_WIDE_TERM_VARS) { err = "term has more than BNLP_MAX_WIDE_TERM_VARS = " + std::to_string(BNLP_MAX_WIDE_TERM_VARS) + " variables"; return false; }
    for (int v : T.vars) if (v < 0 || v >= n) { err = "term variable out of range"; return false; }
    const int a = d->term_op_start[t], b = d->term_op_start[t + 1];
    T.op.assign(d->term_op + a, d->term_op + b);
    T.iarg.assign(d->term_iarg + a, d->term_iarg + b);
    T.darg.assign(d->term_darg + a, d->term_darg + b);
    // validate stack discipline
    int depth = 0, maxd = 0;
    for (size_t k = 0; k < T.op.size(); ++k) {
      switch (T.op[k]) {
        case BNLP_OP_CONST: depth++; break;
        case BNLP_OP_VAR:
          if (T.iarg[k] < 0 || T.iarg[k] >= (int)T.vars.size()) { err = "term VAR slot out of range"; return false; }
          depth++; break;
        case BNLP_OP_ADD: case BNLP_OP_SUB: case BNLP_OP_MUL: case BNLP_OP_DIV: case BNLP_OP_POW:
          if (depth < 2) { err = "term stack underflow"; return false; }
          depth--; break;
        case BNLP_OP_NEG: case BNLP_OP_POWC: case BNLP_OP_EXP: case BNLP_OP_LOG: case BNLP_OP_SQRT:
        case BNLP_OP_SIN: case BNLP_OP_COS: case BNLP_OP_SQR: case BNLP_OP_ABS: case BNLP_OP_ATAN: case BNLP_OP_ASIN: case BNLP_OP_ACOS:
          if (depth < 1) { err = "term stack underflow"; return false; }
          break;
        default: err = "unknown term opcode"; return false;
      }
      maxd = std::max(maxd, depth);
    }
    if (depth != 1) { err = "term program does not leave exactly one value"; return false; }
    if (maxd > BNLP_MAX_TERM_STACK) { err = "term stack deeper than BNLP_MAX_TERM_STACK"; return false; }
    M.terms.push_back(std::move(T));
  }
  if (term_width_class(M, err) < 0) return false;
  M.finalize();
  return true;
}

int term_stack_depth(const Term &T) {
  int depth = 0, maxd = 0;
  for (int op : T.op) {
    if (op == BNLP_OP_CONST || op == BNLP_OP_VAR) depth++;
    else if (op == BNLP_OP_ADD || op == BNLP_OP_SUB || op == BNLP_OP_MUL || op == BNLP_OP_DIV || op == BNLP_OP_POW) depth--;
    maxd = std::max(maxd, depth);
  }
  return maxd;
}

int term_width_class(const HostModel &M, std::string &err) {
  bool wide = false;
  for (const Term &T : M.terms) wide |= (int)T.vars.size() > BNLP_MAX_TERM_VARS;
  for (const Term &T : M.terms) {
    const int nv = (int)T.vars.size(), d = term_stack_depth(T);
    if (nv > (wide ? BNLP_MAX_WIDE_TERM_VARS : BNLP_MAX_TERM_VARS)) {
      err = "a nonlinear element term couples " + std::to_string(nv) + " variables (limit " + std::to_string(BNLP_MAX_WIDE_TERM_VARS) + ")";
      return -1;
    }
    if (d > (wide ? BNLP_MAX_WIDE_TERM_STACK : BNLP_MAX_TERM_STACK)) {
      err = "a nonlinear element term needs a postfix stack of " + std::to_string(d) + " entries (limit " +
            std::to_string(wide ? BNLP_MAX_WIDE_TERM_STACK : BNLP_MAX_TERM_STACK) + (wide ? " in a model with wide terms)" : ")");
      return -1;
    }
  }
  return wide ? 1 : 0;
}

// ---- TNLP2FPNLP (src/Interfaces/BonTNLP2FPNLP.cpp:75-330) as a tape transformation.
// The rounding target enters as k extra variables p (fixed per child through their bounds), so a whole batch of targets
// is ONE problem whose children differ only in bounds:
//   objective (use_fp_objective):  (1 - lambda) sigma f(x) + s lambda dist(x, p)             (:178-190)
//     norm 2: dist = sum_i (x_i - p_i)^2                                                       (:80-84)
//     norm 1: dist = sum_i [p_i <= 0.1 ? x_i : 1 - x_i]  ==  x_i (1 - 2 p_i) + p_i for p_i in {0, 1}   (:85-94)
//   cutoff row (use_cutoff_constraint): the (new) objective <= cutoff                        (:150-155, :204-214)
//   local-branching row: sum_i x_i (1 - 2 p_i) + p_i <= rhs                                  (:215-223)
// CutStrengthener::StrengtheningTNLP (src/Interfaces/BonCutStrengthener.cpp:306-560) as a model variant: the NLP
//   min  obj_grad . x   s.t.  g_l <= g_i(x) <= g_u            (constraint row i of the base model)
//   min  obj_grad . (x, z)   s.t.  f(x) - z <= 0              (constr_index == -1: the objective cut, z = the alpha column)
// whose optimal value is the tightest valid right-hand side of the cut  obj_grad . x >= bound.  The cut coefficients are
// per-cut data, so they enter as nv EXTRA variables p (variable nv + j multiplies variable j in the objective; nv = n, or
// n + 1 with z); a child fixes them through its bounds, exactly like the feasibility-pump targets, so all the cuts of
// one row are ONE batch.  With p fixed the objective is linear in the free variables (the p-x cross block of the Hessian
// only touches fixed variables).  The reference restricts the NLP to the variables that appear in row i; the caller
// gets the same problem by fixing the other variables at the cut point through the child's bounds.
bool model_cs_variant(const HostModel &B, int ci, HostModel &M, std::string &err) {
  const int n = B.n;
  if (ci < -1 || ci >= B.m) { err = "constraint index out of range"; return false; }
  const int nv = n + (ci == -1 ? 1 : 0), n2 = 2 * nv;
  const int src = ci + 1;                         // function index in the base tape (0 = objective)
  M = HostModel();
  M.n = n2; M.m = 1;
  M.xl.assign(n2, -kInf); M.xu.assign(n2, kInf); M.x0.assign(n2, 0.0); M.vtype.assign(n2, 0);
  for (int j = 0; j < n; ++j) { M.xl[j] = B.xl[j]; M.xu[j] = B.xu[j]; M.x0[j] = B.x0[j]; M.vtype[j] = B.vtype[j]; }
  if (ci == -1) { M.gl = {-kInf}; M.gu = {0.0}; }
  else { M.gl = {B.gl[ci]}; M.gu = {B.gu[ci]}; }
  M.c = {0.0, B.c[src]};
  // objective: sum_j p_j x_j  (1/2 x'Qx convention: Q_{p_j, x_j} = 1)
  for (int j = 0; j < nv; ++j) { M.qs_fn.push_back(0); M.qs_i.push_back(nv + j); M.qs_j.push_back(j); M.qs_v.push_back(1.0); }
  // the row
  M.lin_start = {0, 0};
  for (int e = B.lin_start[src]; e < B.lin_start[src + 1]; ++e) { M.lin_col.push_back(B.lin_col[e]); M.lin_val.push_back(B.lin_val[e]); }
  if (ci == -1) { M.lin_col.push_back(n); M.lin_val.push_back(-1.0); }
  M.lin_start.push_back((int)M.lin_col.size());
  for (size_t e = 0; e < B.qs_fn.size(); ++e)
    if (B.qs_fn[e] == src) { M.qs_fn.push_back(1); M.qs_i.push_back(B.qs_i[e]); M.qs_j.push_back(B.qs_j[e]); M.qs_v.push_back(B.qs_v[e]); }
  const size_t P = (size_t)n * (n + 1) / 2;
  for (size_t q = 0; q < B.qd_fn.size(); ++q)
    if (B.qd_fn[q] == src) {        // dense block -> triplets (the variant has 2 nv variables; the block only spans the first n)
      const double *blk = B.qd.data() + q * P;
      size_t e = 0;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j, ++e)
          if (blk[e] != 0.0) { M.qs_fn.push_back(1); M.qs_i.push_back(i); M.qs_j.push_back(j); M.qs_v.push_back(blk[e]); }
    }
  for (const Term &t : B.terms)
    if (t.fn == src) { Term u = t; u.fn = 1; M.terms.push_back(u); }
  M.finalize();
  return true;
}

bool model_fp_variant(const HostModel &B, const bnlp_fp_desc &d, HostModel &M, std::string &err) {
  const int n = B.n, m = B.m, k = d.n_inds;
  if (k < 0 || (k > 0 && !d.inds)) { err = "bad index set"; return false; }
  for (int i = 0; i < k; ++i) if (d.inds[i] < 0 || d.inds[i] >= n) { err = "index out of range"; return false; }
  if (d.norm != 1 && d.norm != 2) { err = "norm must be 1 or 2"; return false; }
  const int n2 = n + k, madd = (d.use_cutoff_constraint ? 1 : 0) + (d.use_local_branching_constraint ? 1 : 0), m2 = m + madd;
  M = HostModel();
  M.n = n2; M.m = m2;
  M.xl = B.xl; M.xu = B.xu; M.x0 = B.x0; M.vtype = B.vtype;
  M.xl.resize(n2, -kInf); M.xu.resize(n2, kInf); M.x0.resize(n2, 0.0); M.vtype.resize(n2, 0);
  M.gl = B.gl; M.gu = B.gu;
  const double fs = d.use_fp_objective ? (1.0 - d.lambda) * d.sigma : 1.0;      // scale of the original objective
  const double ds = d.use_fp_objective ? d.objective_scaling_factor * d.lambda : 0.0;   // scale of the distance
  // per-function pieces, function 0 rebuilt; rows 1..m copied; extra rows appended
  struct Fn { double c = 0; std::vector<std::pair<int, double>> lin; std::vector<std::tuple<int, int, double>> q; };
  std::vector<Fn> F(m2 + 1);
  for (int fn = 0; fn <= m; ++fn) {
    const double sc = fn == 0 ? fs : 1.0;
    F[fn].c = sc * B.c[fn];
    for (int e = B.lin_start[fn]; e < B.lin_start[fn + 1]; ++e) F[fn].lin.push_back({B.lin_col[e], sc * B.lin_val[e]});
  }
  for (size_t e = 0; e < B.qs_fn.size(); ++e)
    F[B.qs_fn[e]].q.push_back({B.qs_i[e], B.qs_j[e], (B.qs_fn[e] == 0 ? fs : 1.0) * B.qs_v[e]});
  auto add_dist = [&](Fn &f, double w) {       // w * dist(x, p)
    for (int i = 0; i < k; ++i) {
      const int xi = d.inds[i], pi = n + i;
      if (d.norm == 2) {                        // w (x^2 - 2 x p + p^2): 1/2 x'Qx convention -> Q_xx = Q_pp = 2w, Q_px = -2w
        f.q.push_back({xi, xi, 2.0 * w}); f.q.push_back({pi, pi, 2.0 * w}); f.q.push_back({pi, xi, -2.0 * w});
      } else {                                  // w (x - 2 x p + p)
        f.lin.push_back({xi, w}); f.lin.push_back({pi, w}); f.q.push_back({pi, xi, -2.0 * w});
      }
    }
  };
  if (d.use_fp_objective) add_dist(F[0], ds);
  int row = m;
  std::vector<int> copy_obj_rows;              // rows that repeat the objective (dense blocks / terms are duplicated below)
  if (d.use_cutoff_constraint) {
    ++row;
    F[row] = F[0];
    copy_obj_rows.push_back(row);
    M.gl.push_back(-kInf); M.gu.push_back(d.cutoff);
  }
  if (d.use_local_branching_constraint) {
    ++row;
    const int keep = d.norm;
    bnlp_fp_desc d1 = d; d1.norm = 1;
    (void)keep;
    for (int i = 0; i < k; ++i) {
      const int xi = d.inds[i], pi = n + i;
      F[row].lin.push_back({xi, 1.0}); F[row].lin.push_back({pi, 1.0}); F[row].q.push_back({pi, xi, -2.0});
    }
    M.gl.push_back(-kInf); M.gu.push_back(d.rhs_local_branching);
  }
  // dense blocks: re-packed for n2 variables (the packed rows of the first n variables come first, new rows are zero)
  const size_t P = (size_t)n * (n + 1) / 2, P2 = (size_t)n2 * (n2 + 1) / 2;
  auto push_block = [&](int fn, const double *src, double sc) {
    M.qd_fn.push_back(fn);
    const size_t off = M.qd.size();
    M.qd.resize(off + P2, 0.0);
    for (size_t e = 0; e < P; ++e) M.qd[off + e] = sc * src[e];
  };
  for (size_t q = 0; q < B.qd_fn.size(); ++q) {
    const int fn = B.qd_fn[q];
    push_block(fn, B.qd.data() + q * P, fn == 0 ? fs : 1.0);
    if (fn == 0) for (int r : copy_obj_rows) push_block(r, B.qd.data() + q * P, fs);
  }
  // element terms: objective terms are scaled by appending  CONST fs; MUL  to their programs
  auto scaled = [&](const Term &t, int fn, double sc) {
    Term u = t;
    u.fn = fn;
    if (sc != 1.0) { u.op.push_back(BNLP_OP_CONST); u.iarg.push_back(0); u.darg.push_back(sc);
                     u.op.push_back(BNLP_OP_MUL); u.iarg.push_back(0); u.darg.push_back(0.0); }
    return u;
  };
  for (const Term &t : B.terms) {
    M.terms.push_back(scaled(t, t.fn, t.fn == 0 ? fs : 1.0));
    if (t.fn == 0) for (int r : copy_obj_rows) M.terms.push_back(scaled(t, r, fs));
  }
  // assemble CSR + triplets (duplicates merged)
  M.c.assign(m2 + 1, 0.0);
  M.lin_start.assign(m2 + 2, 0);
  for (int fn = 0; fn <= m2; ++fn) {
    M.c[fn] = F[fn].c;
    std::map<int, double> rowm;
    for (auto &kv : F[fn].lin) rowm[kv.first] += kv.second;
    M.lin_start[fn] = (int)M.lin_col.size();
    for (auto &kv : rowm) { M.lin_col.push_back(kv.first); M.lin_val.push_back(kv.second); }
    std::map<std::pair<int, int>, double> qm;
    for (auto &t : F[fn].q) { const int i = std::max(std::get<0>(t), std::get<1>(t)), j = std::min(std::get<0>(t), std::get<1>(t)); qm[{i, j}] += std::get<2>(t); }
    for (auto &kv : qm) { M.qs_fn.push_back(fn); M.qs_i.push_back(kv.first.first); M.qs_j.push_back(kv.first.second); M.qs_v.push_back(kv.second); }
  }
  M.lin_start[m2 + 1] = (int)M.lin_col.size();
  M.finalize();
  return true;
}

}  // namespace bnlp
