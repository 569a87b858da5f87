// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, load or call this file.
//
// What it is: a scalar CPU restatement of the arithmetic that Bonmin's hot path delegates
// to Ipopt (stable/3.13, un-vendored: /root/reference/.coin-or/Dependencies:11), i.e. what
// TNLPSolver::OptimizeTNLP does for ONE child relaxation
//   call site            src/Interfaces/Ipopt/BonIpoptSolver.cpp:134-176
//   option set           src/Interfaces/Ipopt/BonIpoptSolver.cpp:265-297  (gamma_theta 1e-4, gamma_phi 1e-8,
//                        required_infeasibility_reduction 0.1, expect_infeasible_problem yes,
//                        mu_strategy adaptive, mu_oracle probing, honor_original_bounds yes)
//   numeric defaults     doc/options_list_ipopt_content.tex (tol L2689, bound_relax_factor L3647,
//                        bound_push/frac L2871/L2850, tau_min L2596, sigma_min/max L2590/L2584, ...)
//   status map           src/Interfaces/Ipopt/BonIpoptSolver.cpp:304-343
//   result contract      src/Interfaces/BonTMINLP2TNLP.cpp:474-509  (duals = [z_L | z_U | lambda],
//                        infeasible => obj := max-norm constraint violation)
//   all-fixed short cut  src/Interfaces/BonTNLPSolver.cpp:78-161
// The interior-point algorithm itself is not in the reference tree; it is restated from the
// published method (Waechter & Biegler 2006; Nocedal, Waechter & Waltz 2009) under the option
// values above, in the condensed dense form H = W + Sigma_x + J' Sigma_s J of SURVEY.md §8(d).
// Declared deviations from Ipopt (see DESIGN.md §"IPM restatement"):
//   * rows are normalised to one-sided  h_r(x) + s_r = b_r, s_r >= 0 (ranged rows split in two);
//   * initial constraint multipliers y := v (=1) instead of a least-squares estimate;
//   * NLP scaling is computed at the pushed start point;
//   * the restoration phase is replaced by the exact phase-1 problem  min t  s.t. h_r(x) - t <= b_r
//     (SURVEY.md Appendix D item 10 allows this for convex children), left as soon as the
//     violation has dropped to required_infeasibility_reduction * violation-at-entry;
//   * no second-order correction, no watchdog, no fixed-mu fallback (line-search failure enters phase 1).
// PARITY PINNING: pinned against the reference's own known answers on test/mytoy.nl
// (test/InterfaceTest.cpp:122-189) and cross-checked against scipy on the synthetic shapes;
// everything at n>=100 is parity-unpinned by the reference itself.
//
// Build: see oracle/Makefile  (g++ -O3 -march=native -shared -fPIC).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr double kInf = 1e19;

// TNLPSolver::ReturnStatus values (src/Interfaces/BonTNLPSolver.hpp:29-44)
enum Status {
  iterationLimit = -3, doesNotConverge = -8, computationError = -2, illDefinedProblem = -4,
  solvedOptimal = 1, solvedOptimalTol = 2, provenInfeasible = 3, unbounded = 4
};

struct OModel {
  int n, m;
  const double *xl, *xu, *gl, *gu;  // model bounds (children override xl/xu)
  const double *c;                  // m+1 constants (0 = objective)
  const double *A;                  // (m+1) x n dense linear part
  const int *qidx;                  // m+1: index of the dense quadratic block or -1
  const double *Q;                  // nq x n x n symmetric, convention 1/2 x'Qx
  // general nonlinear element terms phi_t(x[vars]) added to function term_fn[t]: postfix programs (include/bnlp.h)
  int n_terms = 0;
  const int *term_fn = nullptr, *term_var_start = nullptr, *term_var = nullptr, *term_op_start = nullptr,
            *term_op = nullptr, *term_iarg = nullptr;
  const double *term_darg = nullptr;
};

// second-order forward mode over the local variables of a term (<= 24: BNLP_MAX_WIDE_TERM_VARS of include/bnlp.h; only the
// leading nv x nv part of a stack entry is ever touched)
constexpr int TVm = 24;
struct HD { double v; double g[TVm]; double h[TVm][TVm]; };
enum { OP_CONST = 0, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_POWC, OP_POW, OP_EXP, OP_LOG, OP_SQRT, OP_SIN,
       OP_COS, OP_SQR, OP_ABS, OP_ATAN, OP_ASIN, OP_ACOS };
static void hd_unary(HD &a, int nv, double f, double f1, double f2) {
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) a.h[i][j] = f2 * a.g[i] * a.g[j] + f1 * a.h[i][j];
  for (int i = 0; i < nv; ++i) a.g[i] *= f1;
  a.v = f;
}
static void hd_mul(HD &a, const HD &b, int nv) {
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) a.h[i][j] = a.v * b.h[i][j] + b.v * a.h[i][j] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  for (int i = 0; i < nv; ++i) a.g[i] = a.v * b.g[i] + b.v * a.g[i];
  a.v *= b.v;
}
static HD eval_term(const OModel &M, int t, const double *x) {
  const int v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
  HD st[16];
  int sp = 0;
  for (int p = M.term_op_start[t]; p < M.term_op_start[t + 1]; ++p) {
    const int op = M.term_op[p];
    if (op == OP_CONST || op == OP_VAR) {
      HD &z = st[sp++];
      z.v = 0.0;
      for (int i = 0; i < nv; ++i) { z.g[i] = 0.0; for (int j = 0; j < nv; ++j) z.h[i][j] = 0.0; }
      if (op == OP_CONST) z.v = M.term_darg[p];
      else { z.v = x[M.term_var[v0 + M.term_iarg[p]]]; z.g[M.term_iarg[p]] = 1.0; }
    } else if (op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV || op == OP_POW) {
      HD &b = st[sp - 1];
      HD &a = st[sp - 2];
      --sp;
      if (op == OP_ADD || op == OP_SUB) {
        const double sg = op == OP_ADD ? 1.0 : -1.0;
        a.v += sg * b.v;
        for (int i = 0; i < nv; ++i) { a.g[i] += sg * b.g[i]; for (int j = 0; j < nv; ++j) a.h[i][j] += sg * b.h[i][j]; }
      } else if (op == OP_MUL) hd_mul(a, b, nv);
      else if (op == OP_DIV) { const double r = 1.0 / b.v; hd_unary(b, nv, r, -r * r, 2 * r * r * r); hd_mul(a, b, nv); }
      else { const double la = std::log(a.v); hd_unary(a, nv, la, 1.0 / a.v, -1.0 / (a.v * a.v)); hd_mul(a, b, nv);
             const double e = std::exp(a.v); hd_unary(a, nv, e, e, e); }
    } else {
      HD &a = st[sp - 1];
      const double v = a.v;
      switch (op) {
        case OP_NEG: hd_unary(a, nv, -v, -1.0, 0.0); break;
        case OP_SQR: hd_unary(a, nv, v * v, 2 * v, 2.0); break;
        case OP_POWC: { const double p_ = M.term_darg[p];
                        hd_unary(a, nv, std::pow(v, p_), p_ == 0 ? 0.0 : p_ * std::pow(v, p_ - 1),
                                 (p_ == 0 || p_ == 1) ? 0.0 : p_ * (p_ - 1) * std::pow(v, p_ - 2)); break; }
        case OP_EXP: { const double e = std::exp(v); hd_unary(a, nv, e, e, e); break; }
        case OP_LOG: hd_unary(a, nv, std::log(v), 1.0 / v, -1.0 / (v * v)); break;
        case OP_SQRT: { const double q = std::sqrt(v); hd_unary(a, nv, q, 0.5 / q, -0.25 / (q * v)); break; }
        case OP_SIN: hd_unary(a, nv, std::sin(v), std::cos(v), -std::sin(v)); break;
        case OP_COS: hd_unary(a, nv, std::cos(v), -std::sin(v), -std::cos(v)); break;
        case OP_ABS: hd_unary(a, nv, std::fabs(v), v >= 0 ? 1.0 : -1.0, 0.0); break;
        case OP_ATAN: { const double q = 1.0 / (1.0 + v * v); hd_unary(a, nv, std::atan(v), q, -2.0 * v * q * q); break; }
        case OP_ASIN: { const double q = 1.0 / std::sqrt(1.0 - v * v); hd_unary(a, nv, std::asin(v), q, v * q * q * q); break; }
        case OP_ACOS: { const double q = 1.0 / std::sqrt(1.0 - v * v); hd_unary(a, nv, std::acos(v), -q, -v * q * q * q); break; }
      }
    }
  }
  return st[0];
}

struct OOpt {
  double tol = 1e-8, acceptable_tol = 1e-6;
  int max_iter = 3000, acceptable_iter = 15;
  double dual_inf_tol = 1.0, constr_viol_tol = 1e-4, compl_inf_tol = 1e-4;
  double bound_relax = 1e-8, bound_push = 0.01, bound_frac = 0.01;
  // warm_start_init_point yes (BonIpoptSolver.cpp:415-419): warm_start_bound_push / _bound_frac / _slack_bound_push /
  // _slack_bound_frac / _mult_bound_push, all 1e-3 (options_list_ipopt_content.tex, "Warm Start")
  double warm_push = 1e-3;
  double tau_min = 0.99, sigma_min = 1e-6, sigma_max = 100, mu_min = 1e-11, mu_max_fact = 1e3, mu_max = 1e5;
  double s_max = 100, kappa_sigma = 1e10;
  double gamma_theta = 1e-4, gamma_phi = 1e-8, eta_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, delta = 1.0;
  double alpha_min_frac = 0.05, alpha_red = 0.5;
  double theta_max_fact = 1e4, theta_min_fact = 1e-4;
  double scal_max_grad = 100.0;
  double req_infeas_red = 0.1;       // required_infeasibility_reduction (Bonmin)
  double infeas_ytol = 1e8, infeas_ctol = 1e-3;
  double infeas_tol = 1e-7;          // phase-1 optimum above this (scaled max violation) => infeasible
  int max_phase1 = 8;
  double mu_safeguard = 1e-4;
};

static int g_verbose = 0;
struct OStats { int iters_main = 0; int iters = 0, phase1_iters = 0, phase1_entries = 0, nfact = 0; };

// dense Cholesky of the lower triangle of a (row-major N x N); returns false on pivot <= 0
bool chol(std::vector<double> &a, int N) {
  for (int j = 0; j < N; ++j) {
    double *aj = &a[(size_t)j * N];
    double d = aj[j];
    for (int k = 0; k < j; ++k) d -= aj[k] * aj[k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    aj[j] = d;
    double inv = 1.0 / d;
    for (int i = j + 1; i < N; ++i) {
      double *ai = &a[(size_t)i * N];
      double v = ai[j];
      for (int k = 0; k < j; ++k) v -= ai[k] * aj[k];
      ai[j] = v * inv;
    }
  }
  return true;
}
void fwd(const std::vector<double> &L, int N, double *b) {  // L w = b
  for (int i = 0; i < N; ++i) {
    const double *li = &L[(size_t)i * N];
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= li[k] * b[k];
    b[i] = v / li[i];
  }
}
void bwd(const std::vector<double> &L, int N, double *b) {  // L' x = b
  for (int i = N - 1; i >= 0; --i) {
    double v = b[i] / L[(size_t)i * N + i];
    b[i] = v;
    for (int k = 0; k < i; ++k) b[k] -= L[(size_t)i * N + k] * v;
  }
}

struct Row { int fn; double sgn; double b; double sc; };

struct Solver {
  const OModel &M;
  const OOpt &o;
  int n, m, N, R, E;
  std::vector<double> L, U, Lr, Ur;   // bounds (N), relaxed
  std::vector<char> fixed, hasL, hasU;
  std::vector<Row> rows, eqs;
  std::vector<double> val, grad;      // raw function values (m+1) and gradients (m+1) x n
  std::vector<double> uq, gd, du;     // Q_k dx, grad_k.dx, dx.Q_k dx per function
  std::vector<double> tval;           // (m+1) sum of element-term values per function at x (val = quadratic part + tval)
  std::vector<HD> tcur;               // per term: value / gradient / Hessian at x
  std::vector<double> x, s, v, zL, zU, ye;
  std::vector<double> h, J, ce, Ae;   // scaled one-sided rows, R x N Jacobian; equalities E x N
  double sf = 1.0;
  bool phase1 = false;
  OStats st;

  Solver(const OModel &M_, const OOpt &o_) : M(M_), o(o_) {}

  void eval_full() {  // raw val/grad of every function at x[0..n)
    for (int k = 0; k <= m; ++k) {
      const double *a = M.A + (size_t)k * n;
      double *g = &grad[(size_t)k * n];
      double vv = M.c[k];
      if (M.qidx[k] >= 0) {
        const double *Q = M.Q + (size_t)M.qidx[k] * n * n;
        for (int i = 0; i < n; ++i) {
          double u = 0;
          const double *qi = Q + (size_t)i * n;
          for (int j = 0; j < n; ++j) u += qi[j] * x[j];
          g[i] = a[i] + u;
          vv += x[i] * (a[i] + 0.5 * u);
        }
      } else {
        for (int i = 0; i < n; ++i) { g[i] = a[i]; vv += a[i] * x[i]; }
      }
      val[k] = vv;
    }
    add_terms(+1.0);
  }
  // evaluate the element terms at x and add (sgn=+1) them to val/grad; remove_terms() undoes the previous addition
  void add_terms(double) {
    tval.assign(m + 1, 0.0);
    tcur.resize(M.n_terms);
    for (int t = 0; t < M.n_terms; ++t) {
      tcur[t] = eval_term(M, t, x.data());
      const int fn = M.term_fn[t], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
      tval[fn] += tcur[t].v;
      for (int a = 0; a < nv; ++a) grad[(size_t)fn * n + M.term_var[v0 + a]] += tcur[t].g[a];
    }
    for (int k = 0; k <= m; ++k) val[k] += tval[k];
  }
  void remove_terms() {
    for (int t = 0; t < M.n_terms; ++t) {
      const int fn = M.term_fn[t], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
      for (int a = 0; a < nv; ++a) grad[(size_t)fn * n + M.term_var[v0 + a]] -= tcur[t].g[a];
    }
    for (int k = 0; k <= m; ++k) val[k] -= tval[k];
  }
  void build_rows_values() {  // h, J, ce, Ae from raw val/grad (+ elastic column)
    for (int r = 0; r < R; ++r) {
      const Row &w = rows[r];
      h[r] = w.sc * w.sgn * val[w.fn] - (phase1 ? x[n] : 0.0);
      double *Jr = &J[(size_t)r * N];
      const double *g = &grad[(size_t)w.fn * n];
      double f = w.sc * w.sgn;
      for (int j = 0; j < n; ++j) Jr[j] = f * g[j];
      Jr[n] = -1.0;
    }
    for (int e = 0; e < E; ++e) {
      const Row &w = eqs[e];
      ce[e] = w.sc * (val[w.fn] - w.b);
      double *Ar = &Ae[(size_t)e * N];
      const double *g = &grad[(size_t)w.fn * n];
      for (int j = 0; j < n; ++j) Ar[j] = w.sc * g[j];
      Ar[n] = 0.0;
    }
  }
  double viol_scaled() const {  // max over rows of scaled violation (without the elastic)
    double t = 0;
    for (int r = 0; r < R; ++r) {
      const Row &w = rows[r];
      t = std::max(t, w.sc * w.sgn * val[w.fn] - w.b);
    }
    return t;
  }
  double viol_l1() const {
    double t = 0;
    for (int r = 0; r < R; ++r) {
      const Row &w = rows[r];
      t += std::max(0.0, w.sc * w.sgn * val[w.fn] - w.b);
    }
    return t;
  }
  // warm = the caller's multipliers [z_L | z_U | lambda] (DUAL0, unscaled) start the solve, every one of them pushed to at
  // least warm_push (warm_start_mult_bound_push), slacks pushed by warm_push instead of bound_push; used for the FIRST
  // initialisation only (a phase switch restarts from unit multipliers as before)
  const double *dual0 = nullptr;
  void init_slacks_mults(bool warm = false) {
    warm = warm && dual0;
    for (int r = 0; r < R; ++r) {
      double sl = rows[r].b - h[r];
      double push = (warm ? o.warm_push : o.bound_push) * std::max(1.0, std::fabs(rows[r].b));
      s[r] = std::max(sl, push);
      v[r] = warm ? std::max(o.warm_push, std::max(0.0, rows[r].sgn * dual0[2 * n + rows[r].fn - 1]) * sf / rows[r].sc) : 1.0;
    }
    for (int j = 0; j < N; ++j) {
      const bool w = warm && j < n;
      zL[j] = (!fixed[j] && hasL[j]) ? (w ? std::max(o.warm_push, dual0[j] * sf) : 1.0) : 0.0;
      zU[j] = (!fixed[j] && hasU[j]) ? (w ? std::max(o.warm_push, dual0[n + j] * sf) : 1.0) : 0.0;
    }
    for (int e = 0; e < E; ++e) ye[e] = warm ? dual0[2 * n + eqs[e].fn - 1] * sf / eqs[e].sc : 0.0;
  }

  int run(const double *xl, const double *xu, const double *x0, double *obj, int *iters, double *xout,
          double *gout, double *dout);
};

int Solver::run(const double *xl, const double *xu, const double *x0, double *obj, int *iters, double *xout,
                double *gout, double *dout) {
  n = M.n; m = M.m; N = n + 1;
  L.assign(N, 0); U.assign(N, 0); Lr.assign(N, 0); Ur.assign(N, 0);
  fixed.assign(N, 0); hasL.assign(N, 0); hasU.assign(N, 0);
  x.assign(N, 0);
  val.assign(m + 1, 0); grad.assign((size_t)(m + 1) * n, 0);
  auto finish = [&](int status, double objv) {
    *obj = objv; *iters = st.iters;
    return status;
  };
  auto write_point = [&](bool with_duals) {
    for (int j = 0; j < n; ++j) xout[j] = std::min(std::max(x[j], xl[j]), xu[j]);  // honor_original_bounds
    for (int i = 0; i < m; ++i) gout[i] = val[i + 1];
    for (int k = 0; k < 2 * n + m; ++k) dout[k] = 0.0;
    if (!with_duals) return;
    for (int j = 0; j < n; ++j) { dout[j] = zL[j] / sf; dout[n + j] = zU[j] / sf; }
    for (int r = 0; r < R; ++r) dout[2 * n + rows[r].fn - 1] += rows[r].sgn * v[r] * rows[r].sc / sf;
    for (int e = 0; e < E; ++e) dout[2 * n + eqs[e].fn - 1] += ye[e] * eqs[e].sc / sf;
  };
  auto orig_violation = [&]() {
    double t = 0;
    for (int i = 0; i < m; ++i) {
      if (M.gu[i] < kInf) t = std::max(t, val[i + 1] - M.gu[i]);
      if (M.gl[i] > -kInf) t = std::max(t, M.gl[i] - val[i + 1]);
    }
    return t;
  };

  // ---- bounds, fixed variables (fixed_variable_treatment make_parameter, L3682)
  bool all_fixed = true;
  for (int j = 0; j < n; ++j) {
    L[j] = xl[j]; U[j] = xu[j];
    hasL[j] = L[j] > -kInf; hasU[j] = U[j] < kInf;
    if (U[j] < L[j]) {  // inconsistent bounds: the child is trivially infeasible
      for (int q = 0; q < n; ++q) x[q] = std::min(std::max(x0[q], std::min(xl[q], xu[q])), std::max(xl[q], xu[q]));
      eval_full();
      R = E = 0; write_point(false);
      return finish(provenInfeasible, std::max(L[j] - U[j], orig_violation()));
    }
    fixed[j] = (U[j] <= L[j]);
    if (U[j] - L[j] > 1e-5) all_fixed = false;
  }
  // ---- all-variables-fixed short cut (BonTNLPSolver.cpp:78-161): evaluate at x_l, feasibility tol 1e-7
  if (all_fixed) {
    for (int j = 0; j < n; ++j) x[j] = L[j];
    eval_full();
    R = E = 0;
    int status = solvedOptimal;
    for (int i = 0; i < m; ++i)
      if (val[i + 1] - M.gl[i] < -1e-7 || val[i + 1] - M.gu[i] > 1e-7) status = provenInfeasible;
    write_point(false);
    for (int j = 0; j < n; ++j) xout[j] = L[j];
    return finish(status, val[0]);
  }
  // elastic variable t (index n): fixed at 0 outside phase 1
  L[n] = 0; U[n] = kInf; hasL[n] = 1; hasU[n] = 0; fixed[n] = 1; x[n] = 0;

  // ---- rows: one-sided normal form
  rows.clear(); eqs.clear();
  for (int i = 0; i < m; ++i) {
    bool lo = M.gl[i] > -kInf, hi = M.gu[i] < kInf;
    if (lo && hi && M.gl[i] == M.gu[i]) { eqs.push_back({i + 1, 1.0, M.gl[i], 1.0}); continue; }
    if (hi) rows.push_back({i + 1, 1.0, M.gu[i], 1.0});
    if (lo) rows.push_back({i + 1, -1.0, -M.gl[i], 1.0});
  }
  R = (int)rows.size(); E = (int)eqs.size();

  // ---- relaxed bounds (bound_relax_factor, L3647) and start point pushed inside (L2871, L2850)
  for (int j = 0; j < n; ++j) {
    Lr[j] = L[j]; Ur[j] = U[j];
    if (fixed[j]) { x[j] = L[j]; continue; }
    if (hasL[j]) Lr[j] = L[j] - o.bound_relax * std::max(1.0, std::fabs(L[j]));
    if (hasU[j]) Ur[j] = U[j] + o.bound_relax * std::max(1.0, std::fabs(U[j]));
    double xx = x0[j];
    const double bp = dual0 ? o.warm_push : o.bound_push, bf = dual0 ? o.warm_push : o.bound_frac;
    if (hasL[j] && hasU[j]) {
      double pl = std::min(bp * std::max(1.0, std::fabs(Lr[j])), bf * (Ur[j] - Lr[j]));
      double pu = std::min(bp * std::max(1.0, std::fabs(Ur[j])), bf * (Ur[j] - Lr[j]));
      xx = std::min(std::max(xx, Lr[j] + pl), Ur[j] - pu);
    } else if (hasL[j]) {
      xx = std::max(xx, Lr[j] + bp * std::max(1.0, std::fabs(Lr[j])));
    } else if (hasU[j]) {
      xx = std::min(xx, Ur[j] - bp * std::max(1.0, std::fabs(Ur[j])));
    }
    x[j] = xx;
  }
  Lr[n] = -o.bound_relax; Ur[n] = kInf;
  eval_full();

  // ---- gradient-based scaling (nlp_scaling_method gradient-based L3767, max gradient 100 L3761)
  auto gmax = [&](int fn) {
    double g = 0;
    for (int j = 0; j < n; ++j) if (!fixed[j]) g = std::max(g, std::fabs(grad[(size_t)fn * n + j]));
    return g;
  };
  { double g = gmax(0); sf = (g > o.scal_max_grad) ? o.scal_max_grad / g : 1.0; }
  for (auto &w : rows) { double g = gmax(w.fn); w.sc = (g > o.scal_max_grad) ? o.scal_max_grad / g : 1.0; }
  for (auto &w : eqs) { double g = gmax(w.fn); w.sc = (g > o.scal_max_grad) ? o.scal_max_grad / g : 1.0; }
  for (auto &w : rows) {
    double b = w.sc * w.sgn * (w.sgn > 0 ? M.gu[w.fn - 1] : M.gl[w.fn - 1]);
    w.b = b + o.bound_relax * std::max(1.0, std::fabs(b));
  }

  s.assign(R, 0); v.assign(R, 0); zL.assign(N, 0); zU.assign(N, 0); ye.assign(E, 0);
  h.assign(R, 0); J.assign((size_t)R * N, 0); ce.assign(E, 0); Ae.assign((size_t)E * N, 0);
  uq.assign((size_t)(m + 1) * n, 0); gd.assign(m + 1, 0); du.assign(m + 1, 0);
  phase1 = false;
  build_rows_values();
  init_slacks_mults(true);

  std::vector<double> H((size_t)N * N), rhs0(N), rhs1(N), dx(N), ds(R), dv(R), dzL(N), dzU(N), dye(E);
  std::vector<double> Sx(N), Ss(R), rc(R), gL(N), tmpR(R), Y, S, w0(N), w1(N), de0(E), de1(E);
  double delta_w_last = 0.0, theta_init = -1, mu_max = -1;
  double th_hist[4] = {0, 0, 0, 0}, a_last[2] = {1, 1};
  int acceptable_cnt = 0;
  int p1_iters_this = 0;
  auto theta_of = [&]() {
    double t = 0;
    for (int r = 0; r < R; ++r) t += std::fabs(h[r] + s[r] - rows[r].b);
    for (int e = 0; e < E; ++e) t += std::fabs(ce[e]);
    return t;
  };
  auto enter_phase1 = [&]() {
    phase1 = true; st.phase1_entries++; p1_iters_this = 0;
    double vmax = std::max(0.0, viol_scaled());
    fixed[n] = 0;
    double T = std::max(1.0, vmax);
    Lr[n] = -T; L[n] = -T;  // t may go negative: finds a strictly interior point quickly when one exists
    x[n] = 1.1 * vmax + 0.1 * T;
    build_rows_values();
    init_slacks_mults();
    theta_init = -1; mu_max = -1; acceptable_cnt = 0;
  };
  auto leave_phase1 = [&]() {
    phase1 = false; fixed[n] = 1; x[n] = 0; st.iters_main = 0;
    build_rows_values();
    init_slacks_mults();
    theta_init = -1; mu_max = -1; acceptable_cnt = 0;
  };

  for (;; ++st.iters) {
    const double of = phase1 ? 0.0 : sf;  // objective weight: phase 1 minimises t only
    // ---- residuals
    double th = 0;
    for (int r = 0; r < R; ++r) { rc[r] = h[r] + s[r] - rows[r].b; th += std::fabs(rc[r]); }
    for (int e = 0; e < E; ++e) th += std::fabs(ce[e]);
    if (theta_init < 0) theta_init = th;
    double dual_inf = 0, compl_inf = 0, sum_mult = 0, sum_z = 0; int nz = 0;
    for (int j = 0; j < N; ++j) {
      if (fixed[j]) { gL[j] = 0; continue; }
      double g = (j < n ? of * grad[j] : (phase1 ? 1.0 : 0.0));
      for (int r = 0; r < R; ++r) g += J[(size_t)r * N + j] * v[r];
      for (int e = 0; e < E; ++e) g += Ae[(size_t)e * N + j] * ye[e];
      gL[j] = g;  // gradient of f + y'h (+ye'c), bound multipliers excluded
      dual_inf = std::max(dual_inf, std::fabs(g - zL[j] + zU[j]));
      if (hasL[j]) { compl_inf = std::max(compl_inf, (x[j] - Lr[j]) * zL[j]); sum_z += zL[j]; nz++; }
      if (hasU[j]) { compl_inf = std::max(compl_inf, (Ur[j] - x[j]) * zU[j]); sum_z += zU[j]; nz++; }
    }
    double cviol = 0, cviol_unsc = 0;
    for (int r = 0; r < R; ++r) {
      compl_inf = std::max(compl_inf, s[r] * v[r]);
      sum_z += v[r]; nz++; sum_mult += v[r];
      cviol = std::max(cviol, std::fabs(rc[r]));
      cviol_unsc = std::max(cviol_unsc, std::fabs(rc[r]) / rows[r].sc);
    }
    for (int e = 0; e < E; ++e) {
      sum_mult += std::fabs(ye[e]);
      cviol = std::max(cviol, std::fabs(ce[e]));
      cviol_unsc = std::max(cviol_unsc, std::fabs(ce[e]) / eqs[e].sc);
    }
    double s_d = std::max(o.s_max, (sum_mult + sum_z) / std::max(1, R + E + nz)) / o.s_max;
    double s_c = std::max(o.s_max, sum_z / std::max(1, nz)) / o.s_max;
    double E0 = std::max(std::max(dual_inf / s_d, cviol), compl_inf / s_c);
    double avg_compl = 0;
    {
      double t = 0;
      for (int j = 0; j < N; ++j) {
        if (fixed[j]) continue;
        if (hasL[j]) t += (x[j] - Lr[j]) * zL[j];
        if (hasU[j]) t += (Ur[j] - x[j]) * zU[j];
      }
      for (int r = 0; r < R; ++r) t += s[r] * v[r];
      avg_compl = nz ? t / nz : 0.0;
    }
    if (mu_max < 0) mu_max = std::min(o.mu_max, o.mu_max_fact * std::max(avg_compl, 1e-300));

    // ---- termination tests (tol L2689; dual_inf_tol/constr_viol_tol/compl_inf_tol L2659/L2647/L2641)
    const double osc = phase1 ? 1.0 : sf;
    bool conv = E0 <= o.tol && dual_inf / osc <= o.dual_inf_tol && cviol_unsc <= o.constr_viol_tol &&
                compl_inf / osc <= o.compl_inf_tol;
    bool acc = E0 <= o.acceptable_tol && cviol_unsc <= 1e-2 && compl_inf / osc <= 1e-2;
    acceptable_cnt = acc ? acceptable_cnt + 1 : 0;
    if (phase1) {
      // leave as soon as the violation has dropped enough (required_infeasibility_reduction 0.1)
      // leave as soon as x is feasible for the (relaxed) rows: stricter than Bonmin's
      // required_infeasibility_reduction 0.1, which made main phase / phase 1 cycle on infeasible children
      if (viol_l1() <= 0.0) { leave_phase1(); continue; }
      if (conv || acceptable_cnt >= o.acceptable_iter) {
        double vmax = viol_scaled();
        // strict convergence: the phase-1 optimum is known to tol, any violation above infeas_tol is real.  Merely
        // "acceptable" termination leaves an error of E0 in the phase-1 optimum: a violation that this error could
        // explain (thin or tangential feasible set) is not a proof, the child goes back to the main phase instead
        const double vthr = conv ? o.infeas_tol : std::max(o.infeas_tol, 10.0 * E0);
        if (vmax > vthr) {  // Infeasible_Problem_Detected -> provenInfeasible
          write_point(false);
          return finish(provenInfeasible, orig_violation());  // BonTMINLP2TNLP.cpp:503-505
        }
        leave_phase1();
        continue;
      }
    } else {
      if (conv) { write_point(true); return finish(solvedOptimal, val[0]); }
      if (acceptable_cnt >= o.acceptable_iter) { write_point(true); return finish(solvedOptimalTol, val[0]); }
    }
    if (st.iters >= o.max_iter) { write_point(true); return finish(iterationLimit, val[0]); }
    if (phase1) { st.phase1_iters++; p1_iters_this++; }

    // ---- expect_infeasible_problem heuristics (BonIpoptSolver.cpp:277-282): enter phase 1 early when the
    // multipliers explode, or when the infeasibility stalls (< 10 % reduction over 3 iterations)
    if (!phase1) {
      const double vl1 = viol_l1();
      th_hist[st.iters_main % 4] = vl1;
      const double vold = th_hist[(st.iters_main + 1) % 4];
      bool stall = st.iters_main >= 3 && vold > 0.0 && vl1 > 0.9 * vold && a_last[0] < 0.3 && a_last[1] < 0.3;
      st.iters_main++;
      double ymax = 0;
      for (int r = 0; r < R; ++r) ymax = std::max(ymax, v[r]);
      if (st.phase1_entries > 0) stall = false;  // a phase-1 pass already produced a feasible point
      if (th > o.infeas_ctol && (ymax > o.infeas_ytol || stall) && vl1 > 0.0) {
        if (st.phase1_entries >= o.max_phase1) { write_point(true); return finish(iterationLimit, val[0]); }
        enter_phase1();
        continue;
      }
    }

    // ---- condensed matrix H = W + Sigma_x + delta_w I + J' Sigma_s J   (lower triangle)
    for (int j = 0; j < N; ++j) {
      double sx = 0;
      if (!fixed[j]) {
        if (hasL[j]) sx += zL[j] / (x[j] - Lr[j]);
        if (hasU[j]) sx += zU[j] / (Ur[j] - x[j]);
      }
      Sx[j] = sx;
    }
    for (int r = 0; r < R; ++r) Ss[r] = v[r] / s[r];
    double delta_w = 0.0;
    for (int attempt = 0;; ++attempt) {
      std::fill(H.begin(), H.end(), 0.0);
      auto addQ = [&](int fn, double wgt) {
        if (wgt == 0.0 || M.qidx[fn] < 0) return;
        const double *Q = M.Q + (size_t)M.qidx[fn] * n * n;
        for (int i = 0; i < n; ++i) {
          double *Hi = &H[(size_t)i * N];
          const double *qi = Q + (size_t)i * n;
          for (int j = 0; j <= i; ++j) Hi[j] += wgt * qi[j];
        }
      };
      addQ(0, of);
      for (int r = 0; r < R; ++r) addQ(rows[r].fn, v[r] * rows[r].sc * rows[r].sgn);
      for (int e = 0; e < E; ++e) addQ(eqs[e].fn, ye[e] * eqs[e].sc);
      if (M.n_terms > 0) {
        std::vector<double> wfn(m + 1, 0.0);
        wfn[0] = of;
        for (int r = 0; r < R; ++r) wfn[rows[r].fn] += v[r] * rows[r].sc * rows[r].sgn;
        for (int e = 0; e < E; ++e) wfn[eqs[e].fn] += ye[e] * eqs[e].sc;
        for (int t = 0; t < M.n_terms; ++t) {
          const int fn = M.term_fn[t], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
          if (wfn[fn] == 0.0) continue;
          for (int a = 0; a < nv; ++a)
            for (int b = 0; b < nv; ++b) {
              const int i = M.term_var[v0 + a], j = M.term_var[v0 + b];
              if (j <= i) H[(size_t)i * N + j] += wfn[fn] * tcur[t].h[a][b];
            }
        }
      }
      for (int r = 0; r < R; ++r) {
        const double *Jr = &J[(size_t)r * N];
        double w = Ss[r];
        for (int i = 0; i < N; ++i) {
          double a = w * Jr[i];
          if (a == 0.0) continue;
          double *Hi = &H[(size_t)i * N];
          for (int j = 0; j <= i; ++j) Hi[j] += a * Jr[j];
        }
      }
      for (int j = 0; j < N; ++j) H[(size_t)j * N + j] += Sx[j] + delta_w;
      for (int j = 0; j < N; ++j)
        if (fixed[j]) {
          for (int k = 0; k < j; ++k) H[(size_t)j * N + k] = 0;
          for (int i = j + 1; i < N; ++i) H[(size_t)i * N + j] = 0;
          H[(size_t)j * N + j] = 1.0;
        }
      st.nfact++;
      if (chol(H, N)) break;
      // inertia correction (first_hessian_perturbation 1e-4 L4125, x100 then x8, /3 between iterations)
      if (delta_w == 0.0) delta_w = (delta_w_last == 0.0) ? 1e-4 : std::max(1e-20, delta_w_last / 3.0);
      else delta_w *= (delta_w_last == 0.0 ? 100.0 : 8.0);
      if (delta_w > 1e20 || attempt > 40) { write_point(true); return finish(computationError, val[0]); }
    }
    if (delta_w > 0) delta_w_last = delta_w;

    // ---- right-hand sides: rhs(mu) = rhs0 + mu rhs1
    for (int j = 0; j < N; ++j) {
      if (fixed[j]) { rhs0[j] = rhs1[j] = 0; continue; }
      double g0 = (j < n ? of * grad[j] : (phase1 ? 1.0 : 0.0));
      for (int e = 0; e < E; ++e) g0 += Ae[(size_t)e * N + j] * ye[e];
      double a0 = -g0, a1 = 0;
      for (int r = 0; r < R; ++r) {
        double Jrj = J[(size_t)r * N + j];
        a0 -= Jrj * (Ss[r] * rc[r]);
        a1 -= Jrj / s[r];
      }
      if (hasL[j]) a1 += 1.0 / (x[j] - Lr[j]);
      if (hasU[j]) a1 -= 1.0 / (Ur[j] - x[j]);
      rhs0[j] = a0; rhs1[j] = a1;
    }
    // solves (equalities through the Schur complement S = Ae H^-1 Ae')
    w0 = rhs0; w1 = rhs1;
    fwd(H, N, w0.data()); fwd(H, N, w1.data());
    if (E > 0) {
      Y.assign((size_t)E * N, 0); S.assign((size_t)E * E, 0);
      for (int e = 0; e < E; ++e) {
        for (int j = 0; j < N; ++j) Y[(size_t)e * N + j] = fixed[j] ? 0.0 : Ae[(size_t)e * N + j];
        fwd(H, N, &Y[(size_t)e * N]);
      }
      for (int a = 0; a < E; ++a)
        for (int b = 0; b <= a; ++b) {
          double t = 0;
          for (int j = 0; j < N; ++j) t += Y[(size_t)a * N + j] * Y[(size_t)b * N + j];
          S[(size_t)a * E + b] = t;
        }
      for (int a = 0; a < E; ++a) S[(size_t)a * E + a] += 1e-14 * (1.0 + S[(size_t)a * E + a]);
      if (!chol(S, E)) { write_point(true); return finish(computationError, val[0]); }
      for (int e = 0; e < E; ++e) {
        double t0 = ce[e], t1 = 0;
        for (int j = 0; j < N; ++j) { t0 += Y[(size_t)e * N + j] * w0[j]; t1 += Y[(size_t)e * N + j] * w1[j]; }
        de0[e] = t0; de1[e] = t1;
      }
      fwd(S, E, de0.data()); bwd(S, E, de0.data());
      fwd(S, E, de1.data()); bwd(S, E, de1.data());
      for (int e = 0; e < E; ++e)
        for (int j = 0; j < N; ++j) { w0[j] -= Y[(size_t)e * N + j] * de0[e]; w1[j] -= Y[(size_t)e * N + j] * de1[e]; }
    }
    bwd(H, N, w0.data()); bwd(H, N, w1.data());  // w0 = d0, w1 = d1

    // ---- probing oracle for mu (mu_strategy adaptive, mu_oracle probing): sigma = (mu_aff/mu)^3
    auto direction = [&](double mu) {
      for (int j = 0; j < N; ++j) dx[j] = fixed[j] ? 0.0 : (w0[j] + mu * w1[j]);
      for (int e = 0; e < E; ++e) dye[e] = de0.empty() ? 0.0 : de0[e] + mu * de1[e];
      for (int r = 0; r < R; ++r) {
        const double *Jr = &J[(size_t)r * N];
        double jd = 0;
        for (int j = 0; j < N; ++j) jd += Jr[j] * dx[j];
        ds[r] = -rc[r] - jd;
        dv[r] = mu / s[r] - v[r] - Ss[r] * ds[r];
      }
      for (int j = 0; j < N; ++j) {
        dzL[j] = dzU[j] = 0;
        if (fixed[j]) continue;
        if (hasL[j]) { double sl = x[j] - Lr[j]; dzL[j] = mu / sl - zL[j] - zL[j] / sl * dx[j]; }
        if (hasU[j]) { double su = Ur[j] - x[j]; dzU[j] = mu / su - zU[j] + zU[j] / su * dx[j]; }
      }
    };
    auto ftb = [&](double tau, double &ap, double &ad) {  // fraction to the boundary
      ap = 1.0; ad = 1.0;
      for (int j = 0; j < N; ++j) {
        if (fixed[j]) continue;
        if (hasL[j]) {
          if (dx[j] < 0) ap = std::min(ap, -tau * (x[j] - Lr[j]) / dx[j]);
          if (dzL[j] < 0) ad = std::min(ad, -tau * zL[j] / dzL[j]);
        }
        if (hasU[j]) {
          if (dx[j] > 0) ap = std::min(ap, tau * (Ur[j] - x[j]) / dx[j]);
          if (dzU[j] < 0) ad = std::min(ad, -tau * zU[j] / dzU[j]);
        }
      }
      for (int r = 0; r < R; ++r) {
        if (ds[r] < 0) ap = std::min(ap, -tau * s[r] / ds[r]);
        if (dv[r] < 0) ad = std::min(ad, -tau * v[r] / dv[r]);
      }
    };
    double mu;
    {
      direction(0.0);
      double ap, ad;
      ftb(1.0, ap, ad);
      double t = 0;
      for (int j = 0; j < N; ++j) {
        if (fixed[j]) continue;
        if (hasL[j]) t += (x[j] + ap * dx[j] - Lr[j]) * (zL[j] + ad * dzL[j]);
        if (hasU[j]) t += (Ur[j] - x[j] - ap * dx[j]) * (zU[j] + ad * dzU[j]);
      }
      for (int r = 0; r < R; ++r) t += (s[r] + ap * ds[r]) * (v[r] + ad * dv[r]);
      double mu_aff = nz ? t / nz : 0.0;
      double sigma = avg_compl > 0 ? std::pow(mu_aff / avg_compl, 3.0) : o.sigma_min;
      sigma = std::min(std::max(sigma, o.sigma_min), o.sigma_max);
      mu = std::min(std::max(sigma * avg_compl, o.mu_min), mu_max);
      // safeguard: do not drive complementarity far below the current primal/dual infeasibility
      mu = std::max(mu, std::min(avg_compl, o.mu_safeguard * std::max(cviol, dual_inf / s_d)));
    }
    direction(mu);
    double tau = std::max(o.tau_min, 1.0 - mu), ap_max, ad;
    ftb(tau, ap_max, ad);

    // ---- per-function directional data: u_k = Q_k dx, gd_k = grad_k.dx, du_k = dx.u_k
    for (int k = 0; k <= m; ++k) {
      const double *g = &grad[(size_t)k * n];
      double a = 0, b = 0;
      double *u = &uq[(size_t)k * n];
      if (M.qidx[k] >= 0) {
        const double *Q = M.Q + (size_t)M.qidx[k] * n * n;
        for (int i = 0; i < n; ++i) {
          const double *qi = Q + (size_t)i * n;
          double t = 0;
          for (int j = 0; j < n; ++j) t += qi[j] * dx[j];
          u[i] = t;
          b += dx[i] * t;
        }
      }
      for (int i = 0; i < n; ++i) a += g[i] * dx[i];
      gd[k] = a; du[k] = b;
    }
    // closed-form trial values use the quadratic part only: remove the terms' share of grad.dx
    for (int t = 0; t < M.n_terms; ++t) {
      const int fn = M.term_fn[t], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
      for (int a = 0; a < nv; ++a) gd[fn] -= tcur[t].g[a] * dx[M.term_var[v0 + a]];
    }
    std::vector<double> tvt(m + 1, 0.0), xt(M.n_terms > 0 ? n : 0);
    auto trial_terms = [&](double a) {   // sum of element-term values per function at x + a dx
      std::fill(tvt.begin(), tvt.end(), 0.0);
      if (M.n_terms == 0) return;
      for (int j = 0; j < n; ++j) xt[j] = x[j] + a * dx[j];
      for (int t = 0; t < M.n_terms; ++t) tvt[M.term_fn[t]] += eval_term(M, t, xt.data()).v;
    };

    // ---- filter line search on (theta, phi_mu) against the current iterate
    auto fval = [&](int k, double a) { return val[k] - tval[k] + a * gd[k] + 0.5 * a * a * du[k] + tvt[k]; };
    auto merit = [&](double a, double &theta, double &phi) {
      double t = phase1 ? x[n] + a * dx[n] : 0.0;
      trial_terms(a);
      theta = 0;
      phi = phase1 ? t : sf * fval(0, a);
      for (int r = 0; r < R; ++r) {
        const Row &w = rows[r];
        double hv = w.sc * w.sgn * fval(w.fn, a) - t;
        double sv = s[r] + a * ds[r];
        theta += std::fabs(hv + sv - w.b);
        phi -= mu * std::log(sv);
      }
      for (int e = 0; e < E; ++e) {
        const Row &w = eqs[e];
        theta += std::fabs(w.sc * (fval(w.fn, a) - w.b));
      }
      for (int j = 0; j < N; ++j) {
        if (fixed[j]) continue;
        double xv = x[j] + a * dx[j];
        if (hasL[j]) phi -= mu * std::log(xv - Lr[j]);
        if (hasU[j]) phi -= mu * std::log(Ur[j] - xv);
      }
    };
    double theta0, phi0;
    merit(0.0, theta0, phi0);
    double gd0_total = gd[0];
    for (int t = 0; t < M.n_terms; ++t)
      if (M.term_fn[t] == 0) {
        const int v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
        for (int a = 0; a < nv; ++a) gd0_total += tcur[t].g[a] * dx[M.term_var[v0 + a]];
      }
    double dphi = phase1 ? dx[n] : sf * gd0_total;
    for (int r = 0; r < R; ++r) dphi -= mu * ds[r] / s[r];
    for (int j = 0; j < N; ++j) {
      if (fixed[j]) continue;
      if (hasL[j]) dphi -= mu * dx[j] / (x[j] - Lr[j]);
      if (hasU[j]) dphi += mu * dx[j] / (Ur[j] - x[j]);
    }
    const double theta_min = o.theta_min_fact * std::max(1.0, theta_init);
    const double theta_max = o.theta_max_fact * std::max(1.0, theta_init);
    double amin = o.gamma_theta;
    if (dphi < 0) {
      amin = std::min(amin, o.gamma_phi * theta0 / (-dphi));
      if (theta0 <= theta_min) amin = std::min(amin, o.delta * std::pow(theta0, o.s_theta) / std::pow(-dphi, o.s_phi));
    }
    amin *= o.alpha_min_frac;
    double alpha = ap_max;
    bool accepted = false;
    for (int bt = 0; bt < 60; ++bt) {
      double th_t, ph_t;
      merit(alpha, th_t, ph_t);
      bool ok;
      const double fuzz = 10.0 * 2.220446049250313e-16 * std::fabs(phi0);
      bool sw = dphi < 0 && theta0 <= theta_min &&
                alpha * std::pow(-dphi, o.s_phi) > o.delta * std::pow(theta0, o.s_theta);
      if (sw) ok = ph_t - phi0 <= o.eta_phi * alpha * dphi + fuzz;
      else ok = (th_t <= (1 - o.gamma_theta) * theta0) || (ph_t - phi0 <= -o.gamma_phi * theta0 + fuzz);
      if (!(th_t <= theta_max)) ok = false;
      if (!std::isfinite(ph_t)) ok = false;
      if (ok) { accepted = true; break; }
      alpha *= o.alpha_red;
      if (alpha < amin) break;
    }
    if (!accepted) {
      if (phase1 || st.phase1_entries >= o.max_phase1) {
        // Restoration_Failed -> iterationLimit (BonIpoptSolver.cpp:308-311)
        write_point(true);
        return finish(iterationLimit, val[0]);
      }
      enter_phase1();
      continue;
    }

    if (g_verbose) std::printf("it %3d p1=%d E0=%.2e th=%.3e mu=%.2e ap_max=%.2e a=%.2e ad=%.2e dw=%.1e f=%.8g t=%.3e\n", st.iters, (int)phase1, E0, th, mu, ap_max, alpha, ad, delta_w, val[0], x[n]);
    // ---- accept the step
    a_last[1] = a_last[0]; a_last[0] = alpha;
    for (int j = 0; j < N; ++j) {
      if (fixed[j]) continue;
      x[j] += alpha * dx[j];
      zL[j] += ad * dzL[j];
      zU[j] += ad * dzU[j];
    }
    for (int r = 0; r < R; ++r) { s[r] += alpha * ds[r]; v[r] += ad * dv[r]; }
    for (int e = 0; e < E; ++e) ye[e] += alpha * dye[e];
    if (M.n_terms > 0) remove_terms();
    for (int k = 0; k <= m; ++k) {
      val[k] += alpha * gd[k] + 0.5 * alpha * alpha * du[k];
      if (M.qidx[k] >= 0) {
        double *g = &grad[(size_t)k * n];
        const double *u = &uq[(size_t)k * n];
        for (int i = 0; i < n; ++i) g[i] += alpha * u[i];
      }
    }
    if (M.n_terms > 0) add_terms(+1.0);
    build_rows_values();
    // keep bound multipliers within [mu/(kappa s), kappa mu/s] (kappa_sigma 1e10, L3022)
    for (int j = 0; j < N; ++j) {
      if (fixed[j]) continue;
      if (hasL[j]) { double sl = x[j] - Lr[j]; zL[j] = std::min(std::max(zL[j], mu / (o.kappa_sigma * sl)), o.kappa_sigma * mu / sl); }
      if (hasU[j]) { double su = Ur[j] - x[j]; zU[j] = std::min(std::max(zU[j], mu / (o.kappa_sigma * su)), o.kappa_sigma * mu / su); }
    }
    for (int r = 0; r < R; ++r) v[r] = std::min(std::max(v[r], mu / (o.kappa_sigma * s[r])), o.kappa_sigma * mu / s[r]);
  }
}

}  // namespace

extern "C" {
void oracle_set_verbose(int v) { g_verbose = v; }

// Solve B child relaxations (same model, per-child bounds XL/XU [B x n], shared start point x0 [n]
// or per-child X0 [B x n] when x0_per_child != 0) on `threads` host threads.
// DUAL0 (optional, n-shared or per child): multipliers [z_L | z_U | lambda] of a warm start (IpoptSolver with
// warm_start_init_point yes, BonIpoptSolver.cpp:369-419; ReOptimizeTNLP after setWarmStart).
// Outputs follow TMINLP2TNLP's result layout (BonTMINLP2TNLP.cpp:474-509).
int oracle_solve_batch_warm(int n, int m, const double *xl, const double *xu, const double *gl, const double *gu,
                            const double *c, const double *A, const int *qidx, const double *Q, int n_terms,
                            const int *term_fn, const int *term_var_start, const int *term_var, const int *term_op_start,
                            const int *term_op, const int *term_iarg, const double *term_darg, int B,
                            const double *XL, const double *XU, const double *X0, int x0_per_child, const double *DUAL0,
                            int dual0_per_child, int threads, int max_iter, double *OBJ, int *STATUS, int *ITERS, double *X,
                            double *G, double *DUALS, int *P1ITERS) {
  OModel M{n, m, xl, xu, gl, gu, c, A, qidx, Q};
  M.n_terms = n_terms; M.term_fn = term_fn; M.term_var_start = term_var_start; M.term_var = term_var;
  M.term_op_start = term_op_start; M.term_op = term_op; M.term_iarg = term_iarg; M.term_darg = term_darg;
  OOpt opt;
  if (max_iter > 0) opt.max_iter = max_iter;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      Solver S(M, opt);
      const double *x0 = x0_per_child ? X0 + (size_t)b * n : X0;
      if (DUAL0) S.dual0 = dual0_per_child ? DUAL0 + (size_t)b * (2 * n + m) : DUAL0;
      STATUS[b] = S.run(XL + (size_t)b * n, XU + (size_t)b * n, x0, &OBJ[b], &ITERS[b], X + (size_t)b * n,
                        G + (size_t)b * m, DUALS + (size_t)b * (2 * n + m));
      if (P1ITERS) P1ITERS[b] = S.st.phase1_iters;
    }
  };
  if (threads <= 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto &t : th) t.join();
  return 0;
}

int oracle_solve_batch(int n, int m, const double *xl, const double *xu, const double *gl, const double *gu,
                       const double *c, const double *A, const int *qidx, const double *Q, int n_terms,
                       const int *term_fn, const int *term_var_start, const int *term_var, const int *term_op_start,
                       const int *term_op, const int *term_iarg, const double *term_darg, int B,
                       const double *XL, const double *XU, const double *X0, int x0_per_child, int threads,
                       int max_iter, double *OBJ, int *STATUS, int *ITERS, double *X, double *G, double *DUALS,
                       int *P1ITERS) {
  return oracle_solve_batch_warm(n, m, xl, xu, gl, gu, c, A, qidx, Q, n_terms, term_fn, term_var_start, term_var,
                                 term_op_start, term_op, term_iarg, term_darg, B, XL, XU, X0, x0_per_child, nullptr, 0,
                                 threads, max_iter, OBJ, STATUS, ITERS, X, G, DUALS, P1ITERS);
}

// ---- B x OsiTMINLPInterface::getOuterApproximation (src/Interfaces/BonOsiTMINLPInterface.cpp:1924-2107, cleanNnz :1881-1919)
// on `threads` host threads, one point per thread at a time: the CPU baseline of the batched OA kernels (bench.py) and a second
// checker beside oracle/struct_eval.py (same semantics, restated line by line from the same reference lines).
// jac_start / jac_col: structural columns of every row (m + 1 entries / nnz), obj_cols: structural columns of the objective
// gradient are all n (the reference loops over every column, :2063).  Slot s < mnl = the s-th nonlinear row (nl_rows[s]),
// slot mnl = the objective cut.  Fixed-capacity output: slot s of point b writes at cap_start[s] of its block of `cap` entries.
// DUALS (B x (2n+m)) may be null.  get_obj as in the reference (0 none, 1 if the objective is nonlinear, 2 always).
int oracle_oa_batch(int n, int m, const double *xl, const double *xu, const double *gl, const double *gu, const double *c,
                    const double *A, const int *qidx, const double *Q, int n_terms, const int *term_fn,
                    const int *term_var_start, const int *term_var, const int *term_op_start, const int *term_op,
                    const int *term_iarg, const double *term_darg, const int *jac_start, const int *jac_col, int mnl,
                    const int *nl_rows, int obj_nonlinear, const int *cap_start, int B, const double *X, const double *DUALS,
                    int get_obj, int threads, int *CNT, int *COLS, double *VALS, double *LB, double *UB) {
  OModel M{n, m, xl, xu, gl, gu, c, A, qidx, Q};
  M.n_terms = n_terms; M.term_fn = term_fn; M.term_var_start = term_var_start; M.term_var = term_var;
  M.term_op_start = term_op_start; M.term_op = term_op; M.term_iarg = term_iarg; M.term_darg = term_darg;
  const double DMAX = 1.7976931348623157e308, tiny = 1e-8, very_tiny = 1e-17, rhs_relax = 1e-8, infty_ = 1e19;
  const int slots = mnl + 1, cap = cap_start[slots];
  const bool want_obj = get_obj == 2 || (get_obj && obj_nonlinear);
  auto clean_nnz = [&](double value, double colLower, double colUpper, double rowLower, double rowUpper, double colsol,
                       double &lb, double &ub, double vtiny) {
    if (std::fabs(value) >= tiny) return true;
    if (std::fabs(value) < vtiny) return false;
    const bool colUpBounded = colUpper < infty_, colLoBounded = colLower > -infty_;
    const bool rowNotLoBounded = rowLower <= -infty_, rowNotUpBounded = rowUpper >= infty_;
    const bool pos = value > 0;
    if (colUpBounded && pos && rowNotUpBounded) { lb += value * (colsol - colUpper); return false; }
    if (colUpBounded && !pos && rowNotLoBounded) { ub += value * (colsol - colUpper); return false; }
    if (colLoBounded && !pos && rowNotUpBounded) { lb += value * (colsol - colLower); return false; }
    if (colLoBounded && pos && rowNotLoBounded) { ub += value * (colsol - colLower); return false; }
    return true;
  };
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<double> grad(n), u(n);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      const double *x = X + (size_t)b * n;
      // value and dense gradient of function k at x
      auto eval_fn = [&](int k, double &val) {
        val = c[k];
        const double *a = A + (size_t)k * n;
        for (int j = 0; j < n; ++j) { grad[j] = a[j]; val += a[j] * x[j]; }
        if (qidx[k] >= 0) {
          const double *Qk = Q + (size_t)qidx[k] * n * n;
          double q = 0.0;
          for (int i = 0; i < n; ++i) {
            const double *row = Qk + (size_t)i * n;
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc += row[j] * x[j];
            grad[i] += acc;
            q += x[i] * acc;
          }
          val += 0.5 * q;
        }
        for (int t = 0; t < n_terms; ++t) {
          if (term_fn[t] != k) continue;
          const HD r = eval_term(M, t, x);
          const int v0 = term_var_start[t], nv = term_var_start[t + 1] - v0;
          val += r.v;
          for (int q2 = 0; q2 < nv; ++q2) grad[term_var[v0 + q2]] += r.g[q2];
        }
      };
      for (int s = 0; s < mnl; ++s) {
        const int i = nl_rows[s], o = b * slots + s;
        double g;
        eval_fn(i + 1, g);
        double lb = gl[i] > -infty_ ? gl[i] - g : -DMAX, ub = gu[i] < infty_ ? gu[i] - g : DMAX;
        if (DUALS && gl[i] > -DMAX && gu[i] < DMAX) {                              // :1968-1974
          const double d = DUALS[(size_t)b * (2 * n + m) + 2 * n + i];
          if (d >= 0) lb = -DMAX;
          if (d <= 0) ub = DMAX;
        }
        int cnt = 0;
        int *cols = COLS + (size_t)b * cap + cap_start[s];
        double *vals = VALS + (size_t)b * cap + cap_start[s];
        for (int e = jac_start[i]; e < jac_start[i + 1]; ++e) {                    // :1981-1999
          const int j = jac_col[e];
          const double v = grad[j];
          if (clean_nnz(v, xl[j], xu[j], gl[i], gu[i], x[j], lb, ub, very_tiny)) {
            cols[cnt] = j; vals[cnt] = v; ++cnt;
            if (lb > -DMAX) lb += v * x[j];
            if (ub < DMAX) ub += v * x[j];
          }
        }
        if (lb > DMAX) lb -= rhs_relax * std::max(std::fabs(lb), 1.0);             // dead code in the reference (:2043)
        if (ub < DMAX) ub += rhs_relax * std::max(std::fabs(ub), 1.0);
        CNT[o] = cnt; LB[o] = lb; UB[o] = ub;
      }
      {
        const int o = b * slots + mnl;
        CNT[o] = -1; LB[o] = -DMAX; UB[o] = DMAX;
        if (want_obj) {                                                            // :2053-2105
          double f;
          eval_fn(0, f);
          double lb = -f, ub = -f;
          int cnt = 0;
          int *cols = COLS + (size_t)b * cap + cap_start[mnl];
          double *vals = VALS + (size_t)b * cap + cap_start[mnl];
          for (int j = 0; j < n; ++j) {
            if (clean_nnz(grad[j], xl[j], xu[j], -DMAX, 0.0, x[j], lb, ub, 1e-15)) {
              cols[cnt] = j; vals[cnt] = grad[j]; ++cnt;
              lb += grad[j] * x[j]; ub += grad[j] * x[j];
            }
          }
          cols[cnt] = n; vals[cnt] = -1.0; ++cnt;
          CNT[o] = cnt; LB[o] = -DMAX; UB[o] = ub;
        }
      }
    }
  };
  if (threads <= 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto &t : th) t.join();
  return 0;
}

}  // extern "C"
