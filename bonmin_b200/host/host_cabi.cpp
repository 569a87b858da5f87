// C entry points of the host layer (used by bench.py, tests and the `bnlp_bb` command): the B-BB driver with a
// batch-solve callback.  In the product the callback is bnlp_solve_batch on a CUDA engine (bnlp_bb_run_engine in
// bnlp_api.cu); tests may pass the CPU oracle to check that both give the same MINLP optimum.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <sstream>

#include "../../include/bnlp.h"
#include "B200Linearizer.hpp"
#include "B200Oa.hpp"
#include "B200Solver.hpp"

extern "C" {

int bnlp_bb_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                   const double *gu, const double *x0, bnlp_batch_solve_cb cb, void *ctx, const bnlp_bb_options *o,
                   bnlp_bb_result *res, double *best_x) {
  return bnlp_bb_run_cb2(n, m, var_type, xl, xu, gl, gu, x0, cb, nullptr, ctx, o, res, best_x);
}

int bnlp_bb_run_cb2(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                    const double *gu, const double *x0, bnlp_batch_solve_cb cb, bnlp_tqp_solve_cb qp_cb, void *ctx,
                    const bnlp_bb_options *o, bnlp_bb_result *res, double *best_x) {
  return bnlp_bb_run_cb3(n, m, var_type, xl, xu, gl, gu, x0, cb, qp_cb, nullptr, ctx, o, res, best_x);
}

int bnlp_bb_run_cb3(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                    const double *gu, const double *x0, bnlp_batch_solve_cb cb, bnlp_tqp_solve_cb qp_cb,
                    bnlp_children_solve_cb ch_cb, void *ctx, const bnlp_bb_options *o, bnlp_bb_result *res, double *best_x) {
  using namespace Bonmin;
  Ipopt::SmartPtr<TMINLP2TNLP> root_holder = new TMINLP2TNLP();   // SmartPtr-managed objects live on the heap
  TMINLP2TNLP &root = *root_holder;
  root.n = n; root.m = m;
  root.var_types.assign(var_type, var_type + n);
  root.x_l_.assign(xl, xl + n); root.x_u_.assign(xu, xu + n);
  root.g_l_.assign(gl, gl + m); root.g_u_.assign(gu, gu + m);
  root.x_init_.assign(x0, x0 + n);
  BbOptions opt;
  if (o) {
    if (o->number_strong_branch > 0) opt.number_strong_branch = o->number_strong_branch;
    if (o->number_before_trust > 0) opt.number_before_trust = o->number_before_trust;
    if (o->node_limit > 0) opt.node_limit = o->node_limit;
    if (o->integer_tolerance > 0) opt.integer_tolerance = o->integer_tolerance;
    if (o->cutoff_decr > 0) opt.cutoff_decr = o->cutoff_decr;
    if (o->time_limit > 0) opt.time_limit = o->time_limit;
    opt.log_level = o->log_level;
    if (o->parallel_nodes > 0) opt.parallel_nodes = o->parallel_nodes;
    opt.variable_selection = o->variable_selection;
    if (o->priorities) opt.priorities.assign(o->priorities, o->priorities + n);
    opt.nlp_failure_behavior = o->nlp_failure_behavior;
    opt.number_look_ahead = o->number_look_ahead;
    if (o->incumbent_x) { opt.incumbent_x.assign(o->incumbent_x, o->incumbent_x + n); opt.incumbent_obj = o->incumbent_obj; }
  }
  TqpSolveFn qfn;
  if (qp_cb)
    qfn = [qp_cb, ctx](const double *xs, const double *du, int B, const double *XL, const double *XU, double *OBJ, int *ST,
                       int *IT, double *X) { return qp_cb(ctx, xs, du, B, XL, XU, OBJ, ST, IT, X); };
  BatchSolveFn fn = [cb, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT,
                              double *X, double *G, double *D) { return cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D); };
  ChildrenSolveFn cfn;
  if (ch_cb)
    cfn = [ch_cb, ctx](int P, const double *PXL, const double *PXU, int B, const int *cp, const int *cc, const int *cw,
                       const double *cb2, const double *X0, double *OBJ, int *ST, int *IT, double *X) {
      return ch_cb(ctx, P, PXL, PXU, B, cp, cc, cw, cb2, X0, OBJ, ST, IT, X);
    };
  BbResult r = bb_run(root, fn, opt, qfn, cfn);
  res->status = r.status; res->nodes = r.nodes; res->nlp_solves = r.nlp_solves; res->sb_batches = r.sb_batches;
  res->sb_children = r.sb_children; res->served_from_sb_cache = r.served_from_sb_cache; res->failed_nodes = r.failed_nodes;
  res->ipm_iterations = r.ipm_iterations; res->best_obj = r.best_obj; res->best_bound = r.best_bound; res->seconds = r.seconds;
  if (best_x && !r.best_x.empty()) memcpy(best_x, r.best_x.data(), sizeof(double) * n);
  return r.status == 3 ? -1 : 0;   // the arithmetic provider failed: an error, not a search result
}

// Exercises the TNLPSolver contract on stub providers: options, output levels, error objects, and — through a MOCK
// Ipopt::TNLP that records every call — that (Re)OptimizeTNLP drives a TNLP purely through its virtuals in the order
// get_nlp_info -> get_bounds_info -> get_starting_point -> finalize_solution, asks for multipliers only when a warm start is
// enabled, and that the interface sequence of test/InterfaceTest.cpp:165-189 (solve, getWarmStart, change bounds,
// setWarmStart, resolve) reaches the warm-start provider with the saved primal and dual point.  Returns 0, or the source
// line of the first expectation that failed.
namespace {
struct MockTNLP : public Ipopt::TNLP {
  std::string log;                       // one letter per virtual call: i(nfo) b(ounds) s(tart) S(tart with duals) f(inalize)
  int n = 2, m = 1;
  std::vector<double> xl{1.0, 0.0}, xu{4.0, 3.0}, x0{2.0, 1.0}, duals;   // duals empty: get_starting_point(init_z) fails
  Ipopt::SolverReturn status = Ipopt::UNASSIGNED;
  std::vector<double> x, zl, zu, lam, g;
  double obj = 0;
  bool get_nlp_info(Ipopt::Index &nn, Ipopt::Index &mm, Ipopt::Index &nj, Ipopt::Index &nh, IndexStyleEnum &st) override {
    log += 'i'; nn = n; mm = m; nj = 2; nh = 1; st = C_STYLE; return true;
  }
  bool get_bounds_info(Ipopt::Index nn, double *l, double *u, Ipopt::Index mm, double *gl, double *gu) override {
    log += 'b';
    for (int j = 0; j < nn; ++j) { l[j] = xl[j]; u[j] = xu[j]; }
    for (int i = 0; i < mm; ++i) { gl[i] = -1e19; gu[i] = 1.0; }
    return true;
  }
  bool get_starting_point(Ipopt::Index nn, bool init_x, double *xx, bool init_z, double *z_L, double *z_U, Ipopt::Index mm,
                          bool init_lambda, double *lambda) override {
    log += (init_z || init_lambda) ? 'S' : 's';
    if ((init_z || init_lambda) && duals.empty()) return false;
    if (init_x) for (int j = 0; j < nn; ++j) xx[j] = x0[j];
    if (init_z) for (int j = 0; j < nn; ++j) { z_L[j] = duals[j]; z_U[j] = duals[nn + j]; }
    if (init_lambda) for (int i = 0; i < mm; ++i) lambda[i] = duals[2 * nn + i];
    return true;
  }
  bool eval_f(Ipopt::Index, const double *, bool, double &) override { log += 'E'; return false; }
  bool eval_grad_f(Ipopt::Index, const double *, bool, double *) override { log += 'E'; return false; }
  bool eval_g(Ipopt::Index, const double *, bool, Ipopt::Index, double *) override { log += 'E'; return false; }
  bool eval_jac_g(Ipopt::Index, const double *, bool, Ipopt::Index, Ipopt::Index, Ipopt::Index *, Ipopt::Index *, double *) override {
    log += 'E'; return false;
  }
  void finalize_solution(Ipopt::SolverReturn st, Ipopt::Index nn, const double *xx, const double *z_L, const double *z_U,
                         Ipopt::Index mm, const double *gg, const double *lambda, double o, const Ipopt::IpoptData *d,
                         Ipopt::IpoptCalculatedQuantities *q) override {
    log += (d == nullptr && q == nullptr) ? 'f' : 'F';
    status = st; obj = o;
    x.assign(xx, xx + nn); zl.assign(z_L, z_L + nn); zu.assign(z_U, z_U + nn); g.assign(gg, gg + mm); lam.assign(lambda, lambda + mm);
  }
};
}  // namespace

// B-OA (B200Oa.cpp) over caller-supplied arithmetic and master
int bnlp_oa_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *x0,
                   bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts, bnlp_oa_extract_cb extract, bnlp_master_mip_cb master,
                   void *ctx, const bnlp_oa_run_options *o, bnlp_oa_run_result *res, double *best_x) {
  using namespace Bonmin;
  if (!solve || !cuts || !extract || !master || !res || n <= 0 || m < 0) return -1;
  Ipopt::SmartPtr<TMINLP2TNLP> holder = new TMINLP2TNLP();
  TMINLP2TNLP &root = *holder;
  root.n = n; root.m = m;
  root.var_types.assign(var_type, var_type + n);
  root.x_l_.assign(xl, xl + n); root.x_u_.assign(xu, xu + n);
  if (x0) root.x_init_.assign(x0, x0 + n);
  OaOptions opt;
  if (o) {
    if (o->cutoff_increment > 0) opt.cutoff_increment = o->cutoff_increment;
    if (o->gap_tol > 0) opt.gap_tol = o->gap_tol;
    if (o->time_limit > 0) opt.time_limit = o->time_limit;
    if (o->max_passes > 0) opt.max_passes = o->max_passes;
    opt.add_master_point = o->no_master_point ? 0 : 1;
  }
  BatchSolveFn sf = [solve, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *STATUS,
                                 int *ITERS, double *X, double *G, double *DUALS) {
    return solve(ctx, B, XL, XU, X0, OBJ, STATUS, ITERS, X, G, DUALS);
  };
  const size_t rows_cap = (size_t)m + 1, nnz_row = (size_t)n + 1;
  OaCutsFn cf = [cuts, ctx, rows_cap, nnz_row](int B, const double *X, const double *duals, CutRows &out) {
    const int cap_cuts = (int)(B * rows_cap), cap_nnz = (int)(B * rows_cap * nnz_row);
    out.row_start.assign(cap_cuts + 1, 0); out.col.resize(cap_nnz); out.val.resize(cap_nnz);
    out.lb.resize(cap_cuts); out.ub.resize(cap_cuts); out.src_row.resize(cap_cuts);
    int nc = 0;
    const int rc = cuts(ctx, B, X, duals, &nc, out.row_start.data(), out.col.data(), out.val.data(), out.lb.data(), out.ub.data(),
                        out.src_row.data(), cap_cuts, cap_nnz);
    if (rc) return rc;
    if (nc < 0 || nc > cap_cuts) return -2;
    out.row_start.resize(nc + 1); out.lb.resize(nc); out.ub.resize(nc); out.src_row.resize(nc);
    return 0;
  };
  OaExtractFn ef = [extract, ctx, rows_cap, nnz_row, n](const double *x, const double *duals, LpRelaxation &lp) {
    const int cap_rows = (int)rows_cap, cap_nnz = (int)(rows_cap * nnz_row);
    lp.row_start.assign(cap_rows + 1, 0); lp.col.resize(cap_nnz); lp.val.resize(cap_nnz);
    lp.row_lb.resize(cap_rows); lp.row_ub.resize(cap_rows); lp.obj.assign(n + 1, 0.0);
    int nr = 0;
    const int rc = extract(ctx, x, duals, &nr, &lp.has_alpha, &lp.obj_offset, lp.row_start.data(), lp.col.data(), lp.val.data(),
                           lp.row_lb.data(), lp.row_ub.data(), lp.obj.data(), cap_rows, cap_nnz);
    if (rc) return rc;
    if (nr < 0 || nr > cap_rows) return -2;
    lp.n_cols = n + (lp.has_alpha ? 1 : 0);
    lp.row_start.resize(nr + 1); lp.row_lb.resize(nr); lp.row_ub.resize(nr); lp.obj.resize(lp.n_cols);
    return 0;
  };
  MasterMipFn mf = [master, ctx](int nc, int nr, const int *rs, const int *cl, const double *vl, const double *rlb,
                                 const double *rub, const double *clb, const double *cub, const double *ob, double off,
                                 const int *isint, double cutoff, double maxs, double *x, double *bound) {
    return master(ctx, nc, nr, rs, cl, vl, rlb, rub, clb, cub, ob, off, isint, cutoff, maxs, x, bound);
  };
  OaResult R = oa_run(root, sf, cf, ef, mf, opt);
  res->status = R.status; res->passes = R.passes; res->nlp_solves = R.nlp_solves; res->cuts = R.cuts; res->oa_points = R.oa_points;
  res->best_obj = R.best_obj; res->lower_bound = R.lower_bound; res->seconds = R.seconds; res->master_seconds = R.master_seconds;
  if (best_x && !R.best_x.empty()) std::memcpy(best_x, R.best_x.data(), sizeof(double) * n);
  return R.status < 0 ? R.status : 0;
}

namespace {
// shared by the OA entry points: C callbacks -> the std::function providers of B200Oa.hpp, cut rows -> flat CSR outputs
Bonmin::BatchSolveFn oa_wrap_solve(bnlp_batch_solve_cb solve, void *ctx) {
  return [solve, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *STATUS, int *ITERS, double *X,
                      double *G, double *DUALS) { return solve(ctx, B, XL, XU, X0, OBJ, STATUS, ITERS, X, G, DUALS); };
}
Bonmin::OaCutsFn oa_wrap_cuts(bnlp_oa_cuts_cb cuts, void *ctx, int n, int m) {
  const size_t rows_cap = (size_t)m + 1, nnz_row = (size_t)n + 1;
  return [cuts, ctx, rows_cap, nnz_row](int B, const double *X, const double *duals, Bonmin::CutRows &out) {
    const int cc = (int)(B * rows_cap), cn = (int)(B * rows_cap * nnz_row);
    out.row_start.assign(cc + 1, 0); out.col.resize(cn); out.val.resize(cn);
    out.lb.resize(cc); out.ub.resize(cc); out.src_row.resize(cc);
    int nc = 0;
    const int rc = cuts(ctx, B, X, duals, &nc, out.row_start.data(), out.col.data(), out.val.data(), out.lb.data(), out.ub.data(),
                        out.src_row.data(), cc, cn);
    if (rc) return rc;
    if (nc < 0 || nc > cc) return -2;
    out.row_start.resize(nc + 1); out.lb.resize(nc); out.ub.resize(nc); out.src_row.resize(nc);
    return 0;
  };
}
Bonmin::MasterMipFn oa_wrap_master(bnlp_master_mip_cb f, void *ctx) {
  return [f, ctx](int nc, int nr, const int *rs, const int *cl, const double *vl, const double *rlb, const double *rub,
                  const double *clb, const double *cub, const double *ob, double off, const int *isint, double co, double maxs,
                  double *x, double *bound) { return f(ctx, nc, nr, rs, cl, vl, rlb, rub, clb, cub, ob, off, isint, co, maxs, x, bound); };
}
int oa_rows_out(const Bonmin::CutRows &R, int *new_row_start, int *new_col, double *new_val, double *new_lb, double *new_ub,
                int *new_src, int cap_cuts, int cap_nnz) {
  const int nc = R.size(), nnz = (int)R.col.size();
  if (nc > cap_cuts || nnz > cap_nnz) return -2;
  if (new_row_start) std::memcpy(new_row_start, R.row_start.data(), sizeof(int) * (nc + 1));
  if (nc) {
    std::memcpy(new_col, R.col.data(), sizeof(int) * nnz); std::memcpy(new_val, R.val.data(), sizeof(double) * nnz);
    std::memcpy(new_lb, R.lb.data(), sizeof(double) * nc); std::memcpy(new_ub, R.ub.data(), sizeof(double) * nc);
    if (new_src) std::memcpy(new_src, R.src_row.data(), sizeof(int) * nc);
  }
  return 0;
}
Bonmin::OaLp oa_lp_in(int n_cols, int n_rows, const int *row_start, const int *col, const double *val, const double *row_lb,
                      const double *row_ub, const double *col_lb, const double *col_ub, const double *obj, double obj_offset) {
  Bonmin::OaLp lp;
  lp.n_cols = n_cols; lp.obj_offset = obj_offset;
  lp.row_start.assign(row_start, row_start + n_rows + 1);
  lp.col.assign(col, col + row_start[n_rows]); lp.val.assign(val, val + row_start[n_rows]);
  lp.row_lb.assign(row_lb, row_lb + n_rows); lp.row_ub.assign(row_ub, row_ub + n_rows);
  lp.col_lb.assign(col_lb, col_lb + n_cols); lp.col_ub.assign(col_ub, col_ub + n_cols); lp.obj.assign(obj, obj + n_cols);
  return lp;
}
}  // namespace

// OaNlpOptim::generateCuts (B200Oa.cpp: oa_nlp_optim) over caller-supplied arithmetic
int bnlp_oa_nlp_optim_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *node_lb,
                         const double *node_ub, const double *x_lp, bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts, void *ctx,
                         bnlp_nlp_optim_result *res, double *x_nlp, double *duals_out, int *new_row_start, int *new_col,
                         double *new_val, double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz) {
  using namespace Bonmin;
  if (!solve || !cuts || !res || n <= 0 || m < 0) return -1;
  Ipopt::SmartPtr<TMINLP2TNLP> holder = new TMINLP2TNLP();
  TMINLP2TNLP &root = *holder;
  root.n = n; root.m = m;
  root.var_types.assign(var_type, var_type + n);
  root.x_l_.assign(xl, xl + n); root.x_u_.assign(xu, xu + n);
  OaNlpOptimResult R = oa_nlp_optim(root, node_lb, node_ub, x_lp, oa_wrap_solve(solve, ctx), oa_wrap_cuts(cuts, ctx, n, m));
  res->nlp_status = R.nlp_status; res->cuts = R.cuts.size(); res->new_nnz = (int)R.cuts.col.size();
  res->has_solution = R.has_solution; res->obj = R.obj;
  if (R.rc) return R.rc;
  if (x_nlp) std::memcpy(x_nlp, R.x.data(), sizeof(double) * n);
  if (duals_out) std::memcpy(duals_out, R.duals.data(), sizeof(double) * (2 * (size_t)n + m));
  return oa_rows_out(R.cuts, new_row_start, new_col, new_val, new_lb, new_ub, new_src, cap_cuts, cap_nnz);
}

// EcpCuts::generateCuts (B200Oa.cpp: oa_ecp_rounds) over caller-supplied arithmetic and LP solver
int bnlp_ecp_rounds_cb(int n, int m, int n_cols, int n_rows, const int *row_start, const int *col, const double *val,
                       const double *row_lb, const double *row_ub, const double *col_lb, const double *col_ub, const double *obj,
                       double obj_offset, const int *is_integer, const double *duals, int rounds, double abs_tol, double rel_tol,
                       bnlp_oa_cuts_cb cuts, bnlp_violation_cb violation, bnlp_master_mip_cb lp_solve, void *ctx,
                       bnlp_ecp_result *res, double *x_out, int *new_row_start, int *new_col, double *new_val, double *new_lb,
                       double *new_ub, int *new_src, int cap_cuts, int cap_nnz) {
  using namespace Bonmin;
  if (!cuts || !violation || !lp_solve || !res || n <= 0 || m < 0 || n_cols < n || n_rows < 0) return -1;
  OaLp lp = oa_lp_in(n_cols, n_rows, row_start, col, val, row_lb, row_ub, col_lb, col_ub, obj, obj_offset);
  OaViolationFn vf = [violation, ctx](const double *x, double o, double *v) { return violation(ctx, x, o, v); };
  OaEcpResult R = oa_ecp_rounds(n, m, lp, is_integer, duals, rounds, abs_tol, rel_tol, oa_wrap_cuts(cuts, ctx, n, m), vf,
                                oa_wrap_master(lp_solve, ctx));
  res->rounds = R.rounds; res->cuts = R.cs.size(); res->installed = R.installed; res->new_nnz = (int)R.cs.col.size();
  res->lp_infeasible = R.lp_infeasible; res->violation = R.violation; res->obj = R.obj;
  if (R.rc) return R.rc;
  if (x_out) std::memcpy(x_out, R.x.data(), sizeof(double) * n_cols);
  return oa_rows_out(R.cs, new_row_start, new_col, new_val, new_lb, new_ub, new_src, cap_cuts, cap_nnz);
}

// OaFeasibilityChecker::performOa (B200Oa.cpp: oa_feas_check) over caller-supplied arithmetic and LP solver
int bnlp_oa_feas_check_cb(int n, int m, const int *var_type, const double *xl, const double *xu, int n_cols, int n_rows,
                          const int *row_start, const int *col, const double *val, const double *row_lb, const double *row_ub,
                          const double *col_lb, const double *col_ub, const double *obj, double obj_offset, const double *x_lp,
                          double cutoff, const bnlp_oa_feas_options *o, bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts,
                          bnlp_master_mip_cb lp_solve, void *ctx, bnlp_oa_feas_result *res, double *incumbent_x, double *x_lp_out,
                          int *new_row_start, int *new_col, double *new_val, double *new_lb, double *new_ub, int *new_src,
                          int cap_cuts, int cap_nnz) {
  using namespace Bonmin;
  if (!solve || !cuts || !lp_solve || !res || n <= 0 || m < 0 || n_cols < n || n_rows < 0) return -1;
  Ipopt::SmartPtr<TMINLP2TNLP> holder = new TMINLP2TNLP();
  TMINLP2TNLP &root = *holder;
  root.n = n; root.m = m;
  root.var_types.assign(var_type, var_type + n);
  root.x_l_.assign(xl, xl + n); root.x_u_.assign(xu, xu + n);
  OaLp lp;
  lp.n_cols = n_cols; lp.obj_offset = obj_offset;
  lp.row_start.assign(row_start, row_start + n_rows + 1);
  lp.col.assign(col, col + row_start[n_rows]); lp.val.assign(val, val + row_start[n_rows]);
  lp.row_lb.assign(row_lb, row_lb + n_rows); lp.row_ub.assign(row_ub, row_ub + n_rows);
  lp.col_lb.assign(col_lb, col_lb + n_cols); lp.col_ub.assign(col_ub, col_ub + n_cols); lp.obj.assign(obj, obj + n_cols);
  OaFeasOptions opt;
  if (o) {
    if (o->integer_tolerance > 0) opt.integer_tolerance = o->integer_tolerance;
    if (o->cutoff_increment > 0) opt.cutoff_increment = o->cutoff_increment;
    if (o->max_passes > 0) opt.max_passes = o->max_passes;
  }
  BatchSolveFn sf = [solve, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *STATUS,
                                 int *ITERS, double *X, double *G, double *DUALS) {
    return solve(ctx, B, XL, XU, X0, OBJ, STATUS, ITERS, X, G, DUALS);
  };
  const size_t rows_cap = (size_t)m + 1, nnz_row = (size_t)n + 1;
  OaCutsFn cf = [cuts, ctx, rows_cap, nnz_row](int B, const double *X, const double *duals, CutRows &out) {
    const int cc = (int)(B * rows_cap), cn = (int)(B * rows_cap * nnz_row);
    out.row_start.assign(cc + 1, 0); out.col.resize(cn); out.val.resize(cn);
    out.lb.resize(cc); out.ub.resize(cc); out.src_row.resize(cc);
    int nc = 0;
    const int rc = cuts(ctx, B, X, duals, &nc, out.row_start.data(), out.col.data(), out.val.data(), out.lb.data(), out.ub.data(),
                        out.src_row.data(), cc, cn);
    if (rc) return rc;
    if (nc < 0 || nc > cc) return -2;
    out.row_start.resize(nc + 1); out.lb.resize(nc); out.ub.resize(nc); out.src_row.resize(nc);
    return 0;
  };
  MasterMipFn mf = [lp_solve, ctx](int nc, int nr, const int *rs, const int *cl, const double *vl, const double *rlb,
                                   const double *rub, const double *clb, const double *cub, const double *ob, double off,
                                   const int *isint, double co, double maxs, double *x, double *bound) {
    return lp_solve(ctx, nc, nr, rs, cl, vl, rlb, rub, clb, cub, ob, off, isint, co, maxs, x, bound);
  };
  OaFeasResult R = oa_feas_check(root, lp, x_lp, cutoff, sf, cf, mf, opt);
  res->status = R.status; res->cuts = R.cuts; res->nlp_solves = R.nlp_solves; res->fathomed = R.fathomed;
  res->has_incumbent = R.has_incumbent; res->has_x_lp = R.has_x_lp; res->cutoff = R.cutoff; res->incumbent_obj = R.incumbent_obj;
  res->new_nnz = (int)R.new_rows.col.size();
  if (R.status < 0) return R.status;
  if (R.cuts > cap_cuts || res->new_nnz > cap_nnz) return -2;
  if (incumbent_x && R.has_incumbent) std::memcpy(incumbent_x, R.incumbent_x.data(), sizeof(double) * n);
  if (x_lp_out && R.has_x_lp) std::memcpy(x_lp_out, R.x_lp.data(), sizeof(double) * n_cols);
  if (new_row_start) std::memcpy(new_row_start, R.new_rows.row_start.data(), sizeof(int) * (R.cuts + 1));
  if (R.cuts) {
    std::memcpy(new_col, R.new_rows.col.data(), sizeof(int) * res->new_nnz);
    std::memcpy(new_val, R.new_rows.val.data(), sizeof(double) * res->new_nnz);
    std::memcpy(new_lb, R.new_rows.lb.data(), sizeof(double) * R.cuts);
    std::memcpy(new_ub, R.new_rows.ub.data(), sizeof(double) * R.cuts);
    if (new_src) std::memcpy(new_src, R.new_rows.src_row.data(), sizeof(int) * R.cuts);
  }
  return 0;
}

int bnlp_host_selftest(void) {
  using namespace Bonmin;
#define EXPECT(c) do { if (!(c)) return __LINE__; } while (0)
  int calls = 0, warm_calls = 0, fail_next = 0;
  std::vector<double> seen_dual0, seen_x0, seen_xl, seen_xu;
  BatchSolveFn stub = [&](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT, double *X,
                          double *G, double *D) {
    ++calls;
    if (fail_next) return fail_next;
    seen_x0.assign(X0, X0 + 2); seen_xl.assign(XL, XL + 2); seen_xu.assign(XU, XU + 2);
    for (int b = 0; b < B; ++b) {
      OBJ[b] = X0[0] + XL[0]; ST[b] = 1; IT[b] = 7;
      if (X) for (int j = 0; j < 2; ++j) X[2 * b + j] = std::min(std::max(X0[j], XL[2 * b + j]), XU[2 * b + j]);
      if (G) G[b] = 0.5;
      if (D) for (int j = 0; j < 5; ++j) D[5 * b + j] = 0.25 * (j + 1);
    }
    return 0;
  };
  BatchSolveWarmFn wstub = [&](int B, const double *XL, const double *XU, const double *X0, const double *DUAL0, double *OBJ,
                               int *ST, int *IT, double *X, double *G, double *D) {
    ++warm_calls;
    seen_dual0.assign(DUAL0, DUAL0 + 5);
    const int rc = stub(B, XL, XU, X0, OBJ, ST, IT, X, G, D);
    for (int b = 0; b < B; ++b) IT[b] = 3;
    return rc;
  };
  Ipopt::SmartPtr<B200Solver> app_holder = new B200Solver(stub, wstub);
  B200Solver &app = *app_holder;
  std::istringstream opts("# comment\nbonmin.tol 1e-9\nmax_iter 123 # trailing\nprint_level 3\ntol 1e-7\n");
  EXPECT(app.Initialize(opts));
  EXPECT(app.option("tol") && *app.option("tol") == "1e-7" && *app.option("max_iter") == "123" && !app.option("mu_init"));
  EXPECT(app.logLevel() == 3);
  app.forceSolverOutput(5); EXPECT(app.logLevel() == 5);
  app.setOutputToDefault(); EXPECT(app.logLevel() == 3);
  std::istringstream bad("tol\n");
  EXPECT(!app.Initialize(bad));
  EXPECT(app.Initialize(std::string("/nonexistent/bonmin.opt")));

  // ---- a TNLP that is NOT a TMINLP2TNLP: only the Ipopt::TNLP virtuals can be used, and they are, in Ipopt's order
  Ipopt::SmartPtr<MockTNLP> mock = new MockTNLP();
  EXPECT(app.OptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(mock))) == TNLPSolver::solvedOptimal);
  EXPECT(mock->log == "ibsf" && mock->status == Ipopt::SUCCESS && mock->obj == 3.0 && app.IterationCount() == 7);
  EXPECT(mock->x.size() == 2 && mock->x[0] == 2.0 && mock->zl[0] == 0.25 && mock->zu[1] == 1.0 && mock->lam[0] == 1.25 && mock->g[0] == 0.5);
  EXPECT(seen_xl[0] == 1.0 && seen_xu[1] == 3.0 && seen_x0[1] == 1.0 && warm_calls == 0);
  // warm start enabled, but this TNLP has no multipliers to give: asked for them ('S'), then for the primal start alone ('s')
  mock->log.clear();
  app.enableWarmStart();
  EXPECT(app.ReOptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(mock))) == TNLPSolver::solvedOptimal);
  EXPECT(mock->log == "ibSsf" && warm_calls == 0 && !app.lastSolveWasWarmStarted());
  // ... and now it has: one get_starting_point call with init_z = init_lambda = true, the warm provider gets [z_L | z_U | lambda]
  mock->log.clear();
  mock->duals = {0.1, 0.2, 0.3, 0.4, 0.5};
  EXPECT(app.ReOptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(mock))) == TNLPSolver::solvedOptimal);
  EXPECT(mock->log == "ibSf" && warm_calls == 1 && app.lastSolveWasWarmStarted() && app.IterationCount() == 3);
  EXPECT(seen_dual0.size() == 5 && seen_dual0[0] == 0.1 && seen_dual0[3] == 0.4 && seen_dual0[4] == 0.5);
  app.disableWarmStart();

  // ---- TMINLP2TNLP + OsiTMINLPInterface: the sequence of InterfaceTest::testSetMethods (test/InterfaceTest.cpp:165-189)
  Ipopt::SmartPtr<TMINLP2TNLP> tp = new TMINLP2TNLP();
  TMINLP2TNLP &t = *tp;
  t.n = 2; t.m = 1; t.var_types = {0, 2}; t.x_l_ = {1.0, 0.0}; t.x_u_ = {4.0, 3.0}; t.g_l_ = {-1e19}; t.g_u_ = {1.0};
  t.x_init_ = {2.0, 1.0};
  EXPECT(app.getUsedWarmStart(tp) == nullptr);                 // no duals were set
  {
    OsiTMINLPInterface si(tp, Ipopt::SmartPtr<TNLPSolver>(GetRawPtr(app_holder)));
    const int warm_before = warm_calls;
    si.setColLower(1, 1.0);
    EXPECT(si.getColLower()[1] == 1.0);
    si.initialSolve();
    EXPECT(si.isProvenOptimal() && si.getColSolution()[1] == 1.0 && warm_calls == warm_before);   // cold solve
    CoinWarmStart *ws = si.getWarmStart();
    const IpoptWarmStart *iws = dynamic_cast<const IpoptWarmStart *>(ws);
    EXPECT(iws && app.warmStartIsValid(ws) && iws->primalSize() == 2 && iws->dualSize() == 5 && iws->primal()[1] == 1.0 && iws->dual()[4] == 1.25);
    si.setColLower(1, 0.0);
    si.setColUpper(1, 0.0);
    EXPECT(si.getColUpper()[1] == 0.0);
    si.setWarmStart(ws);
    si.resolve();                                               // setWarmStart -> enableWarmStart -> ReOptimizeTNLP
    EXPECT(si.isProvenOptimal() && si.getColSolution()[1] == 0.0);
    EXPECT(warm_calls == warm_before + 1 && app.lastSolveWasWarmStarted());
    EXPECT(seen_x0[0] == 2.0 && seen_x0[1] == 1.0 && seen_dual0[4] == 1.25 && seen_xu[1] == 0.0);   // the saved point and multipliers
    CoinWarmStart *used = app.getUsedWarmStart(tp);
    const IpoptWarmStart *iu = dynamic_cast<const IpoptWarmStart *>(used);
    EXPECT(iu && iu->primal()[1] == 1.0 && iu->dualSize() == 5);
    si.setColUpper(1, 3.0);                                     // (the model's own bound again)
    delete ws; delete used;
  }
  CoinWarmStart *empty = app.getEmptyWarmStart();
  EXPECT(!app.warmStartIsValid(empty) && app.setWarmStart(empty, tp) && !app.warmStartEnabled());
  EXPECT(!app.setWarmStart(nullptr, tp));
  delete empty;
  const int calls_before = calls;
  fail_next = -17;
  EXPECT(app.ReOptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(tp))) == TNLPSolver::exception && app.errorCode() == -17 && app.providerErrors() == 1);
  fail_next = 0;
  TNLPSolver::ReturnStatus st_exc = TNLPSolver::exception, st_ce = TNLPSolver::computationError;
  EXPECT(app.isError(st_exc) && app.isRecoverable(st_exc) && !app.isRecoverable(st_ce));
  TNLPSolver::UnsolvedError *e = app.newUnsolvedError(app.errorCode(), tp, "node17");
  std::ostringstream os;
  e->printError(os);
  EXPECT(e->errorNum() == -17 && GetRawPtr(e->model()) == GetRawPtr(tp) && e->name() == "node17" &&
         os.str().find("B200 batched IPM exited with error code -17") == 0);
  delete e;
  Ipopt::SmartPtr<TNLPSolver> c = app.clone();
  EXPECT(IsValid(c) && c->solverName() == app.solverName() && GetRawPtr(c) != static_cast<TNLPSolver *>(&app));
  EXPECT(calls == calls_before + 1);
  {   // lineariser: CSR rows of the provider -> OsiRowCuts (skipped slots dropped, global flag), LP -> OsiSolverInterface
    OaBatchFn oa = [](int B, const double *X, const double *, const double *, int getObj, double tiny, double, double, double, OaRows &o) {
      o.rows_per_point = 2;
      for (int b = 0; b < B; ++b) {
        o.row_start.push_back((int)o.col.size());
        o.col.push_back(0); o.val.push_back(X[2 * b] + tiny); o.col.push_back(1); o.val.push_back(-1.0);
        o.src_row.push_back(0); o.lb.push_back(-1e300); o.ub.push_back(10.0 + b);
        o.row_start.push_back((int)o.col.size());
        if (getObj) { o.col.push_back(2); o.val.push_back(-1.0); o.src_row.push_back(1); } else o.src_row.push_back(-1);
        o.lb.push_back(-1e300); o.ub.push_back(0.0);
      }
      o.row_start.push_back((int)o.col.size());
      return 0;
    };
    ExtractFn ex = [](const double *, const double *xl, const double *xu, int getObj, double, double, double, double, LpRelaxation &lp) {
      lp.n_cols = 2 + getObj; lp.has_alpha = getObj; lp.obj_offset = 1.5;
      lp.row_start = {0, 2}; lp.col = {0, 1}; lp.val = {1.0, 2.0}; lp.row_lb = {-1e300}; lp.row_ub = {4.0};
      lp.col_lb = {xl[0], xl[1]}; lp.col_ub = {xu[0], xu[1]}; lp.obj = {1.0, 0.0}; lp.is_integer = {0, 1};
      if (getObj) { lp.col_lb.push_back(-1e300); lp.col_ub.push_back(1e300); lp.obj = {0.0, 0.0, 1.0}; lp.is_integer.push_back(0); }
      return 0;
    };
    B200Linearizer lin(oa, ex);
    lin.set_tols(1e-8, 1e-17, 1e-8, 1e19);
    lin.set_model(&t);
    OsiCuts cs;
    const double X2[4] = {1.0, 2.0, 3.0, 4.0};
    lin.get_oas(cs, X2, false, true);
    EXPECT(cs.sizeRowCuts() == 1 && cs.rowCut(0).globallyValid() && cs.rowCut(0).ub() == 10.0 && cs.rowCut(0).indices().size() == 2);
    EXPECT(cs.rowCut(0).elements()[0] == 1.0 + 1e-8);
    lin.get_oas_batch(cs, X2, 2, true, false);
    EXPECT(cs.sizeRowCuts() == 5 && !cs.rowCut(1).globallyValid() && cs.rowCut(3).ub() == 11.0 && cs.rowCut(4).indices()[0] == 2);
    lin.get_refined_oa(cs);
    EXPECT(cs.sizeRowCuts() == 5);
    lin.set_refinement_points({1.0, 2.0, 3.0, 4.0});
    lin.get_refined_oa(cs);
    EXPECT(cs.sizeRowCuts() == 7 && cs.rowCut(6).globallyValid());
    OsiSolverInterface si;
    lin.extract(&si, X2, true);
    EXPECT(si.getNumCols() == 3 && si.getNumRows() == 1 && si.integer_[1] == 1 && si.integer_[0] == 0 && si.obj_offset_ == 1.5 && si.colub_[1] == 3.0);
    TMINLP2OsiLP *c2 = lin.clone();
    OsiCuts cs2;
    c2->get_oas(cs2, X2, true, false);
    EXPECT(cs2.sizeRowCuts() == 2);
    delete c2;
  }
  {  // B-OA loop (B200Oa.cpp: oa_run) on closed-form providers, no Python and no engine: min (x - 1.4)^2, x integer in [0, 3].
     // Objective cuts alpha >= f(xk) + f'(xk)(x - xk); the master is solved by enumeration.  Expected: x = 1, f = 0.16.
    Ipopt::SmartPtr<TMINLP2TNLP> toy_holder = new TMINLP2TNLP();
    TMINLP2TNLP &toy = *toy_holder;
    toy.n = 1; toy.m = 0;
    toy.var_types = {2}; toy.x_l_ = {0.0}; toy.x_u_ = {3.0}; toy.x_init_ = {3.0};
    auto f = [](double x) { return (x - 1.4) * (x - 1.4); };
    BatchSolveFn nlp = [&](int B, const double *XL, const double *XU, const double *, double *OBJ, int *ST, int *IT, double *X, double *,
                           double *D) {
      for (int b = 0; b < B; ++b) {
        const double x = std::min(std::max(1.4, XL[b]), XU[b]);
        OBJ[b] = f(x); ST[b] = 1; IT[b] = 1;
        if (X) X[b] = x;
        if (D) D[2 * b] = D[2 * b + 1] = 0.0;
      }
      return 0;
    };
    auto cut_at = [&](double xk, CutRows &o) {          // 2(xk-1.4) x - alpha <= 2(xk-1.4) xk - f(xk)
      const double s = 2 * (xk - 1.4);
      o.col.push_back(0); o.val.push_back(s); o.col.push_back(1); o.val.push_back(-1.0);
      o.row_start.push_back((int)o.col.size()); o.lb.push_back(-1.7976931348623157e308); o.ub.push_back(s * xk - f(xk));
      o.src_row.push_back(0);
    };
    OaCutsFn cuts = [&](int B, const double *X, const double *, CutRows &o) { for (int b = 0; b < B; ++b) cut_at(X[b], o); return 0; };
    OaExtractFn ext = [&](const double *x, const double *, LpRelaxation &lp) {
      CutRows o; cut_at(x[0], o);
      lp.n_cols = 2; lp.has_alpha = 1; lp.obj_offset = 0; lp.row_start = o.row_start; lp.col = o.col; lp.val = o.val;
      lp.row_lb = o.lb; lp.row_ub = o.ub; lp.obj = {0.0, 1.0};
      return 0;
    };
    int master_calls = 0;
    MasterMipFn master = [&](int nc, int nr, const int *rs, const int *cl, const double *vl, const double *, const double *rub,
                             const double *clb, const double *cub, const double *, double, const int *isint, double cutoff, double,
                             double *x, double *bound) {
      ++master_calls;
      if (nc != 2 || !isint[0] || isint[1]) return -1;
      double best = 1e300; int arg = -1;
      for (int xi = (int)clb[0]; xi <= (int)cub[0]; ++xi) {
        double alpha = -1e300;
        for (int r = 0; r < nr; ++r) {
          double a = 0;
          for (int e = rs[r]; e < rs[r + 1]; ++e) if (cl[e] == 0) a = vl[e];
          alpha = std::max(alpha, a * xi - rub[r]);
        }
        if (alpha < cutoff && alpha < best) { best = alpha; arg = xi; }
      }
      if (arg < 0) return 1;
      x[0] = arg; x[1] = best; *bound = best;
      return 0;
    };
    OaResult R = oa_run(toy, nlp, cuts, ext, master, OaOptions());
    EXPECT(R.status == 0 && R.best_x.size() == 1 && R.best_x[0] == 1.0);
    EXPECT(std::fabs(R.best_obj - 0.16) < 1e-12 && R.nlp_solves == R.passes + 1 && master_calls == R.passes + 1);
    EXPECT(R.cuts == 2 * R.passes && R.oa_points == 2 * R.passes && R.lower_bound <= R.best_obj + 1e-12);
    OaOptions one; one.max_passes = 1;
    R = oa_run(toy, nlp, cuts, ext, master, one);
    EXPECT(R.status == 2 && R.passes == 1);
  }
#undef EXPECT
  return 0;
}

int bnlp_resolve_cb(int mode, int n, int m, const double *xl, const double *xu, const double *x_init, double cur_obj,
                    int cur_status, int numsolve, double max_random_point_radius, unsigned seed,
                    bnlp_batch_solve_cb cb, void *ctx, bnlp_retry_result *res, double *obj, int *status, double *x,
                    double *g, double *duals) {
  using namespace Bonmin;
  Ipopt::SmartPtr<TMINLP2TNLP> t_holder = new TMINLP2TNLP();   // SmartPtr-managed objects live on the heap
  TMINLP2TNLP &t = *t_holder;
  t.n = n; t.m = m;
  t.x_l_.assign(xl, xl + n); t.x_u_.assign(xu, xu + n);
  t.x_init_.assign(x_init, x_init + n);
  t.x_sol_.assign(n, 0.0); t.g_sol_.assign(m, 0.0); t.duals_sol_.assign(2 * n + m, 0.0);
  t.obj_value_ = cur_obj;
  t.return_status = (TNLPSolver::ReturnStatus)cur_status;
  RetryOptions o;
  if (max_random_point_radius > 0) o.max_random_point_radius = max_random_point_radius;
  o.seed = seed;
  BatchSolveStartsFn fn = [cb, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT,
                                    double *X, double *G, double *D) { return cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D); };
  RetryResult r;
  const TNLPSolver::ReturnStatus st = mode == 0 ? resolve_for_robustness(t, fn, numsolve, o, r) : resolve_for_cost(t, fn, numsolve, o, r);
  if (res) { res->used = r.used; res->num_failed = r.num_failed; res->num_infeasible = r.num_infeasible;
             res->nlp_solves = r.nlp_solves; res->mean = r.mean; res->std_dev = r.std_dev; }
  if (obj) *obj = t.obj_value_;
  if (status) *status = (int)st;
  if (x) memcpy(x, t.x_sol_.data(), sizeof(double) * n);
  if (g && m) memcpy(g, t.g_sol_.data(), sizeof(double) * m);
  if (duals) memcpy(duals, t.duals_sol_.data(), sizeof(double) * (2 * n + m));
  return 0;
}

int bnlp_dive_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                     const double *gu, const double *x_sol, const double *g_sol, double obj_sol, const int *column_length,
                     double cutoff, bnlp_batch_solve_cb cb, bnlp_grad_f_cb grad_cb, void *ctx, const bnlp_dive_options *o,
                     bnlp_dive_result *res, double *better_x) {
  using namespace Bonmin;
  Ipopt::SmartPtr<TMINLP2TNLP> node_holder = new TMINLP2TNLP();   // SmartPtr-managed objects live on the heap
  TMINLP2TNLP &node = *node_holder;
  node.n = n; node.m = m;
  node.var_types.assign(var_type, var_type + n);
  node.x_l_.assign(xl, xl + n); node.x_u_.assign(xu, xu + n);
  node.g_l_.assign(gl, gl + m); node.g_u_.assign(gu, gu + m);
  node.x_init_.assign(x_sol, x_sol + n);
  node.x_sol_.assign(x_sol, x_sol + n);
  node.g_sol_.assign(g_sol, g_sol + m);
  node.obj_value_ = obj_sol;
  DiveOptions opt;
  if (o) {
    opt.rule = o->rule; opt.speculate = o->speculate != 0; opt.seed = o->seed;
    if (o->integer_tolerance > 0) opt.integer_tolerance = o->integer_tolerance;
    if (o->percentage_to_fix > 0) opt.percentage_to_fix = o->percentage_to_fix;
  }
  BatchSolveFn fn = [cb, ctx](int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT,
                              double *X, double *G, double *D) { return cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D); };
  GradFFn gfn;
  if (grad_cb) gfn = [grad_cb, ctx](const double *x, double *g) { return grad_cb(ctx, x, g); };
  std::vector<int> len;
  if (column_length) len.assign(column_length, column_length + n);
  DiveResult r = dive_run(node, fn, gfn, len, cutoff, opt);
  res->found = r.found ? 1 : 0; res->steps = r.steps; res->nlp_solves = r.nlp_solves; res->batches = r.batches;
  res->ipm_iterations = r.ipm_iterations; res->obj = r.obj;
  if (better_x && r.found) memcpy(better_x, r.x.data(), sizeof(double) * n);
  return 0;
}

#ifdef BNLP_WITH_ENGINE
int bnlp_resolve_engine(bnlp_engine *E, int mode, const double *xl, const double *xu, const double *x_init, double cur_obj,
                        int cur_status, int numsolve, double max_random_point_radius, unsigned seed, bnlp_retry_result *res,
                        double *obj, int *status, double *x, double *g, double *duals) {
  if (!E) return -1;
  // n, m from the engine's own evaluation entry point would need the problem; the starts callback gets them through ctx
  struct Ctx { bnlp_engine *E; } ctx{E};
  auto cb = [](void *c, int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT, double *X,
               double *G, double *D) {
    return bnlp_solve_batch(((Ctx *)c)->E, B, XL, XU, X0, 1 /* per-child start points */, OBJ, ST, IT, X, G, D);
  };
  int n = 0, m = 0;
  if (bnlp_engine_sizes(E, &n, &m)) return -1;
  return bnlp_resolve_cb(mode, n, m, xl, xu, x_init, cur_obj, cur_status, numsolve, max_random_point_radius, seed, cb, &ctx, res,
                         obj, status, x, g, duals);
}

int bnlp_dive_run_engine(bnlp_engine *E, const bnlp_problem *P, const double *xl, const double *xu, const double *x_sol,
                         const double *g_sol, double obj_sol, double cutoff, const bnlp_dive_options *o,
                         bnlp_dive_result *res, double *better_x) {
  int n = 0, m = 0, nj = 0;
  if (!E || !P || bnlp_problem_sizes(P, &n, &m, &nj, nullptr)) return -1;
  std::vector<int> vt(n), jc(nj), len(n, 0);
  std::vector<double> gl(m), gu(m);
  bnlp_problem_info(P, nullptr, jc.data(), nullptr, nullptr, vt.data(), nullptr, nullptr, nullptr, gl.data(), gu.data(), nullptr);
  for (int e = 0; e < nj; ++e) len[jc[e]]++;   // HeuristicDiveVectorLength::setInternalVariables (:50-91): rows per column
  auto cb = [](void *ctx, int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT,
               double *X, double *G, double *D) {
    return bnlp_solve_batch((bnlp_engine *)ctx, B, XL, XU, X0, 0, OBJ, ST, IT, X, G, D);
  };
  auto gcb = [](void *ctx, const double *x, double *g) {
    return bnlp_eval_batch((bnlp_engine *)ctx, 1, x, nullptr, 1.0, nullptr, nullptr, g, nullptr, nullptr);
  };
  return bnlp_dive_run_cb(n, m, vt.data(), xl, xu, gl.data(), gu.data(), x_sol, g_sol, obj_sol, len.data(), cutoff, cb, gcb, E,
                          o, res, better_x);
}

// The lineariser with the CUDA engine behind it: B x get_oas and one extract, flattened for a C caller (tests, bench).
int bnlp_linearizer_engine(bnlp_engine *E, const bnlp_problem *P, int B, const double *X, int getObj, int global, int *n_cuts,
                           int *row_start, int *col, double *val, double *lb, double *ub, int *is_global, int cap_cuts,
                           int cap_nnz, int *lp_cols, int *lp_rows, int *lp_nnz, int *lp_integers) {
  using namespace Bonmin;
  int n = 0, m = 0;
  if (!E || !P || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  Ipopt::SmartPtr<TMINLP2TNLP> model_holder = new TMINLP2TNLP();   // SmartPtr-managed objects live on the heap
  TMINLP2TNLP &model = *model_holder;
  model.n = n; model.m = m;
  model.x_l_.resize(n); model.x_u_.resize(n);
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, model.x_l_.data(), model.x_u_.data(), nullptr, nullptr, nullptr);
  OaBatchFn oa = [E](int Bp, const double *Xp, const double *xl, const double *xu, int go, double tiny, double vtiny, double relax,
                     double infty, OaRows &o) {
    size_t cap = 0;
    if (bnlp_oa_capacity(E, Bp, &o.rows_per_point, &cap)) return -1;
    const size_t slots = (size_t)Bp * o.rows_per_point;
    o.row_start.resize(slots + 1); o.col.resize(cap); o.val.resize(cap); o.lb.resize(slots); o.ub.resize(slots); o.src_row.resize(slots);
    bnlp_oa_tols tl{tiny, vtiny, relax, infty, 0};
    return bnlp_oa_batch(E, Bp, Xp, nullptr, xl, xu, nullptr, 0.0, go, &tl, o.row_start.data(), o.col.data(), o.val.data(),
                         o.lb.data(), o.ub.data(), o.src_row.data());
  };
  const int mm = m, nn = n;
  ExtractFn ex = [E, mm, nn](const double *x, const double *xl, const double *xu, int go, double tiny, double vtiny, double relax,
                             double infty, LpRelaxation &lp) {
    bnlp_oa_tols tl{tiny, vtiny, relax, infty, 0};
    bnlp_lp_dims d{};
    const size_t rows = (size_t)mm + 1, nnz = rows * ((size_t)nn + 1);
    lp.row_start.resize(rows + 1); lp.col.resize(nnz); lp.val.resize(nnz); lp.row_lb.resize(rows); lp.row_ub.resize(rows);
    std::vector<int> src(rows);
    lp.col_lb.resize(nn + 1); lp.col_ub.resize(nn + 1); lp.obj.resize(nn + 1); lp.is_integer.resize(nn + 1);
    const int rc = bnlp_extract_linear_relaxation(E, x, nullptr, xl, xu, go, &tl, &d, lp.row_start.data(), lp.col.data(), lp.val.data(),
                                                  lp.row_lb.data(), lp.row_ub.data(), src.data(), lp.col_lb.data(), lp.col_ub.data(),
                                                  lp.obj.data(), lp.is_integer.data());
    if (rc) return rc;
    lp.n_cols = d.n_cols; lp.has_alpha = d.has_alpha; lp.obj_offset = d.obj_offset;
    lp.row_start.resize(d.n_rows + 1); lp.row_lb.resize(d.n_rows); lp.row_ub.resize(d.n_rows);
    lp.col_lb.resize(d.n_cols); lp.col_ub.resize(d.n_cols); lp.obj.resize(d.n_cols); lp.is_integer.resize(d.n_cols);
    return 0;
  };
  B200Linearizer lin(oa, ex);
  lin.set_tols(1e-8, 1e-17, 1e-8, 1e19);   // tiny_element, very_tiny_element, oa_rhs_relax, infinity (BonOsiTMINLPInterface.cpp:222-247)
  lin.set_model(&model);
  OsiCuts cs;
  lin.get_oas_batch(cs, X, B, getObj != 0, global != 0);
  if (lin.lastError()) return lin.lastError();
  int nnz = 0;
  if (cs.sizeRowCuts() > cap_cuts) return -2;
  for (int k = 0; k < cs.sizeRowCuts(); ++k) {
    const OsiRowCut &c = cs.rowCut(k);
    if (nnz + (int)c.indices().size() > cap_nnz) return -2;
    row_start[k] = nnz;
    for (size_t e = 0; e < c.indices().size(); ++e) { col[nnz] = c.indices()[e]; val[nnz] = c.elements()[e]; ++nnz; }
    lb[k] = c.lb(); ub[k] = c.ub(); is_global[k] = c.globallyValid();
  }
  row_start[cs.sizeRowCuts()] = nnz;
  *n_cuts = cs.sizeRowCuts();
  OsiSolverInterface si;
  lin.extract(&si, X, getObj != 0);
  if (lin.lastError()) return lin.lastError();
  if (lp_cols) *lp_cols = si.getNumCols();
  if (lp_rows) *lp_rows = si.getNumRows();
  if (lp_nnz) *lp_nnz = si.start_.empty() ? 0 : si.start_.back();
  if (lp_integers) { *lp_integers = 0; for (int v : si.integer_) *lp_integers += v; }
  return 0;
}

// Product path of B-OA: the engine is the arithmetic provider of every step, only the master MILP is the caller's.
namespace {
struct OaEngineCtx { bnlp_engine *E; int n, m; std::vector<double> xl, xu; bnlp_master_mip_cb master; void *mctx; };
int oa_eng_solve(void *p, int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *STATUS, int *ITERS,
                 double *X, double *G, double *DUALS) {
  return bnlp_solve_batch(((OaEngineCtx *)p)->E, B, XL, XU, X0, 0, OBJ, STATUS, ITERS, X, G, DUALS);
}
int oa_eng_cuts(void *p, int B, const double *X, const double *duals, int *n_cuts, int *row_start, int *col, double *val,
                double *lb, double *ub, int *src_row, int cap_cuts, int cap_nnz) {
  OaEngineCtx &c = *(OaEngineCtx *)p;
  int rpp = 0;
  size_t cap = 0;
  if (bnlp_oa_capacity(c.E, B, &rpp, &cap)) return -1;
  const size_t slots = (size_t)B * rpp;
  std::vector<int> rs(slots + 1), ci(cap), src(slots);
  std::vector<double> cv(cap), LB(slots), UB(slots);
  bnlp_oa_tols tl{1e-8, 1e-17, 1e-8, 1e19, 0};   // tiny_element, very_tiny_element, oa_rhs_relax, infinity
  // bnlp_oa_batch takes one multiplier vector PER POINT; the loop has one for all of them (the interface's getRowPrice())
  const size_t nmul = 2 * (size_t)c.n + c.m;
  std::vector<double> du((size_t)B * nmul);
  for (int b = 0; b < B; ++b) std::copy(duals, duals + nmul, du.begin() + (size_t)b * nmul);
  if (int rc = bnlp_oa_batch(c.E, B, X, du.data(), c.xl.data(), c.xu.data(), nullptr, 0.0, 1, &tl, rs.data(), ci.data(), cv.data(),
                             LB.data(), UB.data(), src.data()))
    return rc;
  int nc = 0, nnz = 0;
  row_start[0] = 0;
  for (size_t k = 0; k < slots; ++k) {
    if (src[k] < 0) continue;                      // row not generated
    if (nc >= cap_cuts || nnz + (rs[k + 1] - rs[k]) > cap_nnz) return -2;
    for (int e = rs[k]; e < rs[k + 1]; ++e) { col[nnz] = ci[e]; val[nnz] = cv[e]; ++nnz; }
    lb[nc] = LB[k]; ub[nc] = UB[k]; src_row[nc] = src[k];
    row_start[++nc] = nnz;
  }
  *n_cuts = nc;
  return 0;
}
int oa_eng_extract(void *p, const double *x, const double *duals, int *n_rows, int *has_alpha, double *obj_offset, int *row_start,
                   int *col, double *val, double *row_lb, double *row_ub, double *obj, int cap_rows, int cap_nnz) {
  OaEngineCtx &c = *(OaEngineCtx *)p;
  if (cap_rows < c.m + 1 || (size_t)cap_nnz < ((size_t)c.m + 1) * ((size_t)c.n + 1)) return -2;
  bnlp_oa_tols tl{1e-8, 1e-17, 1e-8, 1e19, 0};
  bnlp_lp_dims d{};
  std::vector<int> src(c.m + 1), isint(c.n + 1);
  std::vector<double> clb(c.n + 1), cub(c.n + 1);
  if (int rc = bnlp_extract_linear_relaxation(c.E, x, duals, nullptr, nullptr, 1, &tl, &d, row_start, col, val, row_lb, row_ub,
                                              src.data(), clb.data(), cub.data(), obj, isint.data()))
    return rc;
  *n_rows = d.n_rows; *has_alpha = d.has_alpha; *obj_offset = d.obj_offset;
  return 0;
}
int oa_eng_master(void *p, int nc, int nr, const int *rs, const int *cl, const double *vl, const double *rlb, const double *rub,
                  const double *clb, const double *cub, const double *ob, double off, const int *isint, double cutoff, double maxs,
                  double *x, double *bound) {
  OaEngineCtx *c = (OaEngineCtx *)p;
  return c->master(c->mctx, nc, nr, rs, cl, vl, rlb, rub, clb, cub, ob, off, isint, cutoff, maxs, x, bound);
}
}  // namespace
int bnlp_oa_run_engine(bnlp_engine *E, const bnlp_problem *P, bnlp_master_mip_cb master, void *ctx, const bnlp_oa_run_options *o,
                       bnlp_oa_run_result *res, double *best_x) {
  int n = 0, m = 0;
  if (!E || !P || !master || !res || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  std::vector<int> vt(n);
  std::vector<double> x0(n);
  OaEngineCtx C{E, n, m, std::vector<double>(n), std::vector<double>(n), master, ctx};
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, vt.data(), nullptr, C.xl.data(), C.xu.data(), nullptr, nullptr, x0.data());
  return bnlp_oa_run_cb(n, m, vt.data(), C.xl.data(), C.xu.data(), x0.data(), oa_eng_solve, oa_eng_cuts, oa_eng_extract,
                        oa_eng_master, &C, o, res, best_x);
}
int bnlp_oa_nlp_optim_engine(bnlp_engine *E, const bnlp_problem *P, const double *node_lb, const double *node_ub, const double *x_lp,
                             bnlp_nlp_optim_result *res, double *x_nlp, double *duals_out, int *new_row_start, int *new_col,
                             double *new_val, double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz) {
  int n = 0, m = 0;
  if (!E || !P || !res || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  std::vector<int> vt(n);
  OaEngineCtx C{E, n, m, std::vector<double>(n), std::vector<double>(n), nullptr, nullptr};
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, vt.data(), nullptr, C.xl.data(), C.xu.data(), nullptr, nullptr, nullptr);
  return bnlp_oa_nlp_optim_cb(n, m, vt.data(), C.xl.data(), C.xu.data(), node_lb, node_ub, x_lp, oa_eng_solve, oa_eng_cuts, &C, res,
                              x_nlp, duals_out, new_row_start, new_col, new_val, new_lb, new_ub, new_src, cap_cuts, cap_nnz);
}
int bnlp_ecp_rounds_engine(bnlp_engine *E, const bnlp_problem *P, int n_cols, int n_rows, const int *row_start, const int *col,
                           const double *val, const double *row_lb, const double *row_ub, const double *col_lb,
                           const double *col_ub, const double *obj, double obj_offset, const int *is_integer, const double *duals,
                           int rounds, double abs_tol, double rel_tol, bnlp_master_mip_cb lp_solve, void *ctx,
                           bnlp_ecp_result *res, double *x_out, int *new_row_start, int *new_col, double *new_val, double *new_lb,
                           double *new_ub, int *new_src, int cap_cuts, int cap_nnz) {
  int n = 0, m = 0;
  if (!E || !P || !lp_solve || !res || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  OaEngineCtx C{E, n, m, std::vector<double>(n), std::vector<double>(n), lp_solve, ctx};
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, C.xl.data(), C.xu.data(), nullptr, nullptr, nullptr);
  auto viol = [](void *p, const double *x, double o, double *v) {    // getNonLinearitiesViolation(x, obj)
    return bnlp_violation_batch(((OaEngineCtx *)p)->E, 1, x, &o, v, nullptr);
  };
  return bnlp_ecp_rounds_cb(n, m, n_cols, n_rows, row_start, col, val, row_lb, row_ub, col_lb, col_ub, obj, obj_offset, is_integer,
                            duals, rounds, abs_tol, rel_tol, oa_eng_cuts, viol, oa_eng_master, &C, res, x_out, new_row_start, new_col,
                            new_val, new_lb, new_ub, new_src, cap_cuts, cap_nnz);
}
int bnlp_oa_feas_check_engine(bnlp_engine *E, const bnlp_problem *P, int n_cols, int n_rows, const int *row_start, const int *col,
                              const double *val, const double *row_lb, const double *row_ub, const double *col_lb,
                              const double *col_ub, const double *obj, double obj_offset, const double *x_lp, double cutoff,
                              const bnlp_oa_feas_options *o, bnlp_master_mip_cb lp_solve, void *ctx, bnlp_oa_feas_result *res,
                              double *incumbent_x, double *x_lp_out, int *new_row_start, int *new_col, double *new_val,
                              double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz) {
  int n = 0, m = 0;
  if (!E || !P || !lp_solve || !res || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  std::vector<int> vt(n);
  OaEngineCtx C{E, n, m, std::vector<double>(n), std::vector<double>(n), lp_solve, ctx};
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, vt.data(), nullptr, C.xl.data(), C.xu.data(), nullptr, nullptr, nullptr);
  return bnlp_oa_feas_check_cb(n, m, vt.data(), C.xl.data(), C.xu.data(), n_cols, n_rows, row_start, col, val, row_lb, row_ub, col_lb,
                               col_ub, obj, obj_offset, x_lp, cutoff, o, oa_eng_solve, oa_eng_cuts, oa_eng_master, &C, res,
                               incumbent_x, x_lp_out, new_row_start, new_col, new_val, new_lb, new_ub, new_src, cap_cuts, cap_nnz);
}

// Product path: the same driver with the CUDA engine as arithmetic provider (no Python in between).
int bnlp_bb_run_engine(bnlp_engine *E, const bnlp_problem *P, const bnlp_bb_options *o, bnlp_bb_result *res,
                       double *best_x) {
  int n = 0, m = 0;
  if (!E || !P || bnlp_problem_sizes(P, &n, &m, nullptr, nullptr)) return -1;
  std::vector<int> vt(n);
  std::vector<double> xl(n), xu(n), gl(m), gu(m), x0(n);
  bnlp_problem_info(P, nullptr, nullptr, nullptr, nullptr, vt.data(), nullptr, xl.data(), xu.data(), gl.data(), gu.data(),
                    x0.data());
  auto cb = [](void *ctx, int B, const double *XL, const double *XU, const double *X0, double *OBJ, int *ST, int *IT,
               double *X, double *G, double *D) {
    return bnlp_solve_batch((bnlp_engine *)ctx, B, XL, XU, X0, 0, OBJ, ST, IT, X, G, D);
  };
  auto qcb = [](void *ctx, const double *xs, const double *du, int B, const double *XL, const double *XU, double *OBJ,
                int *ST, int *IT, double *X) { return bnlp_tqp_solve_batch((bnlp_engine *)ctx, xs, du, B, XL, XU, OBJ, ST, IT, X); };
  auto ccb = [](void *ctx, int P, const double *PXL, const double *PXU, int B, const int *cp, const int *cc, const int *cw,
                const double *cb2, const double *X0, double *OBJ, int *ST, int *IT, double *X) {
    return bnlp_solve_children((bnlp_engine *)ctx, P, PXL, PXU, B, cp, cc, cw, cb2, X0, 0, OBJ, ST, IT, X, nullptr, nullptr);
  };
  bnlp_bb_options o2{};
  if (o) o2 = *o;
  std::vector<int> pri(n);
  if (!o2.priorities && (bnlp_problem_branching_info(P, pri.data(), nullptr, nullptr, nullptr) & 1)) o2.priorities = pri.data();
  return bnlp_bb_run_cb3(n, m, vt.data(), xl.data(), xu.data(), gl.data(), gu.data(), x0.data(), cb, qcb, ccb, E, &o2, res, best_x);
}
#endif

}  // extern "C"
