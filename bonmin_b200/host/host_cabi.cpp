// C entry points of the host layer (used by bench.py, tests and the `bnlp_bb` command): the B-BB driver with a
// batch-solve callback.  In the product the callback is bnlp_solve_batch on a CUDA engine (bnlp_bb_run_engine in
// bnlp_api.cu); tests may pass the CPU oracle to check that both give the same MINLP optimum.
#include <cstring>
#include <sstream>

#include "../../include/bnlp.h"
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
  TMINLP2TNLP root;
  root.n = n; root.m = m;
  root.var_types.assign(var_type, var_type + n);
  root.x_l.assign(xl, xl + n); root.x_u.assign(xu, xu + n);
  root.g_l.assign(gl, gl + m); root.g_u.assign(gu, gu + m);
  root.x_init.assign(x0, x0 + n);
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
  res->sb_children = r.sb_children; res->served_from_sb_cache = r.served_from_sb_cache;
  res->ipm_iterations = r.ipm_iterations; res->best_obj = r.best_obj; res->best_bound = r.best_bound; res->seconds = r.seconds;
  if (best_x && !r.best_x.empty()) memcpy(best_x, r.best_x.data(), sizeof(double) * n);
  return r.status == 3 ? -1 : 0;   // the arithmetic provider failed: an error, not a search result
}

// Exercises the parts of the TNLPSolver contract that have no other C entry point (options, warm starts, error objects,
// output levels) on a stub provider; returns 0, or the source line of the first expectation that failed.
int bnlp_host_selftest(void) {
  using namespace Bonmin;
#define EXPECT(c) do { if (!(c)) return __LINE__; } while (0)
  int calls = 0, fail_next = 0;
  BatchSolveFn stub = [&](int B, const double *XL, const double *, const double *X0, double *OBJ, int *ST, int *IT, double *X,
                          double *G, double *D) {
    ++calls;
    if (fail_next) return fail_next;
    for (int b = 0; b < B; ++b) {
      OBJ[b] = X0[0] + XL[0]; ST[b] = 1; IT[b] = 7;
      if (X) for (int j = 0; j < 2; ++j) X[2 * b + j] = X0[j];
      if (G) G[b] = 0.5;
      if (D) for (int j = 0; j < 5; ++j) D[5 * b + j] = 0.25 * (j + 1);
    }
    return 0;
  };
  B200Solver app(stub);
  std::istringstream opts("# comment\nbonmin.tol 1e-9\nmax_iter 123 # trailing\nprint_level 3\ntol 1e-7\n");
  EXPECT(app.Initialize(opts));
  EXPECT(app.option("tol") && *app.option("tol") == "1e-7" && *app.option("max_iter") == "123" && !app.option("mu_init"));
  EXPECT(app.logLevel() == 3);
  app.forceSolverOutput(5); EXPECT(app.logLevel() == 5);
  app.setOutputToDefault(); EXPECT(app.logLevel() == 3);
  std::istringstream bad("tol\n");
  EXPECT(!app.Initialize(bad));
  EXPECT(app.Initialize(std::string("/nonexistent/bonmin.opt")));
  TMINLP2TNLP t;
  t.n = 2; t.m = 1; t.var_types = {0, 2}; t.x_l = {1.0, 0.0}; t.x_u = {4.0, 3.0}; t.g_l = {-1e19}; t.g_u = {1.0};
  t.x_init = {2.0, 1.0};
  EXPECT(app.getUsedWarmStart(t) == nullptr);                 // no duals were set
  EXPECT(app.OptimizeTNLP(t) == TNLPSolver::solvedOptimal && t.obj_value == 3.0 && app.IterationCount() == 7);
  CoinWarmStart *ws = app.getWarmStart(t);
  EXPECT(ws && app.warmStartIsValid(ws) && ws->primalSize() == 2 && ws->dualSize() == 5 && ws->primal()[0] == 2.0 && ws->dual()[4] == 1.25);
  t.x_init = {0.0, 0.0};
  EXPECT(!app.warmStartEnabled() && app.setWarmStart(ws, t) && app.warmStartEnabled());
  EXPECT(t.x_init[0] == 2.0 && t.duals_init.size() == 5);
  CoinWarmStart *used = app.getUsedWarmStart(t);
  EXPECT(used && used->primal()[1] == 1.0 && used->dualSize() == 5);
  CoinWarmStart *empty = app.getEmptyWarmStart();
  EXPECT(!app.warmStartIsValid(empty) && app.setWarmStart(empty, t) && !app.warmStartEnabled());
  EXPECT(!app.setWarmStart(nullptr, t));
  delete ws; delete used; delete empty;
  fail_next = -17;
  EXPECT(app.ReOptimizeTNLP(t) == TNLPSolver::exception && app.errorCode() == -17 && app.providerErrors() == 1);
  EXPECT(TNLPSolver::isError(t.return_status) && TNLPSolver::isRecoverable(TNLPSolver::exception) &&
         !TNLPSolver::isRecoverable(TNLPSolver::computationError));
  TNLPSolver::UnsolvedError *e = app.newUnsolvedError(app.errorCode(), &t, "node17");
  std::ostringstream os;
  e->printError(os);
  EXPECT(e->errorNum() == -17 && e->model() == &t && e->name() == "node17" && os.str().find("B200 batched IPM exited with error code -17") == 0);
  delete e;
  TNLPSolver *c = app.clone();
  EXPECT(c && c->solverName() == app.solverName());
  delete c;
  EXPECT(calls == 2);
#undef EXPECT
  return 0;
}

int bnlp_resolve_cb(int mode, int n, int m, const double *xl, const double *xu, const double *x_init, double cur_obj,
                    int cur_status, int numsolve, double max_random_point_radius, unsigned seed,
                    bnlp_batch_solve_cb cb, void *ctx, bnlp_retry_result *res, double *obj, int *status, double *x,
                    double *g, double *duals) {
  using namespace Bonmin;
  TMINLP2TNLP t;
  t.n = n; t.m = m;
  t.x_l.assign(xl, xl + n); t.x_u.assign(xu, xu + n);
  t.x_init.assign(x_init, x_init + n);
  t.x_sol.assign(n, 0.0); t.g_sol.assign(m, 0.0); t.duals_sol.assign(2 * n + m, 0.0);
  t.obj_value = cur_obj;
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
  if (obj) *obj = t.obj_value;
  if (status) *status = (int)st;
  if (x) memcpy(x, t.x_sol.data(), sizeof(double) * n);
  if (g && m) memcpy(g, t.g_sol.data(), sizeof(double) * m);
  if (duals) memcpy(duals, t.duals_sol.data(), sizeof(double) * (2 * n + m));
  return 0;
}

int bnlp_dive_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                     const double *gu, const double *x_sol, const double *g_sol, double obj_sol, const int *column_length,
                     double cutoff, bnlp_batch_solve_cb cb, bnlp_grad_f_cb grad_cb, void *ctx, const bnlp_dive_options *o,
                     bnlp_dive_result *res, double *better_x) {
  using namespace Bonmin;
  TMINLP2TNLP node;
  node.n = n; node.m = m;
  node.var_types.assign(var_type, var_type + n);
  node.x_l.assign(xl, xl + n); node.x_u.assign(xu, xu + n);
  node.g_l.assign(gl, gl + m); node.g_u.assign(gu, gu + m);
  node.x_init.assign(x_sol, x_sol + n);
  node.x_sol.assign(x_sol, x_sol + n);
  node.g_sol.assign(g_sol, g_sol + m);
  node.obj_value = obj_sol;
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
