// B-OA decomposition loop (see B200Oa.hpp).  Control flow restated from OACutGenerator2::performOa
// (src/Algorithms/OaGenerators/BonOACutGenerator2.cpp:121-268): solve the master, stop when it is infeasible under the
// cutoff or repeats its integer point, fix the integers (fixIntegers, BonSolverHelp.cpp:49-93), solve NLP(y), update
// incumbent and cutoff (post_nlp_solve, :146-151), linearise, install the cuts, go on.  No arithmetic happens here.
#include "B200Oa.hpp"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>

namespace Bonmin {

namespace {
double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

OaResult oa_run(const TMINLP2TNLP &root, const BatchSolveFn &solve, const OaCutsFn &oa, const OaExtractFn &extract,
                const MasterMipFn &master, const OaOptions &opt) {
  const double t0 = now_s();
  OaResult R;
  const int n = root.n, m = root.m;
  std::vector<int> ints;
  for (int j = 0; j < n; ++j)
    if (root.var_types[j] != 0) ints.push_back(j);
  const std::vector<double> &xl = root.x_l_, &xu = root.x_u_;

  // one NLP through the provider; keeps the multipliers of the LAST solve (the interface's getRowPrice())
  std::vector<double> xs(n), duals(2 * (size_t)n + m, 0.0);
  double nlp_obj = 0;
  int nlp_status = 0, nlp_iters = 0;
  auto solve_one = [&](const double *XL, const double *XU, const double *X0) {
    return solve(1, XL, XU, X0, &nlp_obj, &nlp_status, &nlp_iters, xs.data(), nullptr, duals.data());
  };
  auto finish = [&]() { R.seconds = now_s() - t0; return R; };

  // continuous relaxation; the initial LP is extracted at its optimum (BonOaDecBase / BonminSetup::initializeBHyb)
  std::vector<double> x0(n, 0.0);
  if (!root.x_init_.empty()) x0 = root.x_init_;
  if (int rc = solve_one(xl.data(), xu.data(), x0.data())) { R.status = rc < 0 ? rc : -1; return finish(); }
  R.nlp_solves = 1;
  if (nlp_status != 1 && nlp_status != 2) { R.status = nlp_status == 3 ? 1 : 2; return finish(); }
  LpRelaxation lp;
  if (int rc = extract(xs.data(), duals.data(), lp)) { R.status = rc < 0 ? rc : -1; return finish(); }
  const int ncols = n + (lp.has_alpha ? 1 : 0);
  // the growing master: rows of the relaxation, then every installed cut (installCuts: cuts are global)
  std::vector<int> row_start(lp.row_start), col(lp.col.begin(), lp.col.begin() + lp.row_start.back());
  std::vector<double> val(lp.val.begin(), lp.val.begin() + lp.row_start.back()), row_lb(lp.row_lb), row_ub(lp.row_ub);
  std::vector<double> obj(ncols, 0.0), col_lb(ncols), col_ub(ncols);
  for (int j = 0; j < ncols && j < (int)lp.obj.size(); ++j) obj[j] = lp.obj[j];
  for (int j = 0; j < n; ++j) { col_lb[j] = xl[j]; col_ub[j] = xu[j]; }
  if (ncols > n) { col_lb[n] = -DBL_MAX; col_ub[n] = DBL_MAX; }
  std::vector<int> integer(ncols, 0);
  for (int j : ints) integer[j] = 1;

  double cutoff = 1e100, ub = 1e100;
  std::vector<double> xm(ncols), y(ints.size()), last_y;
  bool have_last = false;
  std::vector<double> XL(xl), XU(xu), pts(2 * (size_t)n);
  CutRows cuts;
  while (R.passes < opt.max_passes && now_s() - t0 < opt.time_limit) {
    double bound = 1e200;
    const double tm = now_s();
    const int st = master(ncols, (int)row_lb.size(), row_start.data(), col.data(), val.data(), row_lb.data(), row_ub.data(),
                          col_lb.data(), col_ub.data(), obj.data(), lp.obj_offset, integer.data(), cutoff,
                          opt.time_limit - (tm - t0), xm.data(), &bound);   // subMip_->optimize(cutoff, ., maxTime)
    R.master_seconds += now_s() - tm;
    if (st < 0) { R.status = st; return finish(); }
    if (st != 0) {                        // master infeasible under the cutoff: the incumbent is optimal (:160-166)
      R.lower_bound = std::max(R.lower_bound, std::min(ub, 1e200));
      R.status = ub < 1e99 ? 0 : 1;
      break;
    }
    R.lower_bound = std::max(R.lower_bound, bound);
    ++R.passes;
    for (size_t k = 0; k < ints.size(); ++k) y[k] = std::nearbyint(xm[ints[k]]);
    if (have_last && y == last_y) {  // isDifferentOnIntegers == false: fathomed (:196-199)
      R.status = ub < 1e99 ? 0 : 1;
      break;
    }
    last_y = y;
    have_last = true;
    if (ub < 1e99 && (ub - bound) / std::max(std::fabs(ub), 1e-10) < opt.gap_tol) { R.status = 0; break; }
    // NLP(y): integers fixed at the master's values
    for (size_t k = 0; k < ints.size(); ++k) XL[ints[k]] = XU[ints[k]] = y[k];
    if (int rc = solve_one(XL.data(), XU.data(), xm.data())) { R.status = rc < 0 ? rc : -1; return finish(); }
    ++R.nlp_solves;
    if (nlp_status == 1 || nlp_status == 2) {    // post_nlp_solve: feasible -> incumbent and new cutoff
      if (nlp_obj < ub) { ub = nlp_obj; R.best_obj = ub; R.best_x = xs; }
      cutoff = ub > 0 ? ub * (1 - opt.cutoff_increment) : ub * (1 + opt.cutoff_increment);
      if (ub == 0) cutoff = -opt.cutoff_increment;
    }
    // linearise at the NLP(y) point (an infeasible NLP(y): at its minimum-violation point) and at the master point
    const int npts = opt.add_master_point ? 2 : 1;
    std::copy(xs.begin(), xs.end(), pts.begin());
    if (npts == 2)
      for (int j = 0; j < n; ++j) pts[n + j] = std::min(std::max(xm[j], xl[j]), xu[j]);
    cuts = CutRows();
    if (int rc = oa(npts, pts.data(), duals.data(), cuts)) { R.status = rc < 0 ? rc : -1; return finish(); }
    R.oa_points += npts;
    for (int k = 0; k < cuts.size(); ++k) {
      if (cuts.src_row[k] == m && ncols == n) continue;   // objective cut of a linear objective: not generated (:2053)
      for (int e = cuts.row_start[k]; e < cuts.row_start[k + 1]; ++e) { col.push_back(cuts.col[e]); val.push_back(cuts.val[e]); }
      row_start.push_back((int)col.size());
      row_lb.push_back(cuts.lb[k]); row_ub.push_back(cuts.ub[k]);
      ++R.cuts;
    }
  }
  return finish();
}

OaFeasResult oa_feas_check(const TMINLP2TNLP &root, OaLp &lp, const double *x_lp_in, double cutoff, const BatchSolveFn &solve,
                           const OaCutsFn &oa, const MasterMipFn &lp_solve, const OaFeasOptions &opt) {
  OaFeasResult R;
  const int n = root.n, m = root.m, ncols = lp.n_cols;
  std::vector<int> ints;
  for (int j = 0; j < n; ++j)
    if (root.var_types[j] != 0) ints.push_back(j);
  R.cutoff = cutoff;
  R.x_lp.assign(x_lp_in, x_lp_in + ncols);
  std::vector<double> xs(n), duals(2 * (size_t)n + m, 0.0), XL(root.x_l_), XU(root.x_u_), x_new(ncols);
  const std::vector<int> no_integer(ncols, 0);
  double nlp_obj = 0;
  int nlp_status = 0, nlp_iters = 0, passes = 0;
  bool is_integer = true, feasible = true;
  while (is_integer && feasible && passes < opt.max_passes) {
    ++passes;
    for (int j : ints) XL[j] = XU[j] = std::nearbyint(R.x_lp[j]);            // fixIntegers
    if (int rc = solve(1, XL.data(), XU.data(), R.x_lp.data(), &nlp_obj, &nlp_status, &nlp_iters, xs.data(), nullptr, duals.data())) {
      R.status = rc < 0 ? rc : -1;
      return R;
    }
    ++R.nlp_solves;
    if ((nlp_status == 1 || nlp_status == 2) && nlp_obj < R.cutoff) {         // post_nlp_solve
      if (!R.has_incumbent || nlp_obj < R.incumbent_obj) { R.has_incumbent = true; R.incumbent_obj = nlp_obj; R.incumbent_x = xs; }
      R.cutoff = nlp_obj > 0 ? nlp_obj * (1 - opt.cutoff_increment) : nlp_obj * (1 + opt.cutoff_increment);
    }
    CutRows cuts;
    if (int rc = oa(1, xs.data(), duals.data(), cuts)) { R.status = rc < 0 ? rc : -1; return R; }
    for (int k = 0; k < cuts.size(); ++k) {
      if (cuts.src_row[k] == m && ncols == n) continue;
      for (int e = cuts.row_start[k]; e < cuts.row_start[k + 1]; ++e) {
        lp.col.push_back(cuts.col[e]); lp.val.push_back(cuts.val[e]);
        R.new_rows.col.push_back(cuts.col[e]); R.new_rows.val.push_back(cuts.val[e]);
      }
      lp.row_start.push_back((int)lp.col.size()); lp.row_lb.push_back(cuts.lb[k]); lp.row_ub.push_back(cuts.ub[k]);
      R.new_rows.row_start.push_back((int)R.new_rows.col.size());
      R.new_rows.lb.push_back(cuts.lb[k]); R.new_rows.ub.push_back(cuts.ub[k]); R.new_rows.src_row.push_back(cuts.src_row[k]);
      ++R.cuts;
    }
    double val = 1e200;
    const int st = lp_solve(ncols, (int)lp.row_lb.size(), lp.row_start.data(), lp.col.data(), lp.val.data(), lp.row_lb.data(),
                            lp.row_ub.data(), lp.col_lb.data(), lp.col_ub.data(), lp.obj.data(), lp.obj_offset, no_integer.data(),
                            1e100, 1e10, x_new.data(), &val);                 // lp->resolve()
    if (st < 0) { R.status = st; return R; }
    feasible = st == 0 && val < R.cutoff;
    is_integer = false;
    if (feasible) {
      double diff = 0, frac = 0;
      for (int j : ints) {
        diff = std::max(diff, std::fabs(xs[j] - x_new[j]));
        frac = std::max(frac, std::fabs(x_new[j] - std::nearbyint(x_new[j])));
      }
      if (diff > 0.1) is_integer = frac <= opt.integer_tolerance;            // isDifferentOnIntegers(.., 0.1, ..)
      else R.fathomed = true;                                                // same integers as the NLP point (:115-119)
      R.x_lp = x_new;
    } else {
      R.fathomed = true;
      R.has_x_lp = false;
    }
  }
  return R;
}

OaNlpOptimResult oa_nlp_optim(const TMINLP2TNLP &root, const double *node_lb, const double *node_ub, const double *x_lp,
                              const BatchSolveFn &solve, const OaCutsFn &oa) {
  OaNlpOptimResult R;
  const int n = root.n, m = root.m;
  std::vector<double> XL(root.x_l_), XU(root.x_u_);
  for (int j = 0; j < n; ++j)
    if (root.var_types[j] != 0) { XL[j] = node_lb[j]; XU[j] = node_ub[j]; }
  R.x.assign(n, 0.0); R.duals.assign(2 * (size_t)n + m, 0.0);
  int iters = 0;
  if (int rc = solve(1, XL.data(), XU.data(), x_lp, &R.obj, &R.nlp_status, &iters, R.x.data(), nullptr, R.duals.data())) {
    R.rc = rc < 0 ? rc : -1;
    return R;
  }
  if (R.nlp_status == -3 || R.nlp_status == -2) { R.rc = -3; return R; }   // isAbandoned() || isIterationLimitReached()
  if (int rc = oa(1, R.x.data(), R.duals.data(), R.cuts)) { R.rc = rc < 0 ? rc : -1; return R; }
  if (R.nlp_status == 1 || R.nlp_status == 2) {
    double frac = 0;
    for (int j = 0; j < n; ++j)
      if (root.var_types[j] != 0) frac = std::max(frac, std::fabs(R.x[j] - std::floor(R.x[j] + 0.5)));
    R.has_solution = frac <= 1e-7;
  }
  return R;
}

OaEcpResult oa_ecp_rounds(int n, int m, OaLp &lp, const int *is_integer, const double *duals, int rounds, double abs_tol,
                          double rel_tol, const OaCutsFn &oa, const OaViolationFn &violation, const MasterMipFn &lp_solve) {
  OaEcpResult R;
  const int ncols = lp.n_cols;
  const std::vector<int> no_integer(ncols, 0);
  const std::vector<double> zero_duals(2 * (size_t)n + m, 0.0);
  if (!duals) duals = zero_duals.data();
  R.x.assign(ncols, 0.0);
  double val = 0;
  auto solve_lp = [&]() {
    return lp_solve(ncols, (int)lp.row_lb.size(), lp.row_start.data(), lp.col.data(), lp.val.data(), lp.row_lb.data(),
                    lp.row_ub.data(), lp.col_lb.data(), lp.col_ub.data(), lp.obj.data(), lp.obj_offset,
                    is_integer ? is_integer : no_integer.data(), 1e100, 1e10, R.x.data(), &val);
  };
  int st = solve_lp();
  if (st < 0) { R.rc = st; return R; }
  if (st != 0) { R.lp_infeasible = true; R.obj = DBL_MAX; return R; }
  auto alpha = [&]() { return ncols > n ? R.x[n] : val; };     // the LP's value of the objective (alpha column)
  double orig = 0;
  if (int rc = violation(R.x.data(), alpha(), &orig)) { R.rc = rc < 0 ? rc : -1; return R; }
  R.violation = orig;
  double viol = orig;
  for (int i = 0; i < rounds; ++i) {
    if (!(viol > abs_tol && viol > rel_tol * orig)) break;
    CutRows cuts;
    if (int rc = oa(1, R.x.data(), duals, cuts)) { R.rc = rc < 0 ? rc : -1; return R; }
    int kept = 0;
    const int first = R.cs.size();
    for (int k = 0; k < cuts.size(); ++k) {
      if (cuts.src_row[k] == m && ncols == n) continue;
      for (int e = cuts.row_start[k]; e < cuts.row_start[k + 1]; ++e) { R.cs.col.push_back(cuts.col[e]); R.cs.val.push_back(cuts.val[e]); }
      R.cs.row_start.push_back((int)R.cs.col.size());
      R.cs.lb.push_back(cuts.lb[k]); R.cs.ub.push_back(cuts.ub[k]); R.cs.src_row.push_back(cuts.src_row[k]);
      ++kept;
    }
    if (!kept || i + 1 >= rounds) break;
    for (int k = first; k < first + kept; ++k) {               // installCuts
      for (int e = R.cs.row_start[k]; e < R.cs.row_start[k + 1]; ++e) { lp.col.push_back(R.cs.col[e]); lp.val.push_back(R.cs.val[e]); }
      lp.row_start.push_back((int)lp.col.size()); lp.row_lb.push_back(R.cs.lb[k]); lp.row_ub.push_back(R.cs.ub[k]);
    }
    R.installed += kept;
    ++R.rounds;
    st = solve_lp();
    if (st < 0) { R.rc = st; return R; }
    if (st != 0) { R.lp_infeasible = true; R.obj = DBL_MAX; return R; }
    if (int rc = violation(R.x.data(), alpha(), &viol)) { R.rc = rc < 0 ? rc : -1; return R; }
  }
  R.violation = viol;
  R.obj = val;
  return R;
}

}  // namespace Bonmin
