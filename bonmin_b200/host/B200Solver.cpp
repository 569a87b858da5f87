#include "B200Solver.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <queue>

namespace Bonmin {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Contract of OptimizeTNLP (BonIpoptSolver.cpp:134-176, Filter/BonFilterSolver.cpp:574-645), driven purely through the
// Ipopt::TNLP virtuals like any Bonmin NLP plugin:
//   get_nlp_info -> get_bounds_info -> get_starting_point(init_x, init_z = init_lambda = warm start enabled)
//   -> solve -> finalize_solution(status, n, x, z_L, z_U, m, g, lambda, obj_value, NULL, NULL)
// (TMINLP2TNLP::finalize_solution, BonTMINLP2TNLP.cpp:474-509, stores x_sol / g_sol / duals_sol = [z_L | z_U | lambda] /
// obj_value).  The functions themselves are not evaluated through the TNLP: they live in the tape the engine was built from,
// so n and m must be the engine's.  The all-fixed short cut of TNLPSolver::zeroDimension (BonTNLPSolver.cpp:78-161) is taken
// inside the batch solve (same tests, same finalize_solution arguments).  A provider error maps to TNLPSolver::exception.
static Ipopt::SolverReturn ipopt_status_of(int st) {
  switch (st) {   // the inverse of IpoptSolver::solverReturnStatus (BonIpoptSolver.cpp:304-343)
    case TNLPSolver::solvedOptimal: return Ipopt::SUCCESS;
    case TNLPSolver::solvedOptimalTol: return Ipopt::STOP_AT_ACCEPTABLE_POINT;
    case TNLPSolver::provenInfeasible: return Ipopt::LOCAL_INFEASIBILITY;
    case TNLPSolver::iterationLimit: return Ipopt::MAXITER_EXCEEDED;
    case TNLPSolver::computationError: return Ipopt::ERROR_IN_STEP_COMPUTATION;
    case TNLPSolver::notEnoughFreedom: return Ipopt::TOO_FEW_DEGREES_OF_FREEDOM;
    case TNLPSolver::illegalOption: return Ipopt::INVALID_OPTION;
    default: return Ipopt::INTERNAL_ERROR;
  }
}

TNLPSolver::ReturnStatus B200Solver::solve_tnlp(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp, bool warm) {
  const double t0 = now_s();
  Ipopt::Index n = 0, m = 0, nnz_jac = 0, nnz_h = 0;
  Ipopt::TNLP::IndexStyleEnum style;
  iterations_ = 0;
  if (!tnlp->get_nlp_info(n, m, nnz_jac, nnz_h, style)) return illDefinedProblem;
  std::vector<double> xl(n), xu(n), gl(std::max(m, 1)), gu(std::max(m, 1)), x0(n), d0((size_t)2 * n + m, 0.0);
  if (!tnlp->get_bounds_info(n, xl.data(), xu.data(), m, gl.data(), gu.data())) return illDefinedProblem;
  // warm start: primal AND dual start (warm_start_init_point yes, BonIpoptSolver.cpp:415-419); a TNLP that has no
  // multipliers to give answers false, and the solve falls back to the primal start alone
  bool have_duals = warm && tnlp->get_starting_point(n, true, x0.data(), true, d0.data(), d0.data() + n, m, true, d0.data() + 2 * n);
  if (!have_duals && !tnlp->get_starting_point(n, true, x0.data(), false, nullptr, nullptr, m, false, nullptr)) return illDefinedProblem;
  std::vector<double> x(n, 0.0), g(std::max(m, 1), 0.0), du((size_t)2 * n + m, 0.0);
  double obj = 0;
  int st = 0, it = 0, rc;
  if (have_duals && solve_warm_) rc = solve_warm_(1, xl.data(), xu.data(), x0.data(), d0.data(), &obj, &st, &it, x.data(), g.data(), du.data());
  else rc = solve_(1, xl.data(), xu.data(), x0.data(), &obj, &st, &it, x.data(), g.data(), du.data());
  warm_started_last_ = have_duals && (bool)solve_warm_;
  cpu_time_ = now_s() - t0;
  iterations_ = it;
  optimized_before_ = true;
  if (rc != 0) { provider_errors_++; error_code_ = rc; return exception; }
  // (for an infeasible child the engine's objective is the maximal constraint violation, which is what finalize_solution
  // computes from ip_cq in the reference, BonTMINLP2TNLP.cpp:503-505)
  tnlp->finalize_solution(ipopt_status_of(st), n, x.data(), du.data(), du.data() + n, m, g.data(), du.data() + 2 * n, obj, nullptr, nullptr);
  return (ReturnStatus)st;
}

TNLPSolver::ReturnStatus B200Solver::OptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) {
  // IpoptSolver::OptimizeTNLP (:134-176): a warm start is used only when it is enabled AND the solver has optimised before
  return solve_tnlp(tnlp, enable_warm_start_ && optimized_before_);
}
TNLPSolver::ReturnStatus B200Solver::ReOptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) {
  // IpoptSolver::ReOptimizeTNLP (:180-220): re-optimises whenever it has optimised before; whether multipliers are used is
  // the warm_start_init_point option that enableWarmStart / disableWarmStart set (:415-427)
  return solve_tnlp(tnlp, enable_warm_start_ && optimized_before_);
}

bool B200Solver::Initialize(std::istream &is) {
  std::string line;
  while (std::getline(is, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.erase(h);
    std::istringstream ls(line);
    std::string name, value;
    if (!(ls >> name)) continue;
    if (!(ls >> value)) return false;          // an option without a value is an error for Ipopt's reader too
    if (name.rfind("bonmin.", 0) == 0) name.erase(0, 7);
    bool found = false;
    for (auto &kv : options_) if (kv.first == name) { kv.second = value; found = true; }
    if (!found) options_.push_back({name, value});
  }
  if (const std::string *v = option("print_level")) default_log_level_ = log_level_ = atoi(v->c_str());
  return true;
}

bool B200Solver::Initialize(std::string params_file) {
  std::ifstream f(params_file);
  if (!f) return true;                         // a missing options file is not an error (Ipopt: "no options file found")
  return Initialize(f);
}

const std::string *B200Solver::option(const std::string &name) const {
  for (auto &kv : options_) if (kv.first == name) return &kv.second;
  return nullptr;
}

bool B200Solver::setWarmStart(const CoinWarmStart *warmstart, Ipopt::SmartPtr<TMINLP2TNLP> tnlp) {   // BonIpoptSolver.cpp:369-401
  if (!warmstart) return false;
  const IpoptWarmStart *ws = dynamic_cast<const IpoptWarmStart *>(warmstart);
  if (ws == nullptr) return false;
  if (ws->empty()) { disableWarmStart(); return true; }   // reset initial point and leave
  if (ws->dualSize() > 0) { tnlp->setDualsInit(ws->dualSize(), ws->dual()); enableWarmStart(); }
  else disableWarmStart();
  if (ws->primalSize() != tnlp->num_variables()) return false;
  tnlp->setxInit(ws->primalSize(), ws->primal());
  return true;
}

CoinWarmStart *B200Solver::getUsedWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const {        // :347-356
  if (tnlp->x_init() == nullptr || tnlp->duals_init() == nullptr) return nullptr;
  return new IpoptWarmStart(tnlp->num_variables(), 2 * tnlp->num_variables() + tnlp->num_constraints(), tnlp->x_init(),
                            tnlp->duals_init());
}

CoinWarmStart *B200Solver::getWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const {            // :357-366: the last solution and its multipliers
  return new IpoptWarmStart(tnlp, nullptr);
}

bool B200Solver::warmStartIsValid(const CoinWarmStart *ws) const {                            // :403-410
  const IpoptWarmStart *iws = dynamic_cast<const IpoptWarmStart *>(ws);
  return iws && !iws->empty();
}

const std::string &B200Solver::UnsolvedB200Error::errorName() const {
  static const std::string provider = "the batch-solve provider failed (see bnlp_last_error)", solver = "unrecoverable solver status";
  return errorNum() > -100 && errorNum() < 0 ? solver : provider;
}
const std::string &B200Solver::UnsolvedB200Error::solverName() const {
  static const std::string name = "B200 batched IPM";
  return name;
}

std::shared_ptr<std::vector<double>> B200Solver::x_block(size_t cnt) {
  // the previous batch's block is reused once nobody refers to it any more
  if (!x_pool_ || x_pool_.use_count() > 1 || x_pool_->size() < cnt) x_pool_ = std::make_shared<std::vector<double>>(cnt);
  return x_pool_;
}

int B200Solver::unpack(int rc, int B, int n, std::vector<Child> &out) {
  provider_errors_ += (rc != 0);
  out.resize(B);
  for (int b = 0; b < B; ++b) {
    out[b].obj = obj_buf_[b];
    out[b].status = rc ? exception : (ReturnStatus)st_buf_[b];
    out[b].iters = it_buf_[b];
    out[b].block = x_pool_;
    out[b].off = (size_t)b * n;
    total_batch_iterations_ += it_buf_[b];
  }
  return rc;
}

int B200Solver::OptimizeBatch(const TMINLP2TNLP &p, const std::vector<double> &XL, const std::vector<double> &XU, int B,
                              std::vector<Child> &out) {
  const int n = p.n;
  if ((int)obj_buf_.size() < B) { obj_buf_.resize(B); st_buf_.resize(B); it_buf_.resize(B); }
  double *X = x_block((size_t)B * n)->data();
  const double ts0 = now_s();
  const int rc = solve_(B, XL.data(), XU.data(), p.x_init_.data(), obj_buf_.data(), st_buf_.data(), it_buf_.data(), X, nullptr, nullptr);
  solve_seconds_ += now_s() - ts0;
  return unpack(rc, B, n, out);
}

int B200Solver::OptimizeChildren(const TMINLP2TNLP &p, int P, const std::vector<double> &PXL, const std::vector<double> &PXU,
                                 const std::vector<int> &cp, const std::vector<int> &cc, const std::vector<int> &cw,
                                 const std::vector<double> &cb, std::vector<Child> &out) {
  const int n = p.n, B = (int)cp.size();
  if ((int)obj_buf_.size() < B) { obj_buf_.resize(B); st_buf_.resize(B); it_buf_.resize(B); }
  double *X = x_block((size_t)B * n)->data();
  int rc;
  const double ts0 = now_s();
  if (children_) {
    rc = children_(P, PXL.data(), PXU.data(), B, cp.data(), cc.data(), cw.data(), cb.data(), p.x_init_.data(), obj_buf_.data(),
                   st_buf_.data(), it_buf_.data(), X);
  } else {
    if (xl_buf_.size() < (size_t)B * n) { xl_buf_.resize((size_t)B * n); xu_buf_.resize((size_t)B * n); }
    for (int b = 0; b < B; ++b) {
      std::copy(&PXL[(size_t)cp[b] * n], &PXL[(size_t)cp[b] * n] + n, &xl_buf_[(size_t)b * n]);
      std::copy(&PXU[(size_t)cp[b] * n], &PXU[(size_t)cp[b] * n] + n, &xu_buf_[(size_t)b * n]);
      (cw[b] == 0 ? xu_buf_ : xl_buf_)[(size_t)b * n + cc[b]] = cb[b];
    }
    rc = solve_(B, xl_buf_.data(), xu_buf_.data(), p.x_init_.data(), obj_buf_.data(), st_buf_.data(), it_buf_.data(), X, nullptr, nullptr);
  }
  solve_seconds_ += now_s() - ts0;
  return unpack(rc, B, n, out);
}

void BatchedStrongBranchingSolver::markHotStart(OsiTMINLPInterface *si) {
  const int n = si->getNumCols();
  xl0_.assign(si->getColLower(), si->getColLower() + n);
  xu0_.assign(si->getColUpper(), si->getColUpper() + n);
  if (qp_) {   // BranchingTQP's constructor: freeze the node optimum and its multipliers (BonBranchingTQP.cpp:41-43)
    xsol0_ = si->problem()->x_sol_;
    duals0_ = si->problem()->duals_sol_;
  }
  if (!preloaded_) cache_.clear();
}

void BatchedStrongBranchingSolver::prefetchMany(OsiTMINLPInterface *si, const std::vector<NodeRequest> &reqs) {
  const int n = si->getNumCols(), P = (int)reqs.size();
  size_t total = 0;
  for (auto &r : reqs) total += 2 * r.cols.size();
  std::vector<double> PXL((size_t)P * n), PXU((size_t)P * n), cb(total);
  std::vector<int> cp(total), cc(total), cw(total);
  size_t row = 0;
  for (int q = 0; q < P; ++q) {
    const NodeRequest &r = reqs[q];
    std::copy(r.xl.begin(), r.xl.end(), &PXL[(size_t)q * n]);
    std::copy(r.xu.begin(), r.xu.end(), &PXU[(size_t)q * n]);
    for (size_t i = 0; i < r.cols.size(); ++i)
      for (int w = 0; w < 2; ++w, ++row) {
        cp[row] = q; cc[row] = r.cols[i]; cw[row] = w;
        cb[row] = (w == 0) ? r.down_ub[i] : r.up_lb[i];
      }
  }
  multi_.clear();   // drop the references to the previous round's result block so that it can be reused
  cache_.clear();
  std::vector<B200Solver::Child> res;
  if (total) app_->OptimizeChildren(*si->problem(), P, PXL, PXU, cp, cc, cw, cb, res);
  multi_.assign(reqs.size(), {});
  multi_bounds_.clear();
  row = 0;
  for (size_t q = 0; q < reqs.size(); ++q) {
    multi_bounds_.push_back({reqs[q].xl, reqs[q].xu});
    for (size_t i = 0; i < reqs[q].cols.size(); ++i)
      for (int w = 0; w < 2; ++w, ++row) multi_[q].emplace_hint(multi_[q].end(), std::make_pair(reqs[q].cols[i], w), std::move(res[row]));
  }
}

void BatchedStrongBranchingSolver::useNode(int i) {
  cache_ = std::move(multi_[i]);
  xl0_ = multi_bounds_[i].first; xu0_ = multi_bounds_[i].second;
  preloaded_ = true;
}

void BatchedStrongBranchingSolver::prefetch(OsiTMINLPInterface *si, const std::vector<int> &cols,
                                            const std::vector<double> &down_ub, const std::vector<double> &up_lb) {
  const int n = si->getNumCols(), K = (int)cols.size();
  std::vector<double> XL((size_t)2 * K * n), XU((size_t)2 * K * n);
  for (int i = 0; i < K; ++i)
    for (int w = 0; w < 2; ++w) {
      double *l = &XL[((size_t)2 * i + w) * n], *u = &XU[((size_t)2 * i + w) * n];
      std::copy(xl0_.begin(), xl0_.end(), l);
      std::copy(xu0_.begin(), xu0_.end(), u);
      if (w == 0) u[cols[i]] = down_ub[i]; else l[cols[i]] = up_lb[i];
    }
  std::vector<B200Solver::Child> res;
  if (qp_ && (int)xsol0_.size() == n && (int)duals0_.size() == 2 * n + si->problem()->m) {
    const int B = 2 * K;
    std::vector<double> OBJ(B);
    auto blk = std::make_shared<std::vector<double>>((size_t)B * n);
    std::vector<int> ST(B), IT(B);
    const int rc = qp_(xsol0_.data(), duals0_.data(), B, XL.data(), XU.data(), OBJ.data(), ST.data(), IT.data(), blk->data());
    res.resize(B);
    for (int b = 0; b < B; ++b) {
      res[b].obj = OBJ[b];
      res[b].status = rc ? TNLPSolver::exception : (TNLPSolver::ReturnStatus)ST[b];
      res[b].iters = IT[b];
      res[b].block = blk; res[b].off = (size_t)b * n;
      qp_iterations_ += IT[b];
    }
  } else {
    app_->OptimizeBatch(*si->problem(), XL, XU, 2 * K, res);
  }
  for (int i = 0; i < K; ++i)
    for (int w = 0; w < 2; ++w) cache_[{cols[i], w}] = std::move(res[2 * i + w]);
}

const B200Solver::Child *BatchedStrongBranchingSolver::lookup(int col, bool up) const {
  auto it = cache_.find({col, up ? 1 : 0});
  return it == cache_.end() ? nullptr : &it->second;
}

// solveFromHotStart must leave the child's objective in problem()->set_obj_value and, if optimal, the point in
// Set_x_sol, because HotInfo::updateInformation reads getObjValue()/isProvenOptimal() right after
// (BonChooseVariable.cpp:1131-1142; cf. BonLpBranchingSolver.cpp:181-182).
TNLPSolver::ReturnStatus BatchedStrongBranchingSolver::solveFromHotStart(OsiTMINLPInterface *si) {
  const int n = si->getNumCols();
  const double *l = si->getColLower(), *u = si->getColUpper();
  int col = -1, way = 0, ndiff = 0;
  for (int j = 0; j < n; ++j) {
    if (u[j] != xu0_[j]) { col = j; way = 0; ndiff++; }
    if (l[j] != xl0_[j]) { col = j; way = 1; ndiff++; }
  }
  const B200Solver::Child *c = (ndiff == 1) ? lookup(col, way == 1) : nullptr;
  if (!c) {  // not a prefetched child: solve it on the spot (batch of one) — same answer, just not amortised
    misses_++;
    return si->solver()->OptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(si->problem()));
  }
  hits_++;
  si->problem()->set_obj_value(c->obj);
  if (c->status == TNLPSolver::solvedOptimal || c->status == TNLPSolver::solvedOptimalTol) si->problem()->Set_x_sol(n, c->x());
  si->problem()->return_status = c->status;
  return c->status;
}

void BatchedStrongBranchingSolver::unmarkHotStart(OsiTMINLPInterface *) { preloaded_ = false; }

// BonChooseVariable::doStrongBranching (BonChooseVariable.cpp:787-957) with the 2 x numberToDo solveFromHotStart calls served
// from ONE prefetched batch.  Everything else is the reference's loop: branch down / solve / updateInformation / restore,
// branch up / solve / updateInformation / restore; returnCode -1 when a candidate is infeasible both ways (the loop goes on,
// :916-919), 1 when exactly one way is infeasible (:920-923), 3 and stop when the time is up (:924-929), stop early by the
// look-ahead rule away from the root (:930-948); results_ is cut to the candidates actually evaluated (:950-952).
int BatchedChooseVariable::doStrongBranching(OsiTMINLPInterface *solver, OsiBranchingInformation *info, int numberToDo,
                                             int /* returnCriterion */) {
  double bestLookAhead = -1.7976931348623157e308;
  int trialsSinceBest = 0;
  const bool isRoot = info->depth_ == 0;                  // isRootNode (:959-962)
  const int numberColumns = solver->getNumCols();
  numberToDo = std::min(numberToDo, (int)results_.size());
  std::vector<int> cols(numberToDo);
  std::vector<double> dn(numberToDo), up(numberToDo);
  for (int i = 0; i < numberToDo; ++i) {
    cols[i] = results_[i].col;
    dn[i] = std::floor(results_[i].value); up[i] = std::ceil(results_[i].value);
  }
  solver->markHotStart();                                 // :799
  if (!sb_->preloaded()) sb_->prefetch(solver, cols, dn, up);   // ONE batch of 2 x numberToDo children
  const std::vector<double> saveLower(info->lower_, info->lower_ + numberColumns), saveUpper(info->upper_, info->upper_ + numberColumns);
  int returnCode = 0;
  const double timeStart = now_s();
  int iDo = 0;
  for (; iDo < numberToDo; ++iDo) {
    StrongResult &r = results_[iDo];
    int status[2];
    for (int w = 0; w < 2; ++w) {
      if (w == 0) solver->setColUpper(r.col, dn[iDo]); else solver->setColLower(r.col, up[iDo]);   // branch->branch(solver) :820, :866
      solver->solveFromHotStart();                        // :822, :869
      // HotInfo::updateInformation (:1118-1145) + OsiHotInfo: 0 finished, 1 infeasible, 2 not finished
      const double obj = solver->getObjValue();
      if (solver->isProvenOptimal()) status[w] = 0;
      else if (solver->isProvenPrimalInfeasible()) status[w] = (std::fabs(obj) < 1e-6) ? 2 : 1;
      else status[w] = 2;
      r.status[w] = status[w];
      r.obj[w] = obj;
      r.change[w] = (status[w] == 0) ? std::max(0.0, obj - info->objectiveValue_) : 0.0;
      numberStrongDone_ += (w == 1);
      for (int j = 0; j < numberColumns; ++j) {           // restore bounds :853-858, :898-903
        if (saveLower[j] != solver->getColLower()[j]) solver->setColLower(j, saveLower[j]);
        if (saveUpper[j] != solver->getColUpper()[j]) solver->setColUpper(j, saveUpper[j]);
      }
    }
    if (status[0] == 1 && status[1] == 1) returnCode = -1;                      // infeasible both ways (no break, :916-919)
    else if (status[0] == 1 || status[1] == 1) { numberStrongFixed_++; returnCode = 1; }
    const bool hitMaxTime = (now_s() - timeStart > info->timeRemaining_) || (now_s() - start_time_ > time_limit_);
    if (hitMaxTime) { returnCode = 3; break; }
    if (!isRoot && numberLookAhead_ && status[0] == 0 && status[1] == 0) {      // look-ahead heuristic (:930-948)
      const double MAXMIN = (info->hasSolution_ && info->depth_ > 10) ? maxmin_crit_have_sol : maxmin_crit_no_sol;
      const double value = MAXMIN * std::min(r.change[1], r.change[0]) + (1.0 - MAXMIN) * std::max(r.change[1], r.change[0]);
      if (value > bestLookAhead) { bestLookAhead = value; trialsSinceBest = 0; }
      else if (++trialsSinceBest >= numberLookAhead_) break;
    }
  }
  if (iDo < numberToDo) iDo++;
  results_.resize(iDo);
  solver->unmarkHotStart();                               // :955
  return returnCode;
}

// the form the host B-BB driver uses: candidates as a column list; every candidate is evaluated
int BatchedChooseVariable::doStrongBranching(const std::vector<int> &cols, double parent_obj,
                                             std::vector<StrongResult> &results) {
  setCandidates(cols);
  OsiBranchingInformation info;
  info.lower_ = si_->getColLower(); info.upper_ = si_->getColUpper(); info.solution_ = si_->getColSolution();
  info.objectiveValue_ = parent_obj; info.depth_ = depth_; info.hasSolution_ = have_solution_;
  info.timeRemaining_ = time_remaining_;
  // (info.lower_ / upper_ point into the live bound arrays of the shim: take the snapshot the reference's info holds)
  const std::vector<double> lo(info.lower_, info.lower_ + si_->getNumCols()), hi(info.upper_, info.upper_ + si_->getNumCols());
  info.lower_ = lo.data(); info.upper_ = hi.data();
  const int rc = doStrongBranching(si_, &info, (int)cols.size(), 0);
  results = results_;
  return rc;
}

void BatchedChooseVariable::setCandidates(const std::vector<int> &cols) {
  results_.assign(cols.size(), StrongResult());
  for (size_t i = 0; i < cols.size(); ++i) {
    results_[i].col = cols[i];
    results_[i].value = si_->getColSolution()[cols[i]];
    results_[i].status[0] = results_[i].status[1] = -1;
  }
}

int BatchedChooseVariable::choose(const std::vector<StrongResult> &results, double parent_obj, double cutoff,
                                  bool have_solution, int depth, int &bestWhichWay,
                                  std::vector<std::pair<int, int>> &fixings) const {
  const double MAXMIN = (have_solution && depth > 10) ? maxmin_crit_have_sol : maxmin_crit_no_sol;   // :965-977
  double bestTrusted = -1e300;
  int best = -1;
  fixings.clear();
  for (size_t i = 0; i < results.size(); ++i) {
    const StrongResult &r = results[i];
    double est[2];
    for (int w = 0; w < 2; ++w) {
      if (r.status[w] != 1) est[w] = r.change[w];
      else {  // infeasible - just say expensive (:692-697, :716-721), and fix the variable the other way
        est[w] = (cutoff < 1e50) ? 2.0 * (cutoff - parent_obj) : 2.0 * std::fabs(parent_obj);
        fixings.push_back({r.col, w});
      }
    }
    const double value = MAXMIN * std::min(est[1], est[0]) + (1.0 - MAXMIN) * std::max(est[1], est[0]);
    if (value > bestTrusted) { bestTrusted = value; best = (int)i; bestWhichWay = est[1] > est[0] ? 0 : 1; }
  }
  return best;
}

// ------------------------------------------------------------------------------------------------ host B-BB driver
namespace {
struct Node {
  std::vector<double> xl, xu;
  double bound;      // parent objective (lower bound)
  int depth;
  // result pre-solved by strong branching (served from the SB cache instead of re-solving: SURVEY Appendix C note 2)
  bool solved = false;
  double obj = 0;
  int status = 0;
  std::vector<double> x;
};
struct Cmp { bool operator()(const Node *a, const Node *b) const { return a->bound > b->bound; } };
}  // namespace

// Integer variables that violate integrality, most fractional first (setupList's second list, BonChooseVariable.cpp:357-360);
// with priorities only those of the best (lowest) priority level stay (:330-352: a better level resets the list).
static std::vector<std::pair<double, int>> fractional_candidates(const std::vector<double> &x, const std::vector<int> &ints,
                                                                  const BbOptions &opt) {
  std::vector<std::pair<double, int>> frac;
  const bool pri = (int)opt.priorities.size() == (int)x.size();
  int best = 0;
  for (int j : ints) {
    const double f = std::fabs(x[j] - std::round(x[j]));
    if (f <= opt.integer_tolerance) continue;
    if (pri) {
      if (frac.empty() || opt.priorities[j] < best) { frac.clear(); best = opt.priorities[j]; }
      else if (opt.priorities[j] > best) continue;
    }
    frac.push_back({-f, j});
  }
  std::sort(frac.begin(), frac.end());
  return frac;
}

BbResult bb_run(TMINLP2TNLP &root, BatchSolveFn solve, const BbOptions &opt, TqpSolveFn tqp, ChildrenSolveFn children) {
  BbResult R;
  const double t0 = now_s();
  const int n = root.n;
  Ipopt::SmartPtr<B200Solver> app_holder = new B200Solver(solve);   // (SmartPtr-managed objects live on the heap)
  B200Solver &app = *app_holder;
  if (children) app.setChildrenSolver(children);
  OsiTMINLPInterface si(Ipopt::SmartPtr<TMINLP2TNLP>(&root), Ipopt::SmartPtr<TNLPSolver>(GetRawPtr(app_holder)));
  BatchedStrongBranchingSolver sb(&app);
  si.SetStrongBrachingSolver(&sb);
  const bool qp_mode = opt.variable_selection == 1 && tqp;
  if (qp_mode) sb.setQpSurrogate(tqp);
  BatchedChooseVariable chooser(&si, &sb);
  chooser.integer_tolerance = opt.integer_tolerance;
  chooser.setNumberLookAhead(opt.number_look_ahead);
  chooser.setTimeLimit(t0, opt.time_limit);
  std::vector<int> ints;
  for (int j = 0; j < n; ++j) if (root.var_types[j] != 0) ints.push_back(j);
  std::vector<double> pc_sum[2], pc_cnt[2];   // pseudo-costs per column and direction
  for (int w = 0; w < 2; ++w) { pc_sum[w].assign(n, 0.0); pc_cnt[w].assign(n, 0.0); }

  // root relaxation (Bab::branchAndBound: model_.initialSolve, then the root primal becomes the start point of every
  // later NLP: src/CbcBonmin/BonCbc.cpp:452,492-505)
  si.initialSolve();
  R.nlp_solves++;
  if (!si.isProvenOptimal()) { R.status = app.providerErrors() ? 3 : si.isProvenPrimalInfeasible() ? 1 : 4; R.failed_nodes += !app.providerErrors() && !si.isProvenPrimalInfeasible(); R.seconds = now_s() - t0; return R; }
  root.setxInit(n, root.x_sol_.data());
  std::priority_queue<Node *, std::vector<Node *>, Cmp> heap;
  Node *first = new Node{root.x_l_, root.x_u_, root.obj_value_, 0};
  first->solved = true; first->obj = root.obj_value_; first->status = (int)root.return_status; first->x = root.x_sol_;
  Node *dive = first;   // node processed next without going through the heap (depth-first plunge on the better child)
  double cutoff = 1e100;
  if (!opt.incumbent_x.empty() && opt.incumbent_obj < 1e99) {   // a heuristic's solution: CbcModel::setBestSolution before the tree
    R.best_obj = opt.incumbent_obj; R.best_x = opt.incumbent_x; cutoff = opt.incumbent_obj - opt.cutoff_decr;
  }
  const std::vector<double> xl_root = root.x_l_, xu_root = root.x_u_;

  double t_prefetch = 0;
  if (opt.parallel_nodes > 1 && !qp_mode) {
    // ---- round-based expansion: the P best open nodes per round; their node NLPs are one batch and the strong-
    // branching children of all of them are one batch (independent relaxations sharing tape and start point)
    heap.push(first);
    dive = nullptr;
    auto candidates_of = [&](const std::vector<double> &x) { return fractional_candidates(x, ints, opt); };
    while (!heap.empty()) {
      if (R.nodes >= opt.node_limit || now_s() - t0 > opt.time_limit) { R.status = 2; break; }
      if (app.providerErrors()) break;   // a failed device call is an error, not an infeasible node
      std::vector<Node *> batch;
      while (!heap.empty() && (int)batch.size() < opt.parallel_nodes) {
        Node *nd = heap.top(); heap.pop();
        if (nd->bound >= cutoff) { delete nd; continue; }
        batch.push_back(nd);
      }
      if (batch.empty()) break;
      // 1. node relaxations that were not pre-solved by strong branching
      std::vector<int> need;
      for (size_t i = 0; i < batch.size(); ++i) if (!batch[i]->solved) need.push_back((int)i);
      if (!need.empty()) {
        std::vector<double> XL(need.size() * n), XU(need.size() * n);
        for (size_t q = 0; q < need.size(); ++q) {
          std::copy(batch[need[q]]->xl.begin(), batch[need[q]]->xl.end(), &XL[q * n]);
          std::copy(batch[need[q]]->xu.begin(), batch[need[q]]->xu.end(), &XU[q * n]);
        }
        std::vector<B200Solver::Child> res;
        app.OptimizeBatch(root, XL, XU, (int)need.size(), res);
        R.nlp_solves += (int)need.size();
        for (size_t q = 0; q < need.size(); ++q) {
          Node *nd = batch[need[q]];
          nd->solved = true; nd->obj = res[q].obj; nd->status = (int)res[q].status; nd->x.assign(res[q].x(), res[q].x() + n);
        }
      } 
      // 2. fathoming, incumbents, candidate lists.  Only a PROVEN infeasible node is fathomed; a node whose NLP ended in an
      // abandoned status (iteration limit, computation error: OsiTMINLPInterface::isAbandoned) is an unsolved NLP — the
      // reference stops (nlp_failure_behavior stop) or drops it with a warning (fathom), it never reads it as infeasible
      std::vector<Node *> open;
      std::vector<BatchedStrongBranchingSolver::NodeRequest> reqs;
      bool stop_unsolved = false;
      for (Node *nd : batch) {
        R.nodes++;
        R.served_from_sb_cache += (nd != first && need.empty());
        const bool optimal = nd->status == TNLPSolver::solvedOptimal || nd->status == TNLPSolver::solvedOptimalTol;
        if (!optimal && nd->status != TNLPSolver::provenInfeasible) {
          R.failed_nodes++;
          if (opt.nlp_failure_behavior == 0) stop_unsolved = true;
          delete nd;
          continue;
        }
        if (!optimal || nd->obj >= cutoff) { delete nd; continue; }
        auto frac = candidates_of(nd->x);
        if (frac.empty()) {
          if (nd->obj < R.best_obj) { R.best_obj = nd->obj; R.best_x = nd->x; cutoff = nd->obj - opt.cutoff_decr; }
          delete nd;
          continue;
        }
        BatchedStrongBranchingSolver::NodeRequest rq;
        rq.xl = nd->xl; rq.xu = nd->xu;
        for (auto &f : frac) {
          if ((int)rq.cols.size() >= opt.number_strong_branch) break;
          rq.cols.push_back(f.second);
          rq.down_ub.push_back(std::floor(nd->x[f.second]));
          rq.up_lb.push_back(std::ceil(nd->x[f.second]));
        }
        open.push_back(nd);
        reqs.push_back(std::move(rq));
      }
      if (stop_unsolved) { for (Node *nd : open) delete nd; R.status = 4; break; }
      if (open.empty()) continue;
      // 3. ONE batch for the strong-branching children of every open node of the round
      const double tp0 = now_s();
      sb.prefetchMany(&si, reqs);
      t_prefetch += now_s() - tp0;
      R.sb_batches++;
      for (auto &rq : reqs) R.sb_children += 2 * (int)rq.cols.size();
      // 4. per node: the reference's sequential loop replayed against its cache, scoring, branching
      for (size_t q = 0; q < open.size(); ++q) {
        Node *nd = open[q];
        if (nd->obj >= cutoff) { delete nd; continue; }
        root.x_l_ = nd->xl; root.x_u_ = nd->xu;
        root.obj_value_ = nd->obj; root.x_sol_ = nd->x; root.return_status = (TNLPSolver::ReturnStatus)nd->status;
        si.optimizationStatus_ = (TNLPSolver::ReturnStatus)nd->status;
        sb.useNode((int)q);
        std::vector<StrongResult> res;
        chooser.setNodeInfo(nd->depth, R.best_obj < 1e99, opt.time_limit - (now_s() - t0));
        const int rc = chooser.doStrongBranching(reqs[q].cols, nd->obj, res);
        if (rc == -1) { delete nd; continue; }
        if (rc == 3) { R.status = 2; heap.push(nd); break; }        // time limit inside strong branching (:924-929)
        int bway = 0;
        std::vector<std::pair<int, int>> fix;
        const int bi = chooser.choose(res, nd->obj, cutoff, R.best_obj < 1e99, nd->depth, bway, fix);
        bool fixed_any = !fix.empty();
        for (auto &f : fix) {
          const double v = nd->x[f.first];
          if (f.second == 0) nd->xl[f.first] = std::max(nd->xl[f.first], std::ceil(v));
          else nd->xu[f.first] = std::min(nd->xu[f.first], std::floor(v));
        }
        const int bcol = res[bi].col;
        for (int w = 0; w < 2; ++w) {
          Node *ch = new Node{nd->xl, nd->xu, nd->obj, nd->depth + 1};
          if (w == 0) ch->xu[bcol] = std::floor(nd->x[bcol]); else ch->xl[bcol] = std::ceil(nd->x[bcol]);
          const B200Solver::Child *c = fixed_any ? nullptr : sb.lookup(bcol, w == 1);
          // a strong-branching solve that did not finish (HotInfo status 2) says nothing about the child: it goes to the
          // heap unsolved and gets its own node NLP
          if (c && (c->status == 1 || c->status == 2 || c->status == TNLPSolver::provenInfeasible)) {
            ch->solved = true; ch->obj = c->obj; ch->status = (int)c->status; ch->x.assign(c->x(), c->x() + n);
            if (c->status == 1 || c->status == 2) ch->bound = std::max(nd->obj, c->obj);
          }
          if (ch->solved && ch->status == TNLPSolver::provenInfeasible) { delete ch; continue; }   // pre-solved and infeasible
          heap.push(ch);
        }
        delete nd;
      }
    }
  }
  while ((opt.parallel_nodes <= 1 || qp_mode) && (dive || !heap.empty())) {
    Node *nd = dive;
    dive = nullptr;
    if (!nd) { nd = heap.top(); heap.pop(); }
    if (nd->bound >= cutoff) { delete nd; continue; }
    if (R.nodes >= opt.node_limit || now_s() - t0 > opt.time_limit) { R.status = 2; delete nd; break; }
    if (app.providerErrors()) { delete nd; break; }
    R.nodes++;
    root.x_l_ = nd->xl; root.x_u_ = nd->xu;
    if (nd->solved) {
      R.served_from_sb_cache += (nd != first);
      root.obj_value_ = nd->obj; root.x_sol_ = nd->x; root.return_status = (TNLPSolver::ReturnStatus)nd->status;
      si.optimizationStatus_ = (TNLPSolver::ReturnStatus)nd->status;
    } else {
      si.resolve();
      R.nlp_solves++;
    }
    if (!si.isProvenOptimal() && !si.isProvenPrimalInfeasible()) {
      // unsolved NLP (abandoned status): never fathomed as infeasible.  nlp_failure_behavior stop = the reference's
      // newUnsolvedError (BonOsiTMINLPInterface.cpp:2639-2644), fathom = pretendFailIsInfeasible_
      R.failed_nodes++;
      delete nd;
      if (opt.nlp_failure_behavior == 0) { R.status = 4; break; }
      continue;
    }
    if (!si.isProvenOptimal()) { delete nd; continue; }           // proven infeasible node: fathom
    const double obj = root.obj_value_;
    if (obj >= cutoff) { delete nd; continue; }
    const std::vector<double> x = root.x_sol_;
    // integer infeasibilities
    std::vector<std::pair<double, int>> frac = fractional_candidates(x, ints, opt);
    if (frac.empty()) {   // new incumbent; Cbc's cutoff = incumbent - cutoff_decr
      if (obj < R.best_obj) { R.best_obj = obj; R.best_x = x; cutoff = obj - opt.cutoff_decr; }
      delete nd;
      continue;
    }
    // candidates needing strong branching: untrusted pseudo-costs (:633-643), at most number_strong_branch
    std::vector<int> cand;
    for (auto &f : frac) {
      const int j = f.second;
      if ((int)cand.size() >= opt.number_strong_branch) break;
      if (std::min(pc_cnt[0][j], pc_cnt[1][j]) < opt.number_before_trust) cand.push_back(j);
    }
    int bcol = -1, bway = 0;
    std::vector<StrongResult> res;
    if (!cand.empty()) {
      chooser.setNodeInfo(nd->depth, R.best_obj < 1e99, opt.time_limit - (now_s() - t0));
      const int rc = chooser.doStrongBranching(cand, obj, res);
      R.sb_batches++; R.sb_children += 2 * (int)cand.size();
      if (rc == -1) { delete nd; continue; }                       // both ways infeasible: node infeasible
      if (rc == 3) { R.status = 2; delete nd; break; }             // time limit inside strong branching (:924-929)
      std::vector<std::pair<int, int>> fix;
      const int bi = chooser.choose(res, obj, cutoff, R.best_obj < 1e99, nd->depth, bway, fix);
      for (auto &r : res)                                           // pseudo-cost update from SB (:994-1026)
        for (int w = 0; w < 2; ++w)
          if (r.status[w] == 0) {
            const double dist = (w == 0) ? r.value - std::floor(r.value) : std::ceil(r.value) - r.value;
            if (dist > 1e-6) { pc_sum[w][r.col] += r.change[w] / dist; pc_cnt[w][r.col] += 1.0; }
          }
      for (auto &f : fix) {   // a side proven infeasible fixes the variable to the other side (:698-704, 722-728)
        const double v = x[f.first];
        if (f.second == 0) nd->xl[f.first] = std::max(nd->xl[f.first], std::ceil(v));
        else nd->xu[f.first] = std::min(nd->xu[f.first], std::floor(v));
      }
      if (bi >= 0) bcol = res[bi].col;
    }
    if (bcol < 0) {   // all candidates trusted: pseudo-cost score with the same max-min rule
      double best = -1e300;
      const double MAXMIN = (R.best_obj < 1e99 && nd->depth > 10) ? chooser.maxmin_crit_have_sol : chooser.maxmin_crit_no_sol;
      for (auto &f : frac) {
        const int j = f.second;
        const double dn = (pc_cnt[0][j] > 0 ? pc_sum[0][j] / pc_cnt[0][j] : 1.0) * (x[j] - std::floor(x[j]));
        const double up = (pc_cnt[1][j] > 0 ? pc_sum[1][j] / pc_cnt[1][j] : 1.0) * (std::ceil(x[j]) - x[j]);
        const double v = MAXMIN * std::min(dn, up) + (1 - MAXMIN) * std::max(dn, up);
        if (v > best) { best = v; bcol = j; bway = up > dn ? 0 : 1; }
      }
    }
    // two children; results of the strong-branching solves are reused when the bounds are exactly those that were solved
    Node *ch[2];
    for (int w = 0; w < 2; ++w) {
      ch[w] = new Node{nd->xl, nd->xu, obj, nd->depth + 1};
      if (w == 0) ch[w]->xu[bcol] = std::floor(x[bcol]); else ch[w]->xl[bcol] = std::ceil(x[bcol]);
      const B200Solver::Child *c = nullptr;
      bool same = true;   // fixings applied after SB change the bounds: then the cached result is no longer this child
      for (int j = 0; j < n && same; ++j)
        if (j != bcol && (nd->xl[j] != root.x_l_[j] || nd->xu[j] != root.x_u_[j])) same = false;
      if (same && !res.empty() && !qp_mode) c = sb.lookup(bcol, w == 1);   // QP children are surrogates, never node results
      // (an unfinished strong-branching solve is not a node result: that child is solved again as a node)
      if (c && (c->status == 1 || c->status == 2 || c->status == TNLPSolver::provenInfeasible)) {
        ch[w]->solved = true; ch[w]->obj = c->obj; ch[w]->status = (int)c->status; ch[w]->x.assign(c->x(), c->x() + n);
        if (c->status == 1 || c->status == 2) ch[w]->bound = std::max(obj, c->obj);
      }
    }
    // follow the preferred way first, keep the other on the heap (probed-dive flavour, BonDiver.cpp:255-316)
    dive = ch[bway];
    heap.push(ch[1 - bway]);
    delete nd;
  }
  while (!heap.empty()) { R.best_bound = std::min(R.best_bound < -1e99 ? 1e100 : R.best_bound, heap.top()->bound); delete heap.top(); heap.pop(); }
  if (dive) delete dive;
  if (R.status == 0) { R.status = R.best_obj < 1e99 ? 0 : 1; R.best_bound = R.best_obj; }
  if (R.failed_nodes && opt.log_level > 0)
    fprintf(stderr, "[bb] %d node NLP(s) stayed unsolved (%s)\n", R.failed_nodes, opt.nlp_failure_behavior ? "fathomed" : "search stopped");
  if (app.providerErrors()) R.status = 3;
  R.ipm_iterations = si.totalIterations() + app.totalBatchIterations() + sb.qpIterations();
  R.seconds = now_s() - t0;
  if (getenv("BNLP_BB_TIMERS"))
    fprintf(stderr, "[bb timers] total %.3f s, prefetchMany %.3f s, inside the batch-solve callback %.3f s\n", R.seconds, t_prefetch,
            app.solveSeconds());
  root.x_l_ = xl_root; root.x_u_ = xu_root;
  return R;
}

// ------------------------------------------------------------------------------------------------ retry policies
static bool is_abandoned(int st) {   // OsiTMINLPInterface::isAbandoned (:2907-2920)
  return st == TNLPSolver::iterationLimit || st == TNLPSolver::computationError || st == TNLPSolver::illDefinedProblem ||
         st == TNLPSolver::illegalOption || st == TNLPSolver::externalException || st == TNLPSolver::exception;
}
static bool is_optimal(int st) { return st == TNLPSolver::solvedOptimal || st == TNLPSolver::solvedOptimalTol; }

void random_starting_point(const TMINLP2TNLP &t, double R, unsigned long long &rng, double *sol) {
  for (int i = 0; i < t.n; ++i) {   // uniform mode (:1768-1777)
    double lower = std::min(-R, t.x_u_[i] - R);
    lower = std::max(t.x_l_[i], lower);
    double upper = std::max(R, t.x_l_[i] + R);
    upper = std::min(t.x_u_[i], upper);
    lower = std::min(upper, lower);
    upper = std::max(upper, lower);
    rng = (rng * 0x5DEECE66DULL + 0xB) & ((1ULL << 48) - 1);
    sol[i] = (double)rng / (double)(1ULL << 48) * (upper - lower) + lower;
  }
}

namespace {
struct StartsBatch {
  int B, n, m;
  std::vector<double> OBJ, X, G, D;
  std::vector<int> ST, IT;
  int run(const TMINLP2TNLP &t, BatchSolveStartsFn &solve, const std::vector<double> &X0) {
    n = t.n; m = t.m; B = (int)(X0.size() / n);
    std::vector<double> XL((size_t)B * n), XU((size_t)B * n);
    for (int b = 0; b < B; ++b) {
      std::copy(t.x_l_.begin(), t.x_l_.end(), &XL[(size_t)b * n]);
      std::copy(t.x_u_.begin(), t.x_u_.end(), &XU[(size_t)b * n]);
    }
    OBJ.assign(B, 0.0); ST.assign(B, 0); IT.assign(B, 0);
    X.assign((size_t)B * n, 0.0); G.assign((size_t)B * std::max(m, 1), 0.0); D.assign((size_t)B * (2 * n + m), 0.0);
    const int rc = solve(B, XL.data(), XU.data(), X0.data(), OBJ.data(), ST.data(), IT.data(), X.data(), G.data(), D.data());
    if (rc) ST.assign(B, (int)TNLPSolver::exception);
    return rc;
  }
  void install(TMINLP2TNLP &t, int b) const {
    t.x_sol_.assign(X.begin() + (size_t)b * n, X.begin() + (size_t)(b + 1) * n);
    t.g_sol_.assign(G.begin() + (size_t)b * m, G.begin() + (size_t)(b + 1) * m);
    t.duals_sol_.assign(D.begin() + (size_t)b * (2 * n + m), D.begin() + (size_t)(b + 1) * (2 * n + m));
    t.obj_value_ = OBJ[b];
    t.return_status = (TNLPSolver::ReturnStatus)ST[b];
  }
};
}  // namespace

TNLPSolver::ReturnStatus resolve_for_robustness(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                                RetryResult &r) {
  const int n = t.n, B = 1 + std::max(0, numsolve);
  std::vector<double> X0((size_t)B * n);
  std::copy(t.x_init_.begin(), t.x_init_.begin() + n, X0.begin());          // resetStartingPoint (:1007)
  unsigned long long rng = 0x5DEECE66DULL ^ o.seed;
  for (int f = 1; f < B; ++f) random_starting_point(t, o.max_random_point_radius, rng, &X0[(size_t)f * n]);
  StartsBatch S;
  S.run(t, solve, X0);
  r = RetryResult();
  r.nlp_solves = B;
  for (int f = 0; f < B; ++f) {                                             // first attempt that is not abandoned (:1024, :1057)
    if (!is_abandoned(S.ST[f])) { r.used = f; S.install(t, f); return (TNLPSolver::ReturnStatus)t.return_status; }
    r.num_failed++;
  }
  S.install(t, B - 1);     // all abandoned: the caller throws newUnsolvedError or pretends infeasible (:1076-1091)
  return (TNLPSolver::ReturnStatus)t.return_status;
}

TNLPSolver::ReturnStatus resolve_for_cost(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                          RetryResult &r) {
  const int n = t.n, B = std::max(0, numsolve);
  r = RetryResult();
  if (B == 0) return (TNLPSolver::ReturnStatus)t.return_status;
  std::vector<double> X0((size_t)B * n);
  unsigned long long rng = 0x5DEECE66DULL ^ o.seed;
  for (int f = 0; f < B; ++f) random_starting_point(t, o.max_random_point_radius, rng, &X0[(size_t)f * n]);
  StartsBatch S;
  S.run(t, solve, X0);
  r.nlp_solves = B;
  double bestBound = is_optimal(t.return_status) ? t.obj_value_ : 1.7976931348623157e308;   // :858
  double sum = 0;
  int nopt = 0;
  for (int f = 0; f < B; ++f) {
    if (is_optimal(S.ST[f]) && S.OBJ[f] < bestBound) { bestBound = S.OBJ[f]; r.used = f + 1; S.install(t, f); }   // :890-901
    if (is_abandoned(S.ST[f])) r.num_failed++;
    else if (S.ST[f] == TNLPSolver::provenInfeasible) r.num_infeasible++;
    if (is_optimal(S.ST[f])) { sum += S.OBJ[f]; nopt++; }
  }
  if (nopt) {
    r.mean = sum / nopt;
    double v = 0;
    for (int f = 0; f < B; ++f) if (is_optimal(S.ST[f])) v += (S.OBJ[f] - r.mean) * (S.OBJ[f] - r.mean);
    r.std_dev = std::sqrt(v / nopt);
  }
  return (TNLPSolver::ReturnStatus)t.return_status;
}

// ------------------------------------------------------------------------------------------------ NLP diving
DiveResult dive_run(TMINLP2TNLP &minlp, BatchSolveFn solve, GradFFn grad_f, const std::vector<int> &column_length,
                    double solutionValue, const DiveOptions &opt) {
  DiveResult R;
  const int n = minlp.n, m = minlp.m;
  const double integerTolerance = opt.integer_tolerance;
  double primalTolerance = 1.0e-6;
  // adjustPrimalTolerance (:335-358)
  for (int i = 0; i < m && i < (int)minlp.g_sol_.size(); ++i) {
    if (minlp.g_sol_[i] < minlp.g_l_[i] - primalTolerance) primalTolerance = minlp.g_l_[i] - minlp.g_sol_[i];
    else if (minlp.g_sol_[i] > minlp.g_u_[i] + primalTolerance) primalTolerance = minlp.g_sol_[i] - minlp.g_u_[i];
  }
  std::vector<double> newSolution = minlp.x_sol_, new_g = minlp.g_sol_;
  std::vector<int> integerColumns;
  int numberFractionalVariables = 0;
  for (int j = 0; j < n; ++j)
    if (minlp.var_types[j] != 0) {
      integerColumns.push_back(j);
      if (std::fabs(std::floor(newSolution[j] + 0.5) - newSolution[j]) > integerTolerance) numberFractionalVariables++;
    }
  const int numberIntegers = (int)integerColumns.size();
  const int maxNumberAtBoundToFix = (int)std::floor(opt.percentage_to_fix * numberIntegers);
  unsigned long long rng = 0x5DEECE66DULL ^ opt.seed;
  auto drand = [&]() { rng = (rng * 0x5DEECE66DULL + 0xB) & ((1ULL << 48) - 1); return (double)rng / (double)(1ULL << 48); };
  std::vector<double> &x_l = minlp.x_l_, &x_u = minlp.x_u_;
  std::vector<double> gradient(n, 0.0);
  bool last_success = true;
  double curObj = minlp.obj_value_;   // objective of newSolution

  auto select = [&](int &bestColumn, int &bestRound) {
    bestColumn = -1; bestRound = -1;
    double best = 1e300;
    if (opt.rule == 1 && grad_f) grad_f(newSolution.data(), gradient.data());
    for (int j : integerColumns) {
      const double value = newSolution[j];
      if (!(std::fabs(std::floor(value + 0.5) - value) > integerTolerance)) continue;
      const double below = std::floor(value), above = std::ceil(value);
      const double downFraction = below >= x_l[j] ? value - below : 1e300;
      const double upFraction = above <= x_u[j] ? above - value : 1e300;
      double score;
      int round;
      if (opt.rule == 0) {   // HeuristicDiveFractional::selectVariableToBranch (:55-108)
        if (downFraction < upFraction) { score = downFraction; round = -1; }
        else if (downFraction > upFraction) { score = upFraction; round = 1; }
        else if (drand() < 0.5) { score = downFraction; round = -1; }
        else { score = upFraction; round = 1; }
      } else {               // HeuristicDiveVectorLength::selectVariableToBranch (:93-159)
        const double g = gradient[j];
        double objdelta;
        if (g >= 0.0 && upFraction < 1e300) { objdelta = g * upFraction; round = 1; }
        else if (g < 0.0 && downFraction < 1e300) { objdelta = g * downFraction; round = -1; }
        else if (upFraction < 1e300) { objdelta = g * upFraction; round = 1; }
        else { objdelta = g * downFraction; round = -1; }
        const int len = j < (int)column_length.size() ? column_length[j] : 0;
        score = (objdelta + 1e-6) / ((double)len + 1.0);
      }
      if (score < best) { best = score; bestColumn = j; bestRound = round; }
    }
  };

  struct Attempt { std::vector<double> xl, xu; };
  while (numberFractionalVariables) {
    R.steps++;
    int bestColumn, bestRound;
    select(bestColumn, bestRound);
    // integer variables sitting at a bound get fixed there, up to percentage_to_fix of them (:146-169)
    std::vector<int> columnFixed;
    std::vector<double> originalBound;
    std::vector<char> fixedAtLowerBound;
    for (int j : integerColumns) {
      const double value = newSolution[j];
      if (std::fabs(std::floor(value + 0.5) - value) <= integerTolerance && (int)columnFixed.size() < maxNumberAtBoundToFix) {
        if (std::fabs(x_l[j] - value) <= integerTolerance && x_l[j] != x_u[j]) {
          columnFixed.push_back(j); originalBound.push_back(x_u[j]); fixedAtLowerBound.push_back(1);
          x_u[j] = x_l[j];
        } else if (std::fabs(x_u[j] - value) <= integerTolerance && x_l[j] != x_u[j]) {
          columnFixed.push_back(j); originalBound.push_back(x_l[j]); fixedAtLowerBound.push_back(0);
          x_l[j] = x_u[j];
        }
        if ((int)columnFixed.size() == maxNumberAtBoundToFix) break;
      }
    }
    if (bestColumn < 0) break;
    const double keepL = x_l[bestColumn], keepU = x_u[bestColumn];
    const double dn = std::floor(newSolution[bestColumn]), up = std::ceil(newSolution[bestColumn]);
    // the ladder of :171-207 as explicit bound sets: 0 = fixings + round, 1 = no fixings + round, 2 = no fixings + flip
    std::vector<Attempt> ladder;
    auto with_round = [&](std::vector<double> xl, std::vector<double> xu, int round) {
      xl[bestColumn] = keepL; xu[bestColumn] = keepU;
      if (round < 0) xu[bestColumn] = dn; else xl[bestColumn] = up;
      return Attempt{std::move(xl), std::move(xu)};
    };
    ladder.push_back(with_round(x_l, x_u, bestRound));
    std::vector<double> nl = x_l, nu = x_u;
    for (size_t i = 0; i < columnFixed.size(); ++i) {
      if (fixedAtLowerBound[i]) nu[columnFixed[i]] = originalBound[i]; else nl[columnFixed[i]] = originalBound[i];
    }
    if (!columnFixed.empty()) ladder.push_back(with_round(nl, nu, bestRound));
    ladder.push_back(with_round(nl, nu, -bestRound));
    const int L = (int)ladder.size();
    std::vector<double> OBJ(L), X((size_t)L * n), G((size_t)L * std::max(m, 1));
    std::vector<int> ST(L, 0), IT(L, 0);
    std::vector<char> have(L, 0);
    auto run = [&](int a0, int cnt) {
      std::vector<double> XL((size_t)cnt * n), XU((size_t)cnt * n);
      for (int a = 0; a < cnt; ++a) {
        std::copy(ladder[a0 + a].xl.begin(), ladder[a0 + a].xl.end(), &XL[(size_t)a * n]);
        std::copy(ladder[a0 + a].xu.begin(), ladder[a0 + a].xu.end(), &XU[(size_t)a * n]);
      }
      const int rc = solve(cnt, XL.data(), XU.data(), minlp.x_init_.data(), &OBJ[a0], &ST[a0], &IT[a0], &X[(size_t)a0 * n],
                           &G[(size_t)a0 * std::max(m, 1)], nullptr);
      R.batches++;
      R.nlp_solves += cnt;
      for (int a = 0; a < cnt; ++a) { have[a0 + a] = 1; R.ipm_iterations += IT[a0 + a]; if (rc) ST[a0 + a] = TNLPSolver::exception; }
    };
    if (opt.speculate) run(0, L);
    int used = -1;
    for (int a = 0; a < L; ++a) {       // replay: first attempt whose status is Ipopt SUCCESS (:173)
      if (!have[a]) run(a, 1);
      if (ST[a] == TNLPSolver::solvedOptimal) { used = a; break; }
    }
    last_success = used >= 0;
    if (used >= 0) { x_l = ladder[used].xl; x_u = ladder[used].xu; }
    else { x_l = ladder[L - 1].xl; x_u = ladder[L - 1].xu; break; }
    newSolution.assign(X.begin() + (size_t)used * n, X.begin() + (size_t)(used + 1) * n);
    if (m) new_g.assign(G.begin() + (size_t)used * m, G.begin() + (size_t)(used + 1) * m);
    curObj = OBJ[used];
    if (curObj >= solutionValue) break;              // eval_f at the new point (:216-219)
    numberFractionalVariables = 0;
    for (int j : integerColumns)
      if (std::fabs(std::floor(newSolution[j] + 0.5) - newSolution[j]) > integerTolerance) numberFractionalVariables++;
  }
  // final checks (:231-279): bounds, integrality, rows within the primal tolerance, better than the incumbent
  bool feasible = last_success;
  for (int j = 0; j < n && feasible; ++j) {
    const double value = newSolution[j];
    if (value < x_l[j] || value > x_u[j]) feasible = false;
    else if (minlp.var_types[j] != 0 && std::fabs(std::floor(value + 0.5) - value) > integerTolerance) feasible = false;
  }
  for (int i = 0; i < m && feasible; ++i)
    if (new_g[i] < minlp.g_l_[i] - primalTolerance || new_g[i] > minlp.g_u_[i] + primalTolerance) feasible = false;
  if (feasible && curObj < solutionValue) { R.found = true; R.obj = curObj; R.x = newSolution; }
  return R;
}

}  // namespace Bonmin
