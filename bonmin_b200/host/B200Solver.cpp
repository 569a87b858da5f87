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

// Contract of OptimizeTNLP (BonIpoptSolver.cpp:134-176, Filter/BonFilterSolver.cpp:574-645): read sizes, bounds and
// the start point from the TNLP, solve, and finish by filling x_sol / g_sol / duals_sol / obj_value exactly as
// TMINLP2TNLP::finalize_solution would (BonTMINLP2TNLP.cpp:474-509).  An engine error maps to TNLPSolver::exception.
TNLPSolver::ReturnStatus B200Solver::OptimizeTNLP(TMINLP2TNLP &t) {
  const double t0 = now_s();
  const int n = t.n, m = t.m;
  double obj = 0;
  int st = 0, it = 0;
  t.x_sol.assign(n, 0.0); t.g_sol.assign(m, 0.0); t.duals_sol.assign(2 * n + m, 0.0);
  const int rc = solve_(1, t.x_l.data(), t.x_u.data(), t.x_init.data(), &obj, &st, &it, t.x_sol.data(), t.g_sol.data(),
                        t.duals_sol.data());
  cpu_time_ = now_s() - t0;
  iterations_ = it;
  if (rc != 0) { provider_errors_++; error_code_ = rc; t.return_status = exception; return exception; }
  t.obj_value = obj;
  t.return_status = (ReturnStatus)st;
  return (ReturnStatus)st;
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

bool B200Solver::setWarmStart(const CoinWarmStart *ws, TMINLP2TNLP &tnlp) {   // BonIpoptSolver.cpp:369-401
  if (!ws) return false;
  if (ws->empty()) { disableWarmStart(); return true; }
  if (ws->dualSize() > 0) { tnlp.setDualsInit(ws->dualSize(), ws->dual()); enableWarmStart(); }
  else disableWarmStart();
  if (ws->primalSize() != tnlp.n) return false;
  tnlp.setxInit(ws->primalSize(), ws->primal());
  return true;
}

CoinWarmStart *B200Solver::getUsedWarmStart(TMINLP2TNLP &tnlp) const {        // :347-356
  if ((int)tnlp.x_init.size() < tnlp.n || tnlp.duals_init.empty()) return nullptr;
  return new CoinWarmStart(tnlp.n, 2 * tnlp.n + tnlp.m, tnlp.x_init.data(), tnlp.duals_init.data());
}

CoinWarmStart *B200Solver::getWarmStart(TMINLP2TNLP &tnlp) const {            // :357-366: the last solution and its multipliers
  if ((int)tnlp.x_sol.size() != tnlp.n) return new CoinWarmStart();
  const bool have_duals = (int)tnlp.duals_sol.size() == 2 * tnlp.n + tnlp.m;
  return new CoinWarmStart(tnlp.n, have_duals ? 2 * tnlp.n + tnlp.m : 0, tnlp.x_sol.data(), have_duals ? tnlp.duals_sol.data() : nullptr);
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
  const int rc = solve_(B, XL.data(), XU.data(), p.x_init.data(), obj_buf_.data(), st_buf_.data(), it_buf_.data(), X, nullptr, nullptr);
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
    rc = children_(P, PXL.data(), PXU.data(), B, cp.data(), cc.data(), cw.data(), cb.data(), p.x_init.data(), obj_buf_.data(),
                   st_buf_.data(), it_buf_.data(), X);
  } else {
    if (xl_buf_.size() < (size_t)B * n) { xl_buf_.resize((size_t)B * n); xu_buf_.resize((size_t)B * n); }
    for (int b = 0; b < B; ++b) {
      std::copy(&PXL[(size_t)cp[b] * n], &PXL[(size_t)cp[b] * n] + n, &xl_buf_[(size_t)b * n]);
      std::copy(&PXU[(size_t)cp[b] * n], &PXU[(size_t)cp[b] * n] + n, &xu_buf_[(size_t)b * n]);
      (cw[b] == 0 ? xu_buf_ : xl_buf_)[(size_t)b * n + cc[b]] = cb[b];
    }
    rc = solve_(B, xl_buf_.data(), xu_buf_.data(), p.x_init.data(), obj_buf_.data(), st_buf_.data(), it_buf_.data(), X, nullptr, nullptr);
  }
  solve_seconds_ += now_s() - ts0;
  return unpack(rc, B, n, out);
}

void BatchedStrongBranchingSolver::markHotStart(OsiTMINLPInterface *si) {
  const int n = si->getNumCols();
  xl0_.assign(si->getColLower(), si->getColLower() + n);
  xu0_.assign(si->getColUpper(), si->getColUpper() + n);
  if (qp_) {   // BranchingTQP's constructor: freeze the node optimum and its multipliers (BonBranchingTQP.cpp:41-43)
    xsol0_ = si->problem()->x_sol;
    duals0_ = si->problem()->duals_sol;
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
    return si->solver()->OptimizeTNLP(*si->problem());
  }
  hits_++;
  si->problem()->set_obj_value(c->obj);
  if (c->status == TNLPSolver::solvedOptimal || c->status == TNLPSolver::solvedOptimalTol) si->problem()->Set_x_sol(n, c->x());
  si->problem()->return_status = c->status;
  return c->status;
}

void BatchedStrongBranchingSolver::unmarkHotStart(OsiTMINLPInterface *) { preloaded_ = false; }

int BatchedChooseVariable::doStrongBranching(const std::vector<int> &cols, double parent_obj,
                                             std::vector<StrongResult> &results) {
  const int K = (int)cols.size();
  std::vector<double> xsol(si_->getColSolution(), si_->getColSolution() + si_->getNumCols());
  std::vector<double> dn(K), up(K);
  for (int i = 0; i < K; ++i) { dn[i] = std::floor(xsol[cols[i]]); up[i] = std::ceil(xsol[cols[i]]); }
  si_->markHotStart();                                   // :799
  if (!sb_->preloaded()) sb_->prefetch(si_, cols, dn, up);   // ONE batch of 2K children (Appendix C)
  results.assign(K, StrongResult());
  int returnCode = 0;
  for (int iDo = 0; iDo < K; ++iDo) {                    // :807-948, unchanged control flow
    StrongResult &r = results[iDo];
    r.col = cols[iDo]; r.value = xsol[cols[iDo]];
    for (int w = 0; w < 2; ++w) {
      const double saveL = si_->getColLower()[r.col], saveU = si_->getColUpper()[r.col];
      if (w == 0) si_->setColUpper(r.col, dn[iDo]); else si_->setColLower(r.col, up[iDo]);   // branch->branch(solver) :820
      si_->solveFromHotStart();                          // :822
      // HotInfo::updateInformation (:1118-1145) + OsiHotInfo: 0 feasible, 1 infeasible, 2 not finished
      int status;
      const double obj = si_->getObjValue();
      if (si_->isProvenOptimal()) status = 0;
      else if (si_->isProvenPrimalInfeasible()) status = (std::fabs(obj) < 1e-6) ? 2 : 1;
      else status = 2;
      r.status[w] = status;
      r.obj[w] = obj;
      r.change[w] = (status == 0) ? std::max(0.0, obj - parent_obj) : 0.0;
      si_->setColLower(r.col, saveL); si_->setColUpper(r.col, saveU);   // restore bounds :853-858
    }
    if (r.status[0] == 1 && r.status[1] == 1) returnCode = -1;          // :916-919 (the loop continues)
  }
  si_->unmarkHotStart();                                 // :955
  return returnCode;
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
  B200Solver app(solve);
  if (children) app.setChildrenSolver(children);
  OsiTMINLPInterface si(&root, &app);
  BatchedStrongBranchingSolver sb(&app);
  si.SetStrongBrachingSolver(&sb);
  const bool qp_mode = opt.variable_selection == 1 && tqp;
  if (qp_mode) sb.setQpSurrogate(tqp);
  BatchedChooseVariable chooser(&si, &sb);
  chooser.integer_tolerance = opt.integer_tolerance;
  std::vector<int> ints;
  for (int j = 0; j < n; ++j) if (root.var_types[j] != 0) ints.push_back(j);
  std::vector<double> pc_sum[2], pc_cnt[2];   // pseudo-costs per column and direction
  for (int w = 0; w < 2; ++w) { pc_sum[w].assign(n, 0.0); pc_cnt[w].assign(n, 0.0); }

  // root relaxation (Bab::branchAndBound: model_.initialSolve, then the root primal becomes the start point of every
  // later NLP: src/CbcBonmin/BonCbc.cpp:452,492-505)
  si.initialSolve();
  R.nlp_solves++;
  if (!si.isProvenOptimal()) { R.status = app.providerErrors() ? 3 : si.isProvenPrimalInfeasible() ? 1 : 4; R.failed_nodes += !app.providerErrors() && !si.isProvenPrimalInfeasible(); R.seconds = now_s() - t0; return R; }
  root.setxInit(n, root.x_sol.data());
  std::priority_queue<Node *, std::vector<Node *>, Cmp> heap;
  Node *first = new Node{root.x_l, root.x_u, root.obj_value, 0};
  first->solved = true; first->obj = root.obj_value; first->status = (int)root.return_status; first->x = root.x_sol;
  Node *dive = first;   // node processed next without going through the heap (depth-first plunge on the better child)
  double cutoff = 1e100;
  const std::vector<double> xl_root = root.x_l, xu_root = root.x_u;

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
        root.x_l = nd->xl; root.x_u = nd->xu;
        root.obj_value = nd->obj; root.x_sol = nd->x; root.return_status = (TNLPSolver::ReturnStatus)nd->status;
        si.optimizationStatus_ = (TNLPSolver::ReturnStatus)nd->status;
        sb.useNode((int)q);
        std::vector<StrongResult> res;
        const int rc = chooser.doStrongBranching(reqs[q].cols, nd->obj, res);
        if (rc == -1) { delete nd; continue; }
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
    root.x_l = nd->xl; root.x_u = nd->xu;
    if (nd->solved) {
      R.served_from_sb_cache += (nd != first);
      root.obj_value = nd->obj; root.x_sol = nd->x; root.return_status = (TNLPSolver::ReturnStatus)nd->status;
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
    const double obj = root.obj_value;
    if (obj >= cutoff) { delete nd; continue; }
    const std::vector<double> x = root.x_sol;
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
      const int rc = chooser.doStrongBranching(cand, obj, res);
      R.sb_batches++; R.sb_children += 2 * (int)cand.size();
      if (rc == -1) { delete nd; continue; }                       // both ways infeasible: node infeasible
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
        if (j != bcol && (nd->xl[j] != root.x_l[j] || nd->xu[j] != root.x_u[j])) same = false;
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
  root.x_l = xl_root; root.x_u = xu_root;
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
    double lower = std::min(-R, t.x_u[i] - R);
    lower = std::max(t.x_l[i], lower);
    double upper = std::max(R, t.x_l[i] + R);
    upper = std::min(t.x_u[i], upper);
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
      std::copy(t.x_l.begin(), t.x_l.end(), &XL[(size_t)b * n]);
      std::copy(t.x_u.begin(), t.x_u.end(), &XU[(size_t)b * n]);
    }
    OBJ.assign(B, 0.0); ST.assign(B, 0); IT.assign(B, 0);
    X.assign((size_t)B * n, 0.0); G.assign((size_t)B * std::max(m, 1), 0.0); D.assign((size_t)B * (2 * n + m), 0.0);
    const int rc = solve(B, XL.data(), XU.data(), X0.data(), OBJ.data(), ST.data(), IT.data(), X.data(), G.data(), D.data());
    if (rc) ST.assign(B, (int)TNLPSolver::exception);
    return rc;
  }
  void install(TMINLP2TNLP &t, int b) const {
    t.x_sol.assign(X.begin() + (size_t)b * n, X.begin() + (size_t)(b + 1) * n);
    t.g_sol.assign(G.begin() + (size_t)b * m, G.begin() + (size_t)(b + 1) * m);
    t.duals_sol.assign(D.begin() + (size_t)b * (2 * n + m), D.begin() + (size_t)(b + 1) * (2 * n + m));
    t.obj_value = OBJ[b];
    t.return_status = (TNLPSolver::ReturnStatus)ST[b];
  }
};
}  // namespace

TNLPSolver::ReturnStatus resolve_for_robustness(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                                RetryResult &r) {
  const int n = t.n, B = 1 + std::max(0, numsolve);
  std::vector<double> X0((size_t)B * n);
  std::copy(t.x_init.begin(), t.x_init.begin() + n, X0.begin());          // resetStartingPoint (:1007)
  unsigned long long rng = 0x5DEECE66DULL ^ o.seed;
  for (int f = 1; f < B; ++f) random_starting_point(t, o.max_random_point_radius, rng, &X0[(size_t)f * n]);
  StartsBatch S;
  S.run(t, solve, X0);
  r = RetryResult();
  r.nlp_solves = B;
  for (int f = 0; f < B; ++f) {                                             // first attempt that is not abandoned (:1024, :1057)
    if (!is_abandoned(S.ST[f])) { r.used = f; S.install(t, f); return t.return_status; }
    r.num_failed++;
  }
  S.install(t, B - 1);     // all abandoned: the caller throws newUnsolvedError or pretends infeasible (:1076-1091)
  return t.return_status;
}

TNLPSolver::ReturnStatus resolve_for_cost(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                          RetryResult &r) {
  const int n = t.n, B = std::max(0, numsolve);
  r = RetryResult();
  if (B == 0) return t.return_status;
  std::vector<double> X0((size_t)B * n);
  unsigned long long rng = 0x5DEECE66DULL ^ o.seed;
  for (int f = 0; f < B; ++f) random_starting_point(t, o.max_random_point_radius, rng, &X0[(size_t)f * n]);
  StartsBatch S;
  S.run(t, solve, X0);
  r.nlp_solves = B;
  double bestBound = is_optimal(t.return_status) ? t.obj_value : 1.7976931348623157e308;   // :858
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
  return t.return_status;
}

// ------------------------------------------------------------------------------------------------ NLP diving
DiveResult dive_run(TMINLP2TNLP &minlp, BatchSolveFn solve, GradFFn grad_f, const std::vector<int> &column_length,
                    double solutionValue, const DiveOptions &opt) {
  DiveResult R;
  const int n = minlp.n, m = minlp.m;
  const double integerTolerance = opt.integer_tolerance;
  double primalTolerance = 1.0e-6;
  // adjustPrimalTolerance (:335-358)
  for (int i = 0; i < m && i < (int)minlp.g_sol.size(); ++i) {
    if (minlp.g_sol[i] < minlp.g_l[i] - primalTolerance) primalTolerance = minlp.g_l[i] - minlp.g_sol[i];
    else if (minlp.g_sol[i] > minlp.g_u[i] + primalTolerance) primalTolerance = minlp.g_sol[i] - minlp.g_u[i];
  }
  std::vector<double> newSolution = minlp.x_sol, new_g = minlp.g_sol;
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
  std::vector<double> &x_l = minlp.x_l, &x_u = minlp.x_u;
  std::vector<double> gradient(n, 0.0);
  bool last_success = true;
  double curObj = minlp.obj_value;   // objective of newSolution

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
      const int rc = solve(cnt, XL.data(), XU.data(), minlp.x_init.data(), &OBJ[a0], &ST[a0], &IT[a0], &X[(size_t)a0 * n],
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
    if (new_g[i] < minlp.g_l[i] - primalTolerance || new_g[i] > minlp.g_u[i] + primalTolerance) feasible = false;
  if (feasible && curObj < solutionValue) { R.found = true; R.obj = curObj; R.x = newSolution; }
  return R;
}

}  // namespace Bonmin
