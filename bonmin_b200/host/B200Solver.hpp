// Host adapters that put the batched engine behind Bonmin's own plugin interfaces.
//   B200Solver                     : TNLPSolver            (replaces IpoptSolver, src/Interfaces/Ipopt/BonIpoptSolver.cpp)
//   BatchedStrongBranchingSolver   : StrongBranchingSolver (new value of `variable_selection`, BonBonminSetup.cpp:434-469)
//   BatchedChooseVariable          : the doStrongBranching override of SURVEY.md Appendix C
//                                    (src/Algorithms/Branching/BonChooseVariable.cpp:787-957 replayed against a prefetch)
//   bb_run                         : a small host B-BB driver (Cbc is not available here) restating Bonmin's B-BB
//                                    node loop closely enough to report "B-BB wall time to optimum" (SURVEY.md §7 step 2).
// The arithmetic provider is abstracted as a batch-solve callback so the same host logic runs on the CUDA engine
// (bnlp_solve_batch) in the product and on the CPU oracle in tests — the adapters never compute anything themselves.
#pragma once
#include <functional>
#include <map>
#include <utility>

#include <memory>

#include "bonmin_shim.hpp"

namespace Bonmin {

// B x OptimizeTNLP: same argument meaning as bnlp_solve_batch (include/bnlp.h)
using BatchSolveFn = std::function<int(int B, const double *XL, const double *XU, const double *X0, double *OBJ,
                                       int *STATUS, int *ITERS, double *X, double *G, double *DUALS)>;

// B x QpBranchingSolver::solveFromHotStart (BonQpBranchingSolver.cpp:96-116): the node's optimum (x_sol, duals_sol) and
// the children's bounds in the original variables; same meaning as bnlp_tqp_solve_batch (include/bnlp.h)
using TqpSolveFn = std::function<int(const double *x_sol, const double *duals_sol, int B, const double *XL,
                                     const double *XU, double *OBJ, int *STATUS, int *ITERS, double *X)>;

// The strong-branching form (same meaning as bnlp_solve_children, include/bnlp.h): child b = parent cp[b] of P with one
// bound of column cc[b] replaced by cb[b] (cw[b] 0: upper bound, 1: lower bound); the bound vectors are expanded by the
// provider (on the device in the product)
using ChildrenSolveFn = std::function<int(int P, const double *PXL, const double *PXU, int B, const int *cp, const int *cc,
                                          const int *cw, const double *cb, const double *X0, double *OBJ, int *STATUS,
                                          int *ITERS, double *X)>;

// B x ReOptimizeTNLP from a primal AND dual start: same argument meaning as bnlp_solve_batch_warm (include/bnlp.h),
// DUAL0 = one vector [z_L | z_U | lambda] of 2n+m shared by the B children
using BatchSolveWarmFn = std::function<int(int B, const double *XL, const double *XU, const double *X0, const double *DUAL0,
                                           double *OBJ, int *STATUS, int *ITERS, double *X, double *G, double *DUALS)>;

class B200Solver : public TNLPSolver {
 public:
  explicit B200Solver(BatchSolveFn fn, BatchSolveWarmFn warm = nullptr) : solve_(std::move(fn)), solve_warm_(std::move(warm)) {}
  Ipopt::SmartPtr<TNLPSolver> clone() override {
    B200Solver *c = new B200Solver(solve_, solve_warm_);
    c->children_ = children_; c->options_ = options_; c->enable_warm_start_ = enable_warm_start_;
    c->log_level_ = log_level_; c->default_log_level_ = default_log_level_;
    return Ipopt::SmartPtr<TNLPSolver>(c);
  }
  // the TNLP is driven through its virtuals only: get_nlp_info, get_bounds_info, get_starting_point, finalize_solution
  ReturnStatus OptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) override;
  ReturnStatus ReOptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) override;
  double CPUTime() override { return cpu_time_; }
  int IterationCount() override { return iterations_; }
  std::string &solverName() override { return name_; }
  // --- the rest of the TNLPSolver contract (BonTNLPSolver.hpp:95-205)
  // options: "name value" lines as in bonmin.opt / ipopt.opt; kept for inspection (tol and max_iter are fixed when the
  // engine behind the BatchSolveFn is created, see bnlp_options)
  bool Initialize(std::string params_file) override;
  bool Initialize(std::istream &is) override;
  const std::string *option(const std::string &name) const;
  // warm starts exactly as IpoptSolver handles them (Ipopt/BonIpoptSolver.cpp:345-430): setWarmStart puts the primal point
  // and the multipliers [z_L | z_U | lambda] into the TMINLP2TNLP (setxInit / setDualsInit) and switches the warm-start
  // option on; the next (Re)OptimizeTNLP then asks get_starting_point for x, z_L, z_U and lambda and solves from them
  // (bnlp_solve_batch_warm).  Without a BatchSolveWarmFn only the primal point is used.
  bool setWarmStart(const CoinWarmStart *warm, Ipopt::SmartPtr<TMINLP2TNLP> tnlp) override;
  CoinWarmStart *getUsedWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const override;
  CoinWarmStart *getWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const override;
  CoinWarmStart *getEmptyWarmStart() const override { return new IpoptWarmStart(true); }
  bool warmStartIsValid(const CoinWarmStart *ws) const override;
  void enableWarmStart() override { enable_warm_start_ = true; }
  void disableWarmStart() override { enable_warm_start_ = false; }
  bool warmStartEnabled() const { return enable_warm_start_; }
  bool lastSolveWasWarmStarted() const { return warm_started_last_; }
  void setOutputToDefault() override { log_level_ = default_log_level_; }
  void forceSolverOutput(int log_level) override { log_level_ = log_level; }
  int logLevel() const { return log_level_; }
  int errorCode() const override { return error_code_; }      // the provider's return code of the last failed call
  class UnsolvedB200Error : public UnsolvedError {
   public:
    UnsolvedB200Error(int num, Ipopt::SmartPtr<TMINLP2TNLP> problem, std::string name) : UnsolvedError(num, problem, std::move(name)) {}
    const std::string &errorName() const override;
    const std::string &solverName() const override;
  };
  UnsolvedError *newUnsolvedError(int num, Ipopt::SmartPtr<TMINLP2TNLP> problem, std::string name) override {
    return new UnsolvedB200Error(num, problem, std::move(name));
  }
  // the batched sibling: children differ only in bounds; results in TMINLP2TNLP layout
  // (the primal points of one batch live in ONE shared block; a child refers to its row)
  struct Child {
    double obj; ReturnStatus status; int iters;
    std::shared_ptr<std::vector<double>> block; size_t off = 0;
    const double *x() const { return block->data() + off; }
  };
  int OptimizeBatch(const TMINLP2TNLP &parent, const std::vector<double> &XL, const std::vector<double> &XU, int B,
                    std::vector<Child> &out);
  // children of P parents, each one bound away from its parent (see ChildrenSolveFn); without a ChildrenSolveFn the
  // bound vectors are expanded here and go through the BatchSolveFn
  void setChildrenSolver(ChildrenSolveFn f) { children_ = std::move(f); }
  int OptimizeChildren(const TMINLP2TNLP &parent, int P, const std::vector<double> &PXL, const std::vector<double> &PXU,
                       const std::vector<int> &cp, const std::vector<int> &cc, const std::vector<int> &cw,
                       const std::vector<double> &cb, std::vector<Child> &out);
  long totalBatchIterations() const { return total_batch_iterations_; }
  double solveSeconds() const { return solve_seconds_; }
  int providerErrors() const { return provider_errors_; }   // batch calls whose provider returned non-zero

 private:
  int unpack(int rc, int B, int n, std::vector<Child> &out);
  std::shared_ptr<std::vector<double>> x_block(size_t cnt);
  ReturnStatus solve_tnlp(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp, bool warm);
  BatchSolveFn solve_;
  BatchSolveWarmFn solve_warm_;
  ChildrenSolveFn children_;
  bool optimized_before_ = false, warm_started_last_ = false;
  std::vector<double> obj_buf_, xl_buf_, xu_buf_;   // reused between calls: no fresh pages per round
  std::vector<int> st_buf_, it_buf_;
  std::shared_ptr<std::vector<double>> x_pool_;
  std::string name_ = "B200 batched IPM";
  double cpu_time_ = 0;
  int iterations_ = 0;
  long total_batch_iterations_ = 0;
  double solve_seconds_ = 0;
  int provider_errors_ = 0;
  int error_code_ = 0;
  bool enable_warm_start_ = false;
  int log_level_ = 0, default_log_level_ = 0;
  std::vector<std::pair<std::string, std::string>> options_;
};

// Strong-branching plugin: prefetch() solves all 2K children of a node as ONE batch; solveFromHotStart() then serves
// the child selected by diffing the interface's current bounds against the hot-start snapshot.
class BatchedStrongBranchingSolver : public StrongBranchingSolver {
 public:
  explicit BatchedStrongBranchingSolver(B200Solver *app) : app_(app) {}
  // QpBranchingSolver flavour (BonQpBranchingSolver.cpp:50-116): children are solved on the BranchingTQP surrogate built
  // at the node optimum that markHotStart sees, instead of as full NLPs
  void setQpSurrogate(TqpSolveFn f) { qp_ = std::move(f); }
  bool isQp() const { return (bool)qp_; }
  void markHotStart(OsiTMINLPInterface *si) override;
  TNLPSolver::ReturnStatus solveFromHotStart(OsiTMINLPInterface *si) override;
  void unmarkHotStart(OsiTMINLPInterface *si) override;
  // candidates: (column, down bound = floor(x), up bound = ceil(x)); called between markHotStart and the replay loop
  void prefetch(OsiTMINLPInterface *si, const std::vector<int> &cols, const std::vector<double> &down_ub,
                const std::vector<double> &up_lb);
  // Many nodes at once: the strong-branching children of several open nodes go into ONE batch (they share the start
  // point, BonCbc.cpp:492-505); useNode(i) then makes node i's results the active cache for the replay loop.
  struct NodeRequest { std::vector<double> xl, xu; std::vector<int> cols; std::vector<double> down_ub, up_lb; };
  void prefetchMany(OsiTMINLPInterface *si, const std::vector<NodeRequest> &reqs);
  void useNode(int i);
  bool preloaded() const { return preloaded_; }
  int cacheHits() const { return hits_; }
  int cacheMisses() const { return misses_; }
  long qpIterations() const { return qp_iterations_; }
  const B200Solver::Child *lookup(int col, bool up) const;

 private:
  B200Solver *app_;
  TqpSolveFn qp_;
  std::vector<double> xsol0_, duals0_;   // node optimum at markHotStart (what BranchingTQP's constructor reads)
  std::vector<double> xl0_, xu0_;
  std::map<std::pair<int, int>, B200Solver::Child> cache_;   // (column, way) -> result
  std::vector<std::map<std::pair<int, int>, B200Solver::Child>> multi_;
  std::vector<std::pair<std::vector<double>, std::vector<double>>> multi_bounds_;
  bool preloaded_ = false;
  int hits_ = 0, misses_ = 0;
  long qp_iterations_ = 0;
};

struct StrongResult {  // what HotInfo holds per candidate (BonChooseVariable.hpp:109-113: -1 not done, 0 feasible, 1 infeasible, 2 not finished)
  int col;
  double value;
  int status[2];     // [down, up]
  double change[2];
  double obj[2];
};

// OsiBranchingInformation (Osi/OsiBranchingObject.hpp): the fields BonChooseVariable::doStrongBranching reads
struct OsiBranchingInformation {
  const double *lower_ = nullptr, *upper_ = nullptr, *solution_ = nullptr;   // the node's bounds and LP/NLP point
  double objectiveValue_ = 0.0, cutoff_ = 1e100, timeRemaining_ = 1e100;
  int depth_ = 0;
  bool hasSolution_ = false;
};

class BatchedChooseVariable {
 public:
  BatchedChooseVariable(OsiTMINLPInterface *si, BatchedStrongBranchingSolver *sb) : si_(si), sb_(sb) {}
  double maxmin_crit_no_sol = 0.7, maxmin_crit_have_sol = 0.1;   // BonChooseVariable.cpp:170-175
  double integer_tolerance = 1e-6;                               // BonBabSetupBase.cpp:478-482
  // doStrongBranching with the reference's signature and return codes (BonChooseVariable.cpp:787-957): -1 some candidate is
  // infeasible both ways (node infeasible), 1 some candidate is infeasible one way (fix it), 3 time limit, 0 otherwise;
  // results() is cut to the candidates evaluated (look-ahead, time limit).  The candidates are those of setCandidates().
  int doStrongBranching(OsiTMINLPInterface *solver, OsiBranchingInformation *info, int numberToDo, int returnCriterion);
  void setCandidates(const std::vector<int> &cols);              // setupList's strong-branching list (HotInfo per candidate)
  const std::vector<StrongResult> &results() const { return results_; }
  void setNumberLookAhead(int k) { numberLookAhead_ = k; }       // `number_look_ahead` (:152-158), default 0 = off
  void setTimeLimit(double start_time, double time_limit) { start_time_ = start_time; time_limit_ = time_limit; }
  void setNodeInfo(int depth, bool have_solution, double time_remaining = 1e100) {
    depth_ = depth; have_solution_ = have_solution; time_remaining_ = time_remaining;
  }
  int numberStrongDone() const { return numberStrongDone_; }
  int numberStrongFixed() const { return numberStrongFixed_; }
  // the column-list form the host B-BB driver uses (every candidate evaluated unless look-ahead / time stops the loop)
  int doStrongBranching(const std::vector<int> &cols, double parent_obj, std::vector<StrongResult> &results);
  // scoring of chooseVariable (BonChooseVariable.cpp:682-747); returns the index into results (or -1) and the way
  int choose(const std::vector<StrongResult> &results, double parent_obj, double cutoff, bool have_solution, int depth,
             int &bestWhichWay, std::vector<std::pair<int, int>> &fixings) const;

 private:
  OsiTMINLPInterface *si_;
  BatchedStrongBranchingSolver *sb_;
  std::vector<StrongResult> results_;
  int numberLookAhead_ = 0, numberStrongDone_ = 0, numberStrongFixed_ = 0;
  int depth_ = 0;
  bool have_solution_ = false;
  double time_limit_ = 1e100, start_time_ = 0.0, time_remaining_ = 1e100;
};

struct BbOptions {
  int number_strong_branch = 20;        // BonBabSetupBase.cpp:547-550
  int number_before_trust = 8;          // :554-559
  double integer_tolerance = 1e-6;      // :478-482
  double cutoff_decr = 1e-5;            // :507-513
  double allowable_gap = 0.0, allowable_fraction_gap = 0.0;   // :484-497
  int node_limit = 1000000;
  int parallel_nodes = 1;               // > 1: that many best-bound nodes are expanded per round with ONE batched launch
  int variable_selection = 0;           // 0 NLP strong branching, 1 QP strong branching (needs a TqpSolveFn)
  double time_limit = 1e10;
  int log_level = 0;
  std::vector<int> priorities;          // empty, or n values: only the fractional integers of LOWEST value are candidates
  // nlp_failure_behavior (BonOsiTMINLPInterface.cpp:233-240): 0 = stop (the reference throws newUnsolvedError when a node
  // NLP stays unsolved, :2639-2644), 1 = fathom (pretendFailIsInfeasible_: the node is dropped with a warning)
  int nlp_failure_behavior = 0;
  int number_look_ahead = 0;            // BonChooseVariable.cpp:152-158: stop strong branching after that many non-improving candidates (0 = off)
  std::vector<double> incumbent_x;      // empty, or an integer-feasible point found before the search (heuristics) ...
  double incumbent_obj = 1e100;         // ... and its objective: the search starts with cutoff = incumbent_obj - cutoff_decr
};
struct BbResult {
  int status = 0;               // 0 optimal, 1 infeasible, 2 node/time limit, 3 provider error, 4 a node NLP stayed unsolved
  int failed_nodes = 0;         // node NLPs that ended in an abandoned status even after the re-solve
  double best_obj = 1e100, best_bound = -1e100;
  std::vector<double> best_x;
  int nodes = 0, nlp_solves = 0, sb_batches = 0, sb_children = 0, served_from_sb_cache = 0;
  long ipm_iterations = 0;
  double seconds = 0;
};
// ---- retry policies of OsiTMINLPInterface around one NLP (BonOsiTMINLPInterface.cpp:836-1093).  The reference solves
// the restarts one after the other; they differ only in their start point, so here they are ONE batch (per-child X0)
// and the reference's sequential rule is replayed over the results in order.
using BatchSolveStartsFn = std::function<int(int B, const double *XL, const double *XU, const double *X0 /* B x n */,
                                             double *OBJ, int *STATUS, int *ITERS, double *X, double *G, double *DUALS)>;
struct RetryOptions {
  double max_random_point_radius = 1e5;   // BonOsiTMINLPInterface.cpp:128-134
  unsigned seed = 1;                      // CoinDrand48 in the reference
};
struct RetryResult {
  int used = -1;               // index of the start whose result was kept (0 = the user's start point), -1 = none
  int num_failed = 0, num_infeasible = 0, nlp_solves = 0;
  double mean = 0, std_dev = 0;   // of the optimal objectives (resolveForCost's statistics, :942-961)
};
// randomStartingPoint, uniform mode (:1749-1795)
void random_starting_point(const TMINLP2TNLP &t, double radius, unsigned long long &rng, double *sol);
// resolveForRobustness(numsolve) (:989-1093): the user's start point again, then numsolve random ones; the first that is
// not abandoned wins.  Fills t's solution fields; returns the status (of the last attempt if all were abandoned).
TNLPSolver::ReturnStatus resolve_for_robustness(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                                RetryResult &r);
// resolveForCost(numsolve) (:836-986): numsolve random starts; the best proven-optimal objective replaces the current
// solution if it improves on it.
TNLPSolver::ReturnStatus resolve_for_cost(TMINLP2TNLP &t, BatchSolveStartsFn solve, int numsolve, const RetryOptions &o,
                                          RetryResult &r);

// ---- NLP diving heuristics (src/CbcBonmin/Heuristics/BonHeuristicDive.cpp:67-302): a chain of NLPs with progressive
// rounding.  Each step's retry ladder (with at-bound fixings -> without them -> other rounding direction, :171-207) is up
// to three NLPs that differ only in bounds: with speculate = true they are ONE batch and the ladder is replayed against it.
using GradFFn = std::function<int(const double *x, double *grad_f)>;   // TMINLP::eval_grad_f at one point
struct DiveOptions {
  int rule = 0;                       // 0 = HeuristicDiveFractional (:55-108), 1 = HeuristicDiveVectorLength (:93-159)
  double integer_tolerance = 1e-6;    // CbcModel::CbcIntegerTolerance
  double percentage_to_fix = 0.2;     // BonHeuristicDive.cpp:32
  bool speculate = true;
  unsigned seed = 1;                  // tie-break of the fractional rule (CoinDrand48 in the reference)
};
struct DiveResult {
  bool found = false;
  double obj = 1e100;
  std::vector<double> x;
  int steps = 0, nlp_solves = 0, batches = 0;
  long ipm_iterations = 0;
};
// node = the relaxation the dive starts from (x_sol, g_sol, bounds; modified like the reference modifies its clone);
// column_length = number of rows each column appears in (Jacobian structure), only for rule 1; solution_value = incumbent
DiveResult dive_run(TMINLP2TNLP &node, BatchSolveFn solve, GradFFn grad_f, const std::vector<int> &column_length,
                    double solution_value, const DiveOptions &opt);

BbResult bb_run(TMINLP2TNLP &root, BatchSolveFn solve, const BbOptions &opt, TqpSolveFn tqp = nullptr,
                ChildrenSolveFn children = nullptr);

}  // namespace Bonmin
