// B-OA on the host side of the boundary, in C++: the outer-approximation decomposition loop of
// OACutGenerator2::performOa (src/Algorithms/OaGenerators/BonOACutGenerator2.cpp:121-268; helpers fixIntegers /
// isDifferentOnIntegers / installCuts in BonSolverHelp.cpp:49-174, cutoff handling post_nlp_solve in
// BonOaDecBase.cpp:237-400) around the batched engine.  The reference holds the master MILP behind a SubMipSolver
// (src/Algorithms/OaGenerators/BonSubMipSolver.hpp: optimize(cutoff, loglevel, maxTime), getLastSolution(), lowBound())
// built over an OsiSolverInterface; Cbc / Osi are absent from this image, so the master is the one thing that stays a
// callback: MasterMipFn receives the whole LP (initial relaxation + every cut installed so far, CSR by rows) and the
// cutoff, and returns a point and a bound.  Everything else — the NLP relaxation, every NLP(y), the initial LP and the
// linearisations at SEVERAL points per pass — is served by the arithmetic providers below, which in the product are
// bnlp_solve_batch, bnlp_extract_linear_relaxation and bnlp_oa_batch on the CUDA engine (bnlp_oa_run_engine).
#pragma once
#include <functional>
#include <vector>

#include "B200Linearizer.hpp"
#include "B200Solver.hpp"

namespace Bonmin {

// cuts of one getOuterApproximation call over B points: CSR rows, the source row of each (m = the objective cut)
struct CutRows { std::vector<int> row_start{0}, col, src_row; std::vector<double> val, lb, ub;
                 int size() const { return (int)lb.size(); } };

// B x getOuterApproximation(cs, x, getObj = 1) with the side of ranged rows picked by `duals` = the multipliers of the
// last NLP solve (getRowPrice, BonOsiTMINLPInterface.cpp:1952,1968-1974); rows that are not generated are left out
using OaCutsFn = std::function<int(int B, const double *X, const double *duals, CutRows &out)>;
// extractLinearRelaxation + addObjectiveFunction at (x, duals)
using OaExtractFn = std::function<int(const double *x, const double *duals, LpRelaxation &out)>;
// the SubMipSolver: min obj.x + obj_offset over the rows, integer columns integral, objective below `cutoff`
// (OsiDualObjectiveLimit, BonOACutGenerator2.cpp:150).  Returns 0 solved (x: n_cols, bound), 1 infeasible / cut off,
// < 0 error.
using MasterMipFn = std::function<int(int n_cols, int n_rows, const int *row_start, const int *col, const double *val,
                                      const double *row_lb, const double *row_ub, const double *col_lb, const double *col_ub,
                                      const double *obj, double obj_offset, const int *is_integer, double cutoff,
                                      double max_seconds, double *x, double *bound)>;

struct OaOptions {
  double cutoff_increment = 1e-5;    // relative, as the Python driver (oa.py) and B-BB's cutoff_decr
  double gap_tol = 1e-9;
  double time_limit = 1e10;
  int max_passes = 200;
  int add_master_point = 1;          // also linearise at the master point (ECP style), one engine call for both points
};

struct OaResult {
  int status = 2;                    // 0 optimal, 1 infeasible, 2 limit reached, < 0 provider error
  double best_obj = 1e100, lower_bound = -1e200;
  int passes = 0, nlp_solves = 0, cuts = 0, oa_points = 0;
  double seconds = 0, master_seconds = 0;
  std::vector<double> best_x;
};

OaResult oa_run(const TMINLP2TNLP &root, const BatchSolveFn &solve, const OaCutsFn &oa, const OaExtractFn &extract,
                const MasterMipFn &master, const OaOptions &opt);

// OaFeasibilityChecker::performOa (src/Algorithms/OaGenerators/BonOaFeasChecker.cpp:38-135), the generator B-QG / B-Hyb call
// when a node LP has an integer solution: fix the integers at the LP values, solve NLP(y), update incumbent and cutoff,
// linearise at the NLP point, install the cuts in the node LP and re-solve it (MasterMipFn with no integer column); go on
// while the LP stays feasible below the cutoff, integer, and different on the integers from the NLP point.
// (add_only_violated_oa, default "no" in the reference, is not offered here: OaCutsFn carries no second point.)
struct OaLp {                       // the node LP: CSR rows, columns, objective
  int n_cols = 0; double obj_offset = 0;
  std::vector<int> row_start{0}, col; std::vector<double> val, row_lb, row_ub, col_lb, col_ub, obj;
};
struct OaFeasOptions { double integer_tolerance = 1e-6, cutoff_increment = 1e-5; int max_passes = 50; };
struct OaFeasResult {
  int status = 0;                   // < 0: provider error
  double cutoff = 1e100, incumbent_obj = 1e100;
  bool has_incumbent = false, fathomed = false, has_x_lp = true;
  int cuts = 0, nlp_solves = 0;
  std::vector<double> incumbent_x, x_lp;
  CutRows new_rows;                 // the cuts installed (they are global: the caller keeps them)
};
OaFeasResult oa_feas_check(const TMINLP2TNLP &root, OaLp &lp /* rows are appended */, const double *x_lp, double cutoff,
                           const BatchSolveFn &solve, const OaCutsFn &oa, const MasterMipFn &lp_solve, const OaFeasOptions &opt);

// OaNlpOptim::generateCuts (src/Algorithms/OaGenerators/BonOaNlpOptim.cpp:63-173): the NLP relaxation of the NODE (integer
// columns take the node LP's bounds), OA cuts at its solution (getObj = 1), and the solution itself when it happens to be
// integer (1e-7).  status -3 / -2 of the NLP (abandoned, iteration limit) is the reference's "Unsolved NLP ... exit": rc -3.
struct OaNlpOptimResult { int rc = 0, nlp_status = 0; double obj = 0; bool has_solution = false;
                          std::vector<double> x, duals; CutRows cuts; };
OaNlpOptimResult oa_nlp_optim(const TMINLP2TNLP &root, const double *node_lb, const double *node_ub, const double *x_lp,
                              const BatchSolveFn &solve, const OaCutsFn &oa);

// EcpCuts::generateCuts (src/Algorithms/OaGenerators/BonEcpCuts.cpp:46-143) on an LP: while the LP optimum violates the
// nonlinear rows (getNonLinearitiesViolation), linearise them AT THE LP POINT, install the cuts, re-solve.  `cs` collects the
// cuts of EVERY round — the last round's too, which are generated but neither installed nor re-solved (:86).
using OaViolationFn = std::function<int(const double *x, double obj, double *viol)>;
struct OaEcpResult { int rc = 0, rounds = 0, installed = 0; bool lp_infeasible = false; double violation = 0, obj = 0;
                     CutRows cs; std::vector<double> x; };
OaEcpResult oa_ecp_rounds(int n, int m, OaLp &lp /* installed cuts are appended */, const int *is_integer, const double *duals,
                          int rounds, double abs_violation_tol, double rel_violation_tol, const OaCutsFn &oa,
                          const OaViolationFn &violation, const MasterMipFn &lp_solve);

}  // namespace Bonmin
