// Minimal stand-ins for the COIN-OR types the adapters are written against.  The real headers (IpTNLP.hpp,
// BonTNLPSolver.hpp, BonTMINLP2TNLP.hpp, BonStrongBranchingSolver.hpp, OsiSolverInterface.hpp) are absent from this
// image; these shims reproduce ONLY the names, argument meaning and status values the hot path
// uses, so that B200Solver.cpp reads like a Bonmin plugin and INTEGRATION.md's real-header variant is a mechanical edit.
// Reference signatures:  TNLPSolver            src/Interfaces/BonTNLPSolver.hpp:26-237
//                        TMINLP2TNLP           src/Interfaces/BonTMINLP2TNLP.hpp:32-505 (state: :402-427)
//                        StrongBranchingSolver src/Interfaces/BonStrongBranchingSolver.hpp:17-66
//                        OsiTMINLPInterface    src/Interfaces/BonOsiTMINLPInterface.hpp:48-1346 (subset used here)
#pragma once
#include <istream>
#include <ostream>
#include <string>
#include <vector>

namespace Bonmin {

class TMINLP2TNLP;

// CoinWarmStart as IpoptWarmStart fills it (src/Interfaces/Ipopt/BonIpoptWarmStart.hpp:28-100): the primal point and the
// [z_L | z_U | lambda] multipliers; "empty" = reset to the model's own starting point.
class CoinWarmStart {
 public:
  CoinWarmStart() {}
  CoinWarmStart(int np, int nd, const double *x, const double *duals) : primal_(x, x + np), dual_(duals ? duals : x, duals ? duals + nd : x), empty_(false) {}
  bool empty() const { return empty_; }
  int primalSize() const { return (int)primal_.size(); }
  int dualSize() const { return (int)dual_.size(); }
  const double *primal() const { return primal_.data(); }
  const double *dual() const { return dual_.data(); }

 private:
  std::vector<double> primal_, dual_;
  bool empty_ = true;
};

class TNLPSolver {
 public:
  enum ReturnStatus {  // BonTNLPSolver.hpp:29-44
    iterationLimit = -3, timeLimit = 5, doesNotConverge = -8, computationError = -2, notEnoughFreedom = -1,
    illDefinedProblem = -4, illegalOption = -5, externalException = -6, exception = -7, solvedOptimal = 1,
    solvedOptimalTol = 2, provenInfeasible = 3, unbounded = 4, numReturnCodes
  };
  // thrown BY POINTER by the interface when a solve ends in an unrecoverable status (:50-97, BonOsiTMINLPInterface.cpp:2639-2644)
  class UnsolvedError {
   public:
    UnsolvedError(int errorNum = -10000, TMINLP2TNLP *model = nullptr, std::string name = "") : errorNum_(errorNum), model_(model), name_(std::move(name)) {}
    virtual ~UnsolvedError() {}
    void printError(std::ostream &os) {      // BonTNLPSolver.cpp:165-173
      os << solverName() << " exited with error code " << errorNum_ << " " << errorName() << std::endl;
    }
    virtual const std::string &errorName() const = 0;
    virtual const std::string &solverName() const = 0;
    int errorNum() const { return errorNum_; }
    TMINLP2TNLP *model() const { return model_; }
    const std::string &name() const { return name_; }

   private:
    int errorNum_;
    TMINLP2TNLP *model_;
    std::string name_;
  };
  virtual UnsolvedError *newUnsolvedError(int num, TMINLP2TNLP *problem, std::string name) = 0;   // :95-97
  virtual ~TNLPSolver() {}
  virtual TNLPSolver *clone() = 0;
  virtual bool Initialize(std::string params_file) = 0;         // :118
  virtual bool Initialize(std::istream &is) = 0;                // :122
  virtual ReturnStatus OptimizeTNLP(TMINLP2TNLP &tnlp) = 0;     // :127
  virtual ReturnStatus ReOptimizeTNLP(TMINLP2TNLP &tnlp) = 0;   // :130
  virtual bool setWarmStart(const CoinWarmStart *warm, TMINLP2TNLP &tnlp) = 0;     // :133-134
  virtual CoinWarmStart *getUsedWarmStart(TMINLP2TNLP &tnlp) const = 0;            // :137
  virtual CoinWarmStart *getWarmStart(TMINLP2TNLP &tnlp) const = 0;                // :140
  virtual CoinWarmStart *getEmptyWarmStart() const = 0;                            // :142
  virtual bool warmStartIsValid(const CoinWarmStart *ws) const = 0;                // :145
  virtual void enableWarmStart() = 0;                                              // :148
  virtual void disableWarmStart() = 0;                                             // :151
  virtual double CPUTime() = 0;
  virtual int IterationCount() = 0;
  virtual void setOutputToDefault() = 0;                        // :174
  virtual void forceSolverOutput(int log_level) = 0;            // :176
  virtual const std::string &solverName() const = 0;
  virtual int errorCode() const = 0;                            // :205
  // BonTNLPSolver.cpp:253-255
  static bool isRecoverable(ReturnStatus r) {
    return !(r == illDefinedProblem || r == notEnoughFreedom || r == illegalOption || r == computationError || r == timeLimit);
  }
  static bool isError(ReturnStatus r) { return r < 0; }
};

// The per-relaxation mutable state of BonTMINLP2TNLP.hpp:402-427 (same member meaning).
class TMINLP2TNLP {
 public:
  int n = 0, m = 0;
  std::vector<int> var_types;                 // 0 continuous, 1 binary, 2 integer
  std::vector<double> x_l, x_u, g_l, g_u;     // current bounds
  std::vector<double> x_init;                 // [x | z_L | z_U | lambda] when duals are set (:294-303)
  std::vector<double> duals_init;             // [z_L | z_U | lambda] or empty (setDualsInit, .cpp:286-303)
  void setDualsInit(int nd, const double *d) { duals_init.assign(d, d + nd); }
  std::vector<double> x_sol, g_sol, duals_sol;  // duals_sol = [z_L | z_U | lambda] (.cpp:491-498)
  double obj_value = 1e100;
  TNLPSolver::ReturnStatus return_status = TNLPSolver::exception;
  void SetVariableLowerBound(int i, double v) { x_l[i] = v; }
  void SetVariableUpperBound(int i, double v) { x_u[i] = v; }
  void setxInit(int nn, const double *x) { x_init.assign(x, x + nn); }
  void set_obj_value(double v) { obj_value = v; }
  void Set_x_sol(int nn, const double *x) { x_sol.assign(x, x + nn); }
};

class OsiTMINLPInterface;

class StrongBranchingSolver {  // BonStrongBranchingSolver.hpp:34-40
 public:
  virtual ~StrongBranchingSolver() {}
  virtual void markHotStart(OsiTMINLPInterface *tminlp_interface) = 0;
  virtual TNLPSolver::ReturnStatus solveFromHotStart(OsiTMINLPInterface *tminlp_interface) = 0;
  virtual void unmarkHotStart(OsiTMINLPInterface *tminlp_interface) = 0;
};

// The slice of OsiTMINLPInterface the branching code touches (BonOsiTMINLPInterface.cpp:2779-2955, 3028-3095).
class OsiTMINLPInterface {
 public:
  OsiTMINLPInterface(TMINLP2TNLP *problem, TNLPSolver *app) : problem_(problem), app_(app) {}
  TMINLP2TNLP *problem() { return problem_; }
  TNLPSolver *solver() { return app_; }
  void SetStrongBrachingSolver(StrongBranchingSolver *s) { strong_branching_solver_ = s; }   // sic, :3028-3032
  void setColLower(int i, double v) { problem_->SetVariableLowerBound(i, v); }
  void setColUpper(int i, double v) { problem_->SetVariableUpperBound(i, v); }
  const double *getColLower() const { return problem_->x_l.data(); }
  const double *getColUpper() const { return problem_->x_u.data(); }
  const double *getColSolution() const { return problem_->x_sol.data(); }
  double getObjValue() const { return problem_->obj_value; }
  int getNumCols() const { return problem_->n; }
  void initialSolve() {   // :2779-2836 without the retry ladder (num_retry_unsolved_random_point = 0)
    optimizationStatus_ = app_->OptimizeTNLP(*problem_);
    nCallOptimizeTNLP_++;
    totalIterations_ += app_->IterationCount();
  }
  void resolve() { initialSolve(); }   // warm_start=none: resolve == initialSolve (:2851-2859)
  bool isProvenOptimal() const {        // :2923-2928
    return optimizationStatus_ == TNLPSolver::solvedOptimal || optimizationStatus_ == TNLPSolver::solvedOptimalTol;
  }
  bool isProvenPrimalInfeasible() const { return optimizationStatus_ == TNLPSolver::provenInfeasible; }  // :2930-2935
  bool isAbandoned() const {            // :2907-2920
    return optimizationStatus_ == TNLPSolver::iterationLimit || optimizationStatus_ == TNLPSolver::computationError ||
           optimizationStatus_ == TNLPSolver::illDefinedProblem || optimizationStatus_ == TNLPSolver::illegalOption ||
           optimizationStatus_ == TNLPSolver::externalException || optimizationStatus_ == TNLPSolver::exception;
  }
  void markHotStart() {                 // :3040-3055
    if (strong_branching_solver_) { optimizationStatusBeforeHotStart_ = optimizationStatus_; strong_branching_solver_->markHotStart(this); }
  }
  void solveFromHotStart() {            // :3058-3078
    if (strong_branching_solver_) optimizationStatus_ = strong_branching_solver_->solveFromHotStart(this);
    else resolve();
  }
  void unmarkHotStart() {               // :3081-3095
    if (strong_branching_solver_) { strong_branching_solver_->unmarkHotStart(this); optimizationStatus_ = optimizationStatusBeforeHotStart_; }
  }
  int nCallOptimizeTNLP() const { return nCallOptimizeTNLP_; }
  long totalIterations() const { return totalIterations_; }
  TNLPSolver::ReturnStatus optimizationStatus_ = TNLPSolver::exception;

 private:
  TMINLP2TNLP *problem_;
  TNLPSolver *app_;
  StrongBranchingSolver *strong_branching_solver_ = nullptr;
  TNLPSolver::ReturnStatus optimizationStatusBeforeHotStart_ = TNLPSolver::exception;
  int nCallOptimizeTNLP_ = 0;
  long totalIterations_ = 0;
};

}  // namespace Bonmin
