// Stand-ins for the COIN-OR types the adapters are written against — used ONLY when the real headers are not on the include
// path.  With Ipopt / Bonmin installed (`IpTNLP.hpp` and `BonTNLPSolver.hpp` found by __has_include) this header includes
// the real ones instead and defines BNLP_REAL_COIN 1: B200Solver's TNLPSolver part (B200Solver.hpp: OptimizeTNLP /
// ReOptimizeTNLP / warm starts / clone / newUnsolvedError) is written against the signatures BOTH provide, so the same
// source builds as a Bonmin plugin.  The shim reproduces names, argument lists and status values of
//   Ipopt::TNLP             Ipopt IpTNLP.hpp (stable/3.13): get_nlp_info / get_bounds_info / get_starting_point /
//                           eval_* / finalize_solution, Ipopt::SolverReturn (IpAlgTypes.hpp), Ipopt::SmartPtr / ReferencedObject
//   TNLPSolver              src/Interfaces/BonTNLPSolver.hpp:26-237
//   TMINLP2TNLP             src/Interfaces/BonTMINLP2TNLP.hpp:32-505 (state :402-427; handshake BonTMINLP2TNLP.cpp:335-509)
//   IpoptWarmStart          src/Interfaces/Ipopt/BonIpoptWarmStart.hpp:28-100 (.cpp:20-75)
//   StrongBranchingSolver   src/Interfaces/BonStrongBranchingSolver.hpp:17-66
//   OsiTMINLPInterface      src/Interfaces/BonOsiTMINLPInterface.hpp:48-1346 (the slice the branching code touches)
// The host B-BB / dive / retry drivers (scaffolding for the missing Cbc) use the shim's public data members and are
// compiled only with the shim (BNLP_REAL_COIN 0).
#pragma once
#if defined(__has_include)
#if __has_include("IpTNLP.hpp") && __has_include("BonTNLPSolver.hpp") && !defined(BNLP_FORCE_SHIM)
#define BNLP_REAL_COIN 1
#endif
#endif
#ifndef BNLP_REAL_COIN
#define BNLP_REAL_COIN 0
#endif

#if BNLP_REAL_COIN
#include "IpTNLP.hpp"
#include "BonTNLPSolver.hpp"
#include "BonTMINLP2TNLP.hpp"
#include "BonIpoptWarmStart.hpp"
#include "BonStrongBranchingSolver.hpp"
#include "BonOsiTMINLPInterface.hpp"
#else
#include <cstddef>
#include <istream>
#include <ostream>
#include <string>
#include <vector>

namespace Ipopt {
typedef double Number;
typedef int Index;

// IpReferenced.hpp / IpSmartPtr.hpp: intrusive reference counting (objects handed to a SmartPtr live on the heap)
class ReferencedObject {
 public:
  ReferencedObject() : refs_(0) {}
  virtual ~ReferencedObject() {}
  void AddRef() const { ++refs_; }
  bool ReleaseRef() const { return --refs_ == 0; }
  int ReferenceCount() const { return refs_; }

 private:
  mutable int refs_;
};
template <class T>
class SmartPtr {
 public:
  SmartPtr() : p_(nullptr) {}
  SmartPtr(T *p) : p_(p) { if (p_) p_->AddRef(); }
  SmartPtr(const SmartPtr<T> &o) : p_(o.p_) { if (p_) p_->AddRef(); }
  template <class U>
  SmartPtr(const SmartPtr<U> &o) : p_(o.raw()) { if (p_) p_->AddRef(); }
  ~SmartPtr() { release(); }
  SmartPtr<T> &operator=(const SmartPtr<T> &o) { SmartPtr<T> t(o); std::swap(p_, t.p_); return *this; }
  SmartPtr<T> &operator=(T *p) { SmartPtr<T> t(p); std::swap(p_, t.p_); return *this; }
  T *operator->() const { return p_; }
  T &operator*() const { return *p_; }
  T *raw() const { return p_; }

 private:
  void release() { if (p_ && p_->ReleaseRef()) delete p_; p_ = nullptr; }
  T *p_;
};
template <class T> T *GetRawPtr(const SmartPtr<T> &s) { return s.raw(); }
template <class T> bool IsValid(const SmartPtr<T> &s) { return s.raw() != nullptr; }
template <class T> bool IsNull(const SmartPtr<T> &s) { return s.raw() == nullptr; }

// IpAlgTypes.hpp
enum SolverReturn {
  SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, STOP_AT_ACCEPTABLE_POINT, LOCAL_INFEASIBILITY,
  USER_REQUESTED_STOP, FEASIBLE_POINT_FOUND, DIVERGING_ITERATES, RESTORATION_FAILURE, ERROR_IN_STEP_COMPUTATION,
  INVALID_NUMBER_DETECTED, TOO_FEW_DEGREES_OF_FREEDOM, INVALID_OPTION, OUT_OF_MEMORY, INTERNAL_ERROR, UNASSIGNED
};
class IpoptData;
class IpoptCalculatedQuantities;

// IpTNLP.hpp: the user's NLP as the solver sees it
class TNLP : public ReferencedObject {
 public:
  enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
  virtual bool get_nlp_info(Index &n, Index &m, Index &nnz_jac_g, Index &nnz_h_lag, IndexStyleEnum &index_style) = 0;
  virtual bool get_bounds_info(Index n, Number *x_l, Number *x_u, Index m, Number *g_l, Number *g_u) = 0;
  virtual bool get_starting_point(Index n, bool init_x, Number *x, bool init_z, Number *z_L, Number *z_U, Index m,
                                  bool init_lambda, Number *lambda) = 0;
  virtual bool eval_f(Index n, const Number *x, bool new_x, Number &obj_value) = 0;
  virtual bool eval_grad_f(Index n, const Number *x, bool new_x, Number *grad_f) = 0;
  virtual bool eval_g(Index n, const Number *x, bool new_x, Index m, Number *g) = 0;
  virtual bool eval_jac_g(Index n, const Number *x, bool new_x, Index m, Index nele_jac, Index *iRow, Index *jCol,
                          Number *values) = 0;
  virtual bool eval_h(Index, const Number *, bool, Number, Index, const Number *, bool, Index, Index *, Index *, Number *) {
    return false;
  }
  virtual void finalize_solution(SolverReturn status, Index n, const Number *x, const Number *z_L, const Number *z_U,
                                 Index m, const Number *g, const Number *lambda, Number obj_value,
                                 const IpoptData *ip_data, IpoptCalculatedQuantities *ip_cq) = 0;
};
}  // namespace Ipopt

// CoinWarmStart.hpp
class CoinWarmStart {
 public:
  virtual ~CoinWarmStart() {}
  virtual CoinWarmStart *clone() const = 0;
};

namespace Bonmin {

// The per-relaxation state of BonTMINLP2TNLP.hpp:402-427 behind the TNLP handshake of BonTMINLP2TNLP.cpp:335-509.  The data
// members are public for the host drivers (shim build only); B200Solver's TNLPSolver part uses the methods.
class TMINLP2TNLP : public Ipopt::TNLP {
 public:
  int n = 0, m = 0;
  std::vector<int> var_types;                 // 0 continuous, 1 binary, 2 integer
  std::vector<double> x_l_, x_u_, g_l_, g_u_;   // current bounds
  std::vector<double> x_init_;                // start point
  std::vector<double> duals_init_;            // [z_L | z_U | lambda] or empty (setDualsInit, .cpp:294-303)
  std::vector<double> x_sol_, g_sol_, duals_sol_;  // duals_sol_ = [z_L | z_U | lambda] (.cpp:491-498)
  double obj_value_ = 1e100;
  Ipopt::SolverReturn return_status_ = Ipopt::UNASSIGNED;
  int return_status = -7;                     // TNLPSolver::ReturnStatus of the last solve (exception = -7)
  // --- the reference's accessors, same names (BonTMINLP2TNLP.hpp:70-190)
  Ipopt::Index num_variables() const { return n; }
  Ipopt::Index num_constraints() const { return m; }
  const Ipopt::Number *x_l() const { return x_l_.data(); }
  const Ipopt::Number *x_u() const { return x_u_.data(); }
  const Ipopt::Number *g_l() const { return g_l_.data(); }
  const Ipopt::Number *g_u() const { return g_u_.data(); }
  const Ipopt::Number *x_init() const { return x_init_.empty() ? nullptr : x_init_.data(); }
  const Ipopt::Number *duals_init() const { return duals_init_.empty() ? nullptr : duals_init_.data(); }
  const Ipopt::Number *x_sol() const { return x_sol_.data(); }
  const Ipopt::Number *g_sol() const { return g_sol_.data(); }
  const Ipopt::Number *duals_sol() const { return duals_sol_.data(); }
  Ipopt::Number obj_value() const { return obj_value_; }
  Ipopt::SolverReturn optimization_status() const { return return_status_; }
  void SetVariableLowerBound(int i, double v) { x_l_[i] = v; }
  void SetVariableUpperBound(int i, double v) { x_u_[i] = v; }
  void setxInit(int nn, const double *x) { x_init_.assign(x, x + nn); }
  void setDualsInit(int nd, const double *d) { duals_init_.assign(d, d + nd); }
  void resetStartingPoint() { duals_init_.clear(); }
  void set_obj_value(double v) { obj_value_ = v; }
  void Set_x_sol(int nn, const double *x) { x_sol_.assign(x, x + nn); }
  void Set_dual_sol(int nn, const double *d) { duals_sol_.assign(d, d + nn); }
  // --- Ipopt::TNLP (BonTMINLP2TNLP.cpp:335-509)
  bool get_nlp_info(Ipopt::Index &nn, Ipopt::Index &mm, Ipopt::Index &nnz_jac_g, Ipopt::Index &nnz_h_lag,
                    IndexStyleEnum &index_style) override {
    nn = n; mm = m; nnz_jac_g = 0; nnz_h_lag = 0; index_style = C_STYLE;   // the structure lives in the engine's tape
    return true;
  }
  bool get_bounds_info(Ipopt::Index nn, double *xl, double *xu, Ipopt::Index mm, double *gl, double *gu) override {
    if (nn != n || mm != m) return false;
    for (int i = 0; i < n; ++i) { xl[i] = x_l_[i]; xu[i] = x_u_[i]; }
    for (int i = 0; i < m; ++i) { gl[i] = g_l_[i]; gu[i] = g_u_[i]; }
    return true;
  }
  bool get_starting_point(Ipopt::Index nn, bool init_x, double *x, bool init_z, double *z_L, double *z_U, Ipopt::Index mm,
                          bool init_lambda, double *lambda) override {
    if (nn != n || mm != m) return false;
    if (init_x) for (int i = 0; i < n; ++i) x[i] = x_init_[i];
    if (init_z) {
      if (duals_init_.empty()) return false;                                 // .cpp:377-378
      for (int i = 0; i < n; ++i) { z_L[i] = duals_init_[i]; z_U[i] = duals_init_[n + i]; }
    }
    if (init_lambda) {
      if (duals_init_.empty()) return false;                                 // .cpp:385-386
      for (int i = 0; i < m; ++i) lambda[i] = duals_init_[2 * n + i];
    }
    return true;
  }
  // the functions are evaluated on the device from the tape the engine holds; the host-side TNLP has nothing to evaluate
  bool eval_f(Ipopt::Index, const double *, bool, double &) override { return false; }
  bool eval_grad_f(Ipopt::Index, const double *, bool, double *) override { return false; }
  bool eval_g(Ipopt::Index, const double *, bool, Ipopt::Index, double *) override { return false; }
  bool eval_jac_g(Ipopt::Index, const double *, bool, Ipopt::Index, Ipopt::Index, Ipopt::Index *, Ipopt::Index *, double *) override {
    return false;
  }
  void finalize_solution(Ipopt::SolverReturn status, Ipopt::Index nn, const double *x, const double *z_L, const double *z_U,
                         Ipopt::Index mm, const double *g, const double *lambda, double obj, const Ipopt::IpoptData *,
                         Ipopt::IpoptCalculatedQuantities *) override {           // .cpp:474-509
    x_sol_.assign(x, x + nn);
    if (mm > 0) g_sol_.assign(g, g + mm); else g_sol_.clear();
    duals_sol_.assign((size_t)mm + 2 * nn, 0.0);
    if (lambda) {
      for (int i = 0; i < mm; ++i) duals_sol_[2 * nn + i] = lambda[i];
      for (int i = 0; i < nn; ++i) { duals_sol_[i] = z_L[i]; duals_sol_[nn + i] = z_U[i]; }
    }
    return_status_ = status;
    obj_value_ = obj;
    // (shim only, for the host drivers) the status as IpoptSolver::solverReturnStatus maps it (BonIpoptSolver.cpp:304-343)
    return_status = status == Ipopt::SUCCESS ? 1 : status == Ipopt::STOP_AT_ACCEPTABLE_POINT ? 2
                  : status == Ipopt::LOCAL_INFEASIBILITY ? 3 : status == Ipopt::MAXITER_EXCEEDED ? -3
                  : status == Ipopt::ERROR_IN_STEP_COMPUTATION ? -2 : status == Ipopt::TOO_FEW_DEGREES_OF_FREEDOM ? -1
                  : status == Ipopt::INVALID_OPTION ? -5 : -7;
  }
};

// IpoptWarmStart as IpoptSolver fills and reads it (BonIpoptWarmStart.hpp:28-100): the primal point and the
// [z_L | z_U | lambda] multipliers; "empty" = reset to the model's own starting point.
class IpoptWarmStart : public CoinWarmStart {
 public:
  explicit IpoptWarmStart(bool empty = true) : empty_(empty) {}
  IpoptWarmStart(int primal_size, int dual_size, const double *primal, const double *dual)
      : primal_(primal, primal + primal_size), dual_(dual ? dual : primal, dual ? dual + dual_size : primal), empty_(false) {}
  IpoptWarmStart(const Ipopt::SmartPtr<TMINLP2TNLP> tnlp, void * /* warm_starter */)      // from the last solution (.cpp:31-44)
      : primal_(tnlp->x_sol(), tnlp->x_sol() + tnlp->num_variables()),
        dual_(tnlp->duals_sol(), tnlp->duals_sol() + 2 * tnlp->num_variables() + tnlp->num_constraints()), empty_(false) {}
  CoinWarmStart *clone() const override { return new IpoptWarmStart(*this); }
  bool empty() const { return empty_; }
  int primalSize() const { return (int)primal_.size(); }
  int dualSize() const { return (int)dual_.size(); }
  const double *primal() const { return primal_.data(); }
  const double *dual() const { return dual_.data(); }

 private:
  std::vector<double> primal_, dual_;
  bool empty_;
};

class TNLPSolver : public Ipopt::ReferencedObject {
 public:
  enum ReturnStatus {  // BonTNLPSolver.hpp:29-44
    iterationLimit = -3, timeLimit = 5, doesNotConverge = -8, computationError = -2, notEnoughFreedom = -1,
    illDefinedProblem = -4, illegalOption = -5, externalException = -6, exception = -7, solvedOptimal = 1,
    solvedOptimalTol = 2, provenInfeasible = 3, unbounded = 4, numReturnCodes
  };
  // thrown BY POINTER by the interface when a solve ends in an unrecoverable status (:50-97, BonOsiTMINLPInterface.cpp:2639-2644)
  class UnsolvedError {
   public:
    UnsolvedError(int errorNum = -10000, Ipopt::SmartPtr<TMINLP2TNLP> model = nullptr, std::string name = "")
        : errorNum_(errorNum), model_(model), name_(std::move(name)) {}
    virtual ~UnsolvedError() {}
    void printError(std::ostream &os) {      // BonTNLPSolver.cpp:165-173
      os << solverName() << " exited with error code " << errorNum_ << " " << errorName() << std::endl;
    }
    virtual const std::string &errorName() const = 0;
    virtual const std::string &solverName() const = 0;
    int errorNum() const { return errorNum_; }
    Ipopt::SmartPtr<TMINLP2TNLP> model() const { return model_; }
    const std::string &name() const { return name_; }

   private:
    int errorNum_;
    Ipopt::SmartPtr<TMINLP2TNLP> model_;
    std::string name_;
  };
  virtual UnsolvedError *newUnsolvedError(int num, Ipopt::SmartPtr<TMINLP2TNLP> problem, std::string name) = 0;   // :95-97
  virtual ~TNLPSolver() {}
  virtual Ipopt::SmartPtr<TNLPSolver> clone() = 0;                                       // :110
  virtual bool Initialize(std::string params_file) = 0;                                  // :118
  virtual bool Initialize(std::istream &is) = 0;                                         // :122
  virtual ReturnStatus OptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) = 0;        // :127
  virtual ReturnStatus ReOptimizeTNLP(const Ipopt::SmartPtr<Ipopt::TNLP> &tnlp) = 0;      // :130
  virtual bool setWarmStart(const CoinWarmStart *warm, Ipopt::SmartPtr<TMINLP2TNLP> tnlp) = 0;      // :133-134
  virtual CoinWarmStart *getUsedWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const = 0;             // :137
  virtual CoinWarmStart *getWarmStart(Ipopt::SmartPtr<TMINLP2TNLP> tnlp) const = 0;                 // :140
  virtual CoinWarmStart *getEmptyWarmStart() const = 0;                                  // :142
  virtual bool warmStartIsValid(const CoinWarmStart *ws) const = 0;                      // :145
  virtual void enableWarmStart() = 0;                                                    // :148
  virtual void disableWarmStart() = 0;                                                   // :151
  virtual double CPUTime() = 0;                                                          // :181
  virtual int IterationCount() = 0;                                                      // :184
  virtual void setOutputToDefault() = 0;                                                 // :188
  virtual void forceSolverOutput(int log_level) = 0;                                     // :190
  virtual std::string &solverName() = 0;                                                 // :192
  virtual int errorCode() const = 0;                                                     // :210
  // BonTNLPSolver.cpp:253-255
  bool isRecoverable(ReturnStatus &r) {
    return !(r == illDefinedProblem || r == notEnoughFreedom || r == illegalOption || r == computationError || r == timeLimit);
  }
  bool isError(ReturnStatus &r) { return r < 0; }
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
  OsiTMINLPInterface(Ipopt::SmartPtr<TMINLP2TNLP> problem, Ipopt::SmartPtr<TNLPSolver> app) : problem_(problem), app_(app) {}
  TMINLP2TNLP *problem() { return GetRawPtr(problem_); }
  TNLPSolver *solver() { return GetRawPtr(app_); }
  void SetStrongBrachingSolver(StrongBranchingSolver *s) { strong_branching_solver_ = s; }   // sic, :3028-3032
  void setColLower(int i, double v) { problem_->SetVariableLowerBound(i, v); }
  void setColUpper(int i, double v) { problem_->SetVariableUpperBound(i, v); }
  const double *getColLower() const { return problem_->x_l(); }
  const double *getColUpper() const { return problem_->x_u(); }
  const double *getColSolution() const { return problem_->x_sol(); }
  const double *getRowPrice() const { return problem_->duals_sol(); }               // [z_L | z_U | lambda] (:1304-1308)
  double getObjValue() const { return problem_->obj_value(); }
  int getNumCols() const { return problem_->n; }
  void initialSolve() {   // :2779-2836 without the retry ladder (num_retry_unsolved_random_point = 0)
    app_->disableWarmStart();
    optimizationStatus_ = app_->OptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(problem_)));
    nCallOptimizeTNLP_++;
    totalIterations_ += app_->IterationCount();
  }
  // resolve (:2839-2890): with a warm start that carries multipliers the solver re-optimises from it (setWarmStart,
  // enableWarmStart, ReOptimizeTNLP); without one (`warm_start none`, the default) it is initialSolve
  void resolve() {
    if (!warmstart_ || !app_->warmStartIsValid(warmstart_)) { initialSolve(); return; }
    app_->setWarmStart(warmstart_, problem_);
    delete warmstart_;
    warmstart_ = nullptr;
    if (problem_->duals_init() != nullptr) app_->enableWarmStart(); else app_->disableWarmStart();
    optimizationStatus_ = app_->ReOptimizeTNLP(Ipopt::SmartPtr<Ipopt::TNLP>(GetRawPtr(problem_)));
    nCallOptimizeTNLP_++;
    totalIterations_ += app_->IterationCount();
  }
  CoinWarmStart *getWarmStart() const { return app_->getWarmStart(problem_); }           // :1381-1389
  bool setWarmStart(const CoinWarmStart *ws) {                                           // :1392-1402
    delete warmstart_;
    warmstart_ = ws ? ws->clone() : nullptr;
    return true;
  }
  ~OsiTMINLPInterface() { delete warmstart_; }
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
  Ipopt::SmartPtr<TMINLP2TNLP> problem_;
  Ipopt::SmartPtr<TNLPSolver> app_;
  CoinWarmStart *warmstart_ = nullptr;
  StrongBranchingSolver *strong_branching_solver_ = nullptr;
  TNLPSolver::ReturnStatus optimizationStatusBeforeHotStart_ = TNLPSolver::exception;
  int nCallOptimizeTNLP_ = 0;
  long totalIterations_ = 0;
};

}  // namespace Bonmin
#endif  // BNLP_REAL_COIN
