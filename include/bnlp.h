/* bnlp.h — C ABI of the B200 batched NLP-relaxation engine (libbnlp.so).
 *
 * This is the drop-in boundary for Bonmin's data-parallel hot path.  Every entry point states the
 * reference interface it replaces (paths relative to the Bonmin source tree).  Plain pointers and
 * sizes only; no exceptions cross the ABI; naming/ownership mirror the reference's own C API
 * (src/Interfaces/BonStdCInterface.h:133-204: opaque handles, Create/Free pairs, caller-owned arrays).
 *
 * Conventions (identical to the reference):
 *   - duals layout [z_L (n) | z_U (n) | lambda (m)]         src/Interfaces/BonTMINLP2TNLP.cpp:491-498
 *   - lambda sign: L = f + lambda'g (Ipopt)                  src/Interfaces/BonOsiTMINLPInterface.cpp:1968-1974
 *   - STATUS values are Bonmin::TNLPSolver::ReturnStatus     src/Interfaces/BonTNLPSolver.hpp:29-44
 *   - an infeasible child reports obj = max-norm violation   src/Interfaces/BonTMINLP2TNLP.cpp:503-505
 *   - |bound| >= 1e19 means infinite                         src/Interfaces/BonOsiTMINLPInterface.cpp:3002-3005
 *   - all structure indices are 0-based (C_STYLE)            src/Interfaces/BonTMINLP.hpp:286-315
 * Return value of every int function: 0 = ok, <0 = engine error (message via bnlp_last_error()); the
 * host adapters map a negative return to TNLPSolver::exception.  One host thread drives one engine.
 * There is NO CPU fallback: without a CUDA device bnlp_engine_create fails.
 */
#ifndef BNLP_H
#define BNLP_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bnlp_problem bnlp_problem;
typedef struct bnlp_engine bnlp_engine;

/* TNLPSolver::ReturnStatus (src/Interfaces/BonTNLPSolver.hpp:29-44) */
enum {
  BNLP_iterationLimit = -3, BNLP_timeLimit = 5, BNLP_doesNotConverge = -8, BNLP_computationError = -2,
  BNLP_notEnoughFreedom = -1, BNLP_illDefinedProblem = -4, BNLP_illegalOption = -5,
  BNLP_externalException = -6, BNLP_exception = -7, BNLP_solvedOptimal = 1, BNLP_solvedOptimalTol = 2,
  BNLP_provenInfeasible = 3, BNLP_unbounded = 4
};

/* opcodes of a nonlinear element term (postfix program over the term's local variable slots) */
enum {
  BNLP_OP_CONST = 0, BNLP_OP_VAR = 1, BNLP_OP_ADD = 2, BNLP_OP_SUB = 3, BNLP_OP_MUL = 4, BNLP_OP_DIV = 5,
  BNLP_OP_NEG = 6, BNLP_OP_POWC = 7 /* x^darg */, BNLP_OP_POW = 8, BNLP_OP_EXP = 9, BNLP_OP_LOG = 10,
  BNLP_OP_SQRT = 11, BNLP_OP_SIN = 12, BNLP_OP_COS = 13, BNLP_OP_SQR = 14, BNLP_OP_ABS = 15,
  BNLP_OP_ATAN = 16, BNLP_OP_ASIN = 17, BNLP_OP_ACOS = 18
};
#define BNLP_MAX_TERM_VARS 6
#define BNLP_MAX_TERM_STACK 12
/* WIDE element terms: a non-quadratic expression that couples more than BNLP_MAX_TERM_VARS variables (a log-sum-exp over 20
 * variables, sqrt(sum x^2), ...; the reference forwards any ASL expression DAG, src/Interfaces/Ampl/BonAmplTMINLP.cpp:629-710).
 * A model that holds at least one such term is evaluated by a second build of the term machine (same second-order forward
 * mode, dense local Hessian of up to 24 x 24): every term of that model may then couple up to BNLP_MAX_WIDE_TERM_VARS
 * variables with a postfix stack of at most BNLP_MAX_WIDE_TERM_STACK entries. */
#define BNLP_MAX_WIDE_TERM_VARS 24
#define BNLP_MAX_WIDE_TERM_STACK 8

/* Structured model ("flat tape" input): every function k (k = 0 objective, k = 1..m rows) is
 *     F_k(x) = c[k] + a_k.x + 1/2 x'Q_k x + sum_t phi_t(x[term vars])
 * This is the closed-form statement of what a Bonmin::TMINLP exposes through callbacks
 * (src/Interfaces/BonTMINLP.hpp:226-333: get_nlp_info, get_bounds_info, get_variables_types,
 *  get_starting_point, eval_f/grad_f/g/jac_g/h). */
typedef struct bnlp_tape_desc {
  int n, m;
  const double *xl, *xu;       /* n */
  const double *gl, *gu;       /* m */
  const int *var_type;         /* n : 0 continuous, 1 binary, 2 integer (TMINLP::VariableType) */
  const double *x0;            /* n : starting point (may be NULL = 0) */
  const double *c;             /* m+1 */
  const int *lin_start;        /* m+2 : CSR over functions 0..m */
  const int *lin_col;
  const double *lin_val;
  int n_qdense;                /* dense quadratic blocks */
  const int *qd_fn;            /* n_qdense : function index of each block */
  const double *qd;            /* n_qdense x n(n+1)/2 : packed lower triangle, row-major (r >= c) */
  int n_qsparse;               /* sparse quadratic triplets, i >= j, Q_ij = Q_ji = v */
  const int *qs_fn, *qs_i, *qs_j;
  const double *qs_v;
  int n_terms;                 /* nonlinear element terms */
  const int *term_fn;          /* n_terms */
  const int *term_var_start;   /* n_terms+1 -> term_var */
  const int *term_var;
  const int *term_op_start;    /* n_terms+1 -> term_op / term_iarg / term_darg */
  const int *term_op;
  const int *term_iarg;
  const double *term_darg;
} bnlp_tape_desc;

typedef struct bnlp_options {
  double tol;                  /* Ipopt tol, default 1e-8 (doc/options_list_ipopt_content.tex L2689) */
  int max_iter;                /* default 3000 (L2671) */
  int ctas_per_device;         /* 0 = one persistent CTA per SM */
} bnlp_options;

const char *bnlp_last_error(void);

/* --- model -------------------------------------------------------------------------------------
 * bnlp_problem_from_nl   replaces  AmplTMINLP::Initialize -> new Ipopt::AmplTNLP -> ASL reader
 *                        (src/Interfaces/Ampl/BonAmplTMINLP.cpp:83-204,473-564)
 * bnlp_problem_from_tape replaces a user TMINLP subclass (examples/CppExample/MyTMINLP.cpp:12-217) */
bnlp_problem *bnlp_problem_from_nl(const char *nl_text, size_t len, char *err, size_t errlen);
bnlp_problem *bnlp_problem_from_tape(const bnlp_tape_desc *desc, char *err, size_t errlen);
void bnlp_problem_free(bnlp_problem *);

/* Feasibility-pump / local-branching variant of a problem = Bonmin::TNLP2FPNLP (src/Interfaces/BonTNLP2FPNLP.cpp:75-330):
 *   objective  (1 - lambda) sigma f(x) + objective_scaling_factor lambda dist(x, target)   when use_fp_objective
 *              (norm 2: sum (x_i - t_i)^2; norm 1, binaries: sum [t_i <= 0.1 ? x_i : 1 - x_i]),
 *   a cutoff row  (that objective) <= cutoff            when use_cutoff_constraint (pass the value set_cutoff computes, :52-63),
 *   a local-branching row  sum [t_i <= 0.1 ? x_i : 1 - x_i] <= rhs   when use_local_branching_constraint.
 * The target enters as n_inds EXTRA variables appended after the n original ones (variable n + i is the target of
 * variable inds[i]); a child fixes them through its bounds (XL = XU = t_i), so a whole batch of rounding targets is one
 * bnlp_solve_batch call on the returned problem.  Returns NULL on error. */
typedef struct bnlp_fp_desc {
  int n_inds;
  const int *inds;
  int norm;                                  /* 1 or 2 (BonTNLP2FPNLP.hpp: norm_) */
  double lambda, sigma, objective_scaling_factor;
  int use_fp_objective, use_cutoff_constraint, use_local_branching_constraint;
  double cutoff, rhs_local_branching;
} bnlp_fp_desc;
bnlp_problem *bnlp_problem_fp(const bnlp_problem *base, const bnlp_fp_desc *desc, char *err, size_t errlen);

/* Cut-strengthening variant of a problem = CutStrengthener::StrengtheningTNLP
 * (src/Interfaces/BonCutStrengthener.cpp:219-560):  min obj_grad.x  s.t.  g_l <= g_i(x) <= g_u  over the box — its optimal
 * value is the tightest valid bound of the cut  obj_grad.x >= bound  (obj_grad = +/- the cut row, :349-354).
 * constr_index = -1: the objective cut,  min obj_grad.(x, z)  s.t.  f(x) - z <= 0  (z = variable n, :438-441, :531-533).
 * The cut coefficients are nv EXTRA variables (nv = n, or n + 1 with z; variable nv + j is the coefficient of variable
 * j) that a child fixes through its bounds (XL = XU = obj_grad_j), so all cuts of one row are one bnlp_solve_batch
 * call; variables that do not appear in row i are fixed at the cut point the same way (the reference leaves them out
 * of the NLP, :362-376).  Returns NULL on error. */
bnlp_problem *bnlp_problem_cut_strengthening(const bnlp_problem *base, int constr_index, char *err, size_t errlen);

/* get_nlp_info (BonTMINLP.hpp:226-229) */
int bnlp_problem_sizes(const bnlp_problem *, int *n, int *m, int *nnz_jac, int *nnz_hess);
/* structure + bounds + types: get_bounds_info :238, get_variables_types :208, get_constraints_linearity :220,
 * eval_jac_g/eval_h structure calls :286-315, get_starting_point :251.  Any pointer may be NULL. */
int bnlp_problem_info(const bnlp_problem *, int *jac_irow, int *jac_jcol, int *hes_irow, int *hes_jcol,
                      int *var_type, int *con_nonlinear, double *xl, double *xu, double *gl, double *gu,
                      double *x0);

/* Branching information of the model = TMINLP::branchingInfo() as AmplTMINLP::read_priorities fills it from the .nl
 * variable suffixes (src/Interfaces/Ampl/BonAmplTMINLP.cpp:215-252): priorities (9999 - suffix value: the LOWEST number is
 * branched on first), branching directions, up / down pseudo-costs (one given: used for both).  Arrays of n or NULL.
 * Returns a bit mask of what the model has (1 priorities, 2 directions, 4 up, 8 down), 0 for none, -1 on error. */
/* TMINLP::sosConstraints() (src/Interfaces/BonTMINLP.hpp:114-150 SosInfo) as AmplTMINLP::read_sos fills it from the .nl
 * variable suffixes .sosno / .ref (src/Interfaces/Ampl/BonAmplTMINLP.cpp:254-296): num sets, num_nz members; per set its type
 * (1 or 2: sign of .sosno), priority (9999 - the largest .priority of its members) and starts[num + 1] into indices /
 * weights (members by increasing .ref, made strictly increasing as ampl_utils::sos_kludge does, sos_kludge.cpp:29-45).
 * Call once with null arrays for the sizes.  The grouping itself is ASL's suf_sos, which is not in the reference tree: restated
 * from its documented meaning (parity-unpinned).  The engine does not branch on the sets (SOS branching objects are Cbc's). */
int bnlp_problem_sos_info(const bnlp_problem *, int *num, int *num_nz, char *types, int *priorities, int *starts, int *indices,
                          double *weights);
int bnlp_problem_branching_info(const bnlp_problem *, int *priorities, int *directions, double *up_pseudo,
                                double *down_pseudo);

/* --- engine ------------------------------------------------------------------------------------
 * replaces the construction of an IpoptSolver (src/Interfaces/BonOsiTMINLPInterface.cpp:448-496). The tape
 * is uploaded (replicated) to every listed device. */
bnlp_engine *bnlp_engine_create(const bnlp_problem *, const int *device_ids, int ndev, const bnlp_options *);
void bnlp_engine_free(bnlp_engine *);

/* Optional: page-lock a caller-owned host buffer that is handed to the batch calls repeatedly (cudaHostRegister /
 * cudaHostUnregister): the copies inside the calls then run as DMA.  The caller unregisters before freeing the buffer. */
int bnlp_host_register(void *p, size_t bytes);
int bnlp_host_unregister(void *p);

/* B x {eval_f, eval_grad_f, eval_g, eval_jac_g values, eval_h values} (BonTMINLP.hpp:268-315).
 * X: B x n.  LAM: B x m or NULL (then 0), sigma = obj_factor.  Outputs (any may be NULL):
 * F: B, G: B x m, GRADF: B x n, JVAL: B x nnz_jac, HVAL: B x nnz_hess.  Host pointers. */
int bnlp_eval_batch(bnlp_engine *, int B, const double *X, const double *LAM, double sigma, double *F,
                    double *G, double *GRADF, double *JVAL, double *HVAL);

/* B child relaxations = B x TNLPSolver::OptimizeTNLP (src/Interfaces/BonTNLPSolver.hpp:127,
 * src/Interfaces/Ipopt/BonIpoptSolver.cpp:134-176), differing only in their bound vectors
 * (src/Algorithms/Branching/BonChooseVariable.cpp:807-905).  XL, XU: B x n.  X0: n (shared, x0_per_child = 0)
 * or B x n.  Outputs: OBJ B, STATUS B, ITERS B, X B x n, G B x m, DUALS B x (2n+m); X/G/DUALS may be NULL.
 * Host pointers; copies are inside the call. */
int bnlp_solve_batch(bnlp_engine *, int B, const double *XL, const double *XU, const double *X0,
                     int x0_per_child, double *OBJ, int *STATUS, int *ITERS, double *X, double *G,
                     double *DUALS);
/* The strong-branching form of the same call: child b is parent child_parent[b] (one of P open nodes, bounds
 * parent_XL/parent_XU: P x n) with ONE bound of column child_col[b] replaced by child_bound[b] — its upper bound when
 * child_way[b] == 0 (the down branch, setColUpper(floor(x_j))), its lower bound when 1 (up branch, setColLower(ceil(x_j)))
 * — exactly what BonChooseVariable::doStrongBranching does between markHotStart and unmarkHotStart
 * (src/Algorithms/Branching/BonChooseVariable.cpp:807-905).  The bound vectors of the B children are expanded on the
 * device, so a round of 128 nodes x 64 children moves 0.3 MB to the GPU instead of 13 MB.  X0: n (shared) or, when
 * x0_per_parent, P x n (each child starts from its parent's point).  Outputs as bnlp_solve_batch. */
int bnlp_solve_children(bnlp_engine *, int P, const double *parent_XL, const double *parent_XU, int B,
                        const int *child_parent, const int *child_col, const int *child_way, const double *child_bound,
                        const double *X0, int x0_per_parent, double *OBJ, int *STATUS, int *ITERS, double *X, double *G,
                        double *DUALS);
/* Warm-started re-solve = IpoptSolver::ReOptimizeTNLP after setWarmStart (src/Interfaces/Ipopt/BonIpoptSolver.cpp:180-220,
 * 369-419; OsiTMINLPInterface::resolve, src/Interfaces/BonOsiTMINLPInterface.cpp:2851-2874; exercised by
 * test/InterfaceTest.cpp:165-189).  DUAL0: the multipliers [z_L | z_U | lambda] the solve starts from — one vector of
 * 2n+m for every child (dual0_per_child = 0) or B rows — with Ipopt's warm_start_init_point = yes treatment: the start point
 * and the slacks are pushed inside by 1e-3 (warm_start_bound_push / _bound_frac / _slack_bound_push / _slack_bound_frac)
 * instead of 1e-2, every multiplier is taken from DUAL0 and pushed to at least 1e-3 (warm_start_mult_bound_push).
 * DUAL0 == NULL is bnlp_solve_batch. */
int bnlp_solve_batch_warm(bnlp_engine *, int B, const double *XL, const double *XU, const double *X0, int x0_per_child,
                          const double *DUAL0, int dual0_per_child, double *OBJ, int *STATUS, int *ITERS, double *X,
                          double *G, double *DUALS);
/* Same with DEVICE pointers on device `dev_index` of the engine (inputs already resident in HBM; runs on `stream`,
 * a cudaStream_t passed as void*, no synchronisation). */
int bnlp_solve_batch_device(bnlp_engine *, int dev_index, int B, const double *dXL, const double *dXU,
                            const double *dX0, int x0_per_child, double *dOBJ, int *dSTATUS, int *dITERS,
                            double *dX, double *dG, double *dDUALS, void *stream);

/* Outer-approximation rows at B points = B x OsiTMINLPInterface::getOuterApproximation
 * (src/Interfaces/BonOsiTMINLPInterface.cpp:1924-2107 incl. cleanNnz :1881-1919).
 * X: B x n; DUALS: B x (2n+m) or NULL (multipliers pick the side of ranged rows); XL/XU: n column bounds used by
 * cleanNnz; X2 + theta: violation filter, X2 = the LP points the cuts are tested against, B x bnlp_oa_x2_width() or
 * NULL: the width is n + 1 whenever the objective cut is generated (getObj == 2, or getObj == 1 on a model with a
 * nonlinear objective) because that cut is tested against the LP's alpha column x2[n] (:2077-2090), else n;
 * getObj as in the reference (0/1/2).
 * Output is CSR with a fixed capacity of (m_nl + 1) rows per point:
 *   row_start: B*(m_nl+1)+1, col_idx/val: capacity from bnlp_oa_capacity(), LB/UB: B*(m_nl+1),
 *   src_row: B*(m_nl+1) (constraint index, m for the objective cut, -1 = row not generated). */
typedef struct bnlp_oa_tols {
  double tiny;        /* tiny_element 1e-8       (BonOsiTMINLPInterface.cpp:236-241) */
  double very_tiny;   /* very_tiny_element 1e-17 (:243-247) */
  double rhs_relax;   /* oa_rhs_relax 1e-8       (:222-228) */
  double infty;       /* 1e19 */
  int skip_non_violated; /* 0 = reference behaviour at oa log level 0: the x2/theta filter never drops a
                            constraint cut (BonOsiTMINLPInterface.cpp:2018-2020); 1 = drop cuts violated by < theta */
} bnlp_oa_tols;
int bnlp_oa_capacity(const bnlp_engine *, int B, int *rows_per_point, size_t *nnz_capacity);
int bnlp_oa_x2_width(const bnlp_engine *, int getObj);   /* doubles per row of X2 for this getObj (n or n + 1) */
int bnlp_oa_batch(bnlp_engine *, int B, const double *X, const double *DUALS, const double *XL, const double *XU,
                  const double *X2, double theta, int getObj, const bnlp_oa_tols *tols, int *row_start,
                  int *col_idx, double *val, double *LB, double *UB, int *src_row);

/* B x OsiTMINLPInterface::getConstraintsViolation (OBJ == NULL) or getNonLinearitiesViolation (OBJ: B objective values)
 * (src/Interfaces/BonOsiTMINLPInterface.cpp:1839-1876): max bound violation over the NONLINEAR rows (bounds beyond
 * +-1e10 ignored), and max with f(x) - obj.  VIOL: B.  F (optional): f(x) per point. */
int bnlp_violation_batch(bnlp_engine *, int B, const double *X, const double *OBJ, double *VIOL, double *F);
/* B x OsiTMINLPInterface::getBendersCut (src/Interfaces/BonOsiTMINLPInterface.cpp:2111-2202): one aggregated cut per point,
 * sum over nonlinear rows with |lambda| > 1e-6 of lambda * (linearisation), plus the objective cut when the objective is
 * nonlinear and with_objective != 0 (the reference's isProvenOptimal()).  Output: fixed capacity n+1 per point
 * (col_idx/val: B x (n+1), row_len: B), UB: B; the cut's lower bound is -COIN_DBL_MAX. */
int bnlp_benders_batch(bnlp_engine *, int B, const double *X, const double *DUALS, const double *XL, const double *XU,
                       int with_objective, const bnlp_oa_tols *tols, int *row_len, int *col_idx, double *val, double *UB);

/* QP strong branching.
 * bnlp_branching_tqp   = Bonmin::BranchingTQP (src/Interfaces/BonBranchingTQP.cpp:22-108 construction, :136-153 bounds,
 *   :192-309 callbacks): the node NLP frozen at its optimum (x_sol: n, duals_sol: 2n+m = [z_L|z_U|lambda]) as a QP in the
 *   displacement d = x - x_sol:  min grad f(x_sol).d + 1/2 d'W d  (W = Hessian of the Lagrangian, obj_factor 1),
 *   g_L - g(x_sol) <= J(x_sol) d <= g_U - g(x_sol),  x_L - x_sol <= d <= x_U - x_sol.  Returned as an ordinary problem
 *   (caller frees); *obj_val = f(x_sol).
 * bnlp_tqp_solve_batch = B x QpBranchingSolver::solveFromHotStart (src/Algorithms/Branching/BonQpBranchingSolver.cpp:50-116):
 *   children given by their bounds in the ORIGINAL variables (XL, XU: B x n), solved on that QP from d = 0; results as
 *   BranchingTQP::finalize_solution reports them (BonBranchingTQP.cpp:311-333): X = x_sol + d, OBJ = QP objective +
 *   f(x_sol), an infeasible child keeps its violation.  X may be NULL. */
bnlp_problem *bnlp_branching_tqp(bnlp_engine *, const double *x_sol, const double *duals_sol, double *obj_val, char *err,
                                 size_t errlen);
int bnlp_tqp_solve_batch(bnlp_engine *, const double *x_sol, const double *duals_sol, int B, const double *XL,
                         const double *XU, double *OBJ, int *STATUS, int *ITERS, double *X);

/* The whole initial LP relaxation at x = OsiTMINLPInterface::extractLinearRelaxation + addObjectiveFunction
 * (src/Interfaces/BonOsiTMINLPInterface.cpp:2378-2590): linear rows copied, nonlinear rows linearised at x with the
 * +-1e-7 slack (:2430-2440), ranged nonlinear rows with a zero multiplier removed (:2421-2425) and one-sided by the sign
 * of the multiplier otherwise, coefficients cleaned by cleanNnz, upper sides relaxed by oa_rhs_relax; getObj != 0: a linear
 * objective is copied (offset -f(0)), a nonlinear one adds the column alpha (cost 1, free) and the objective cut.
 * duals: 2n+m (getRowPrice layout) or NULL (= all zero); xl/xu: current column bounds or NULL (the model's).
 * Output: row-major CSR; capacities row_start m+2, col_idx/val nnz_jac+n+1, row_lb/row_ub/src_row m+1 (src_row = model
 * row, m for the objective cut), col_lb/col_ub/obj/is_integer n+1.  LP infinity = DBL_MAX (OsiClp's getInfinity()). */
typedef struct bnlp_lp_dims { int n_cols, n_rows, nnz, has_alpha; double obj_offset; } bnlp_lp_dims;
int bnlp_extract_linear_relaxation(bnlp_engine *, const double *x, const double *duals, const double *xl, const double *xu,
                                   int getObj, const bnlp_oa_tols *tols, bnlp_lp_dims *dims, int *row_start, int *col_idx,
                                   double *val, double *row_lb, double *row_ub, int *src_row, double *col_lb,
                                   double *col_ub, double *obj, int *is_integer);

/* B x OsiTMINLPInterface::getConstraintOuterApproximation (src/Interfaces/BonOsiTMINLPInterface.cpp:2207-2275): the
 * linearisation of ONE row rowIdx (linear or nonlinear) at B points, side chosen by the multiplier, coefficients cleaned
 * by cleanNnz, no rhs relaxation.  Output rows have fixed capacity n: col_idx/val B x n, row_len/LB/UB B. */
int bnlp_constraint_oa_batch(bnlp_engine *, int B, int rowIdx, const double *X, const double *DUALS, const double *XL,
                             const double *XU, const bnlp_oa_tols *tols, int *row_len, int *col_idx, double *val,
                             double *LB, double *UB);

/* statistics of the last solve call: total IPM iterations, factorisations, device milliseconds
 * (counterparts of totalIterations_/totalNlpSolveTime_, src/Interfaces/BonOsiTMINLPInterface.hpp:482-495) */
int bnlp_engine_stats(const bnlp_engine *, long long *iterations, long long *factorisations, double *device_ms);
/* sizes of the engine's problem (get_nlp_info, BonTMINLP.hpp:226-229) */
int bnlp_engine_sizes(const bnlp_engine *, int *n, int *m);
/* number of kernel launches issued by the engine since creation (bench.py's gpu_launches) */
long long bnlp_engine_launches(const bnlp_engine *);
/* measured FP64 FMA peak of device dev_index in TFLOP/s (micro-kernel; roofline denominator) */
double bnlp_measure_fp64_peak(bnlp_engine *, int dev_index);

/* --- host B-BB driver (host control flow, NOT accelerated; Cbc is not part of this engine) --------------------
 * Restates the node loop that Bab::branchAndBound + BonChooseVariable run for algorithm B-BB
 * (src/CbcBonmin/BonCbc.cpp:104-690, src/Algorithms/Branching/BonChooseVariable.cpp:599-957): NLP relaxation per node,
 * strong branching with all 2K children of a node solved as ONE batch and the reference's sequential loop replayed
 * against the cached results, max-min scoring, pseudo-costs, best-bound + plunge.  Option names/defaults:
 * src/Algorithms/BonBabSetupBase.cpp:478-559.  0 in an option field means "reference default". */
typedef int (*bnlp_batch_solve_cb)(void *ctx, int B, const double *XL, const double *XU, const double *X0, double *OBJ,
                                   int *STATUS, int *ITERS, double *X, double *G, double *DUALS);
typedef struct bnlp_bb_options {
  int number_strong_branch, number_before_trust, node_limit, log_level;
  double integer_tolerance, cutoff_decr, time_limit;
  int parallel_nodes;          /* > 1: expand that many best-bound nodes per round with one batched launch (not in the
                                  reference: its tree search is strictly one node at a time) */
  int variable_selection;      /* 0 = NLP strong branching (`variable_selection strong-branching`), 1 = QP strong
                                  branching (`qp-strong-branching`, src/Algorithms/BonBonminSetup.cpp:447-455): the 2K
                                  children of a node are solved on the BranchingTQP surrogate; their results rank the
                                  candidates only (trustStrongForSolution = false, never reused as node relaxations) */
  const int *priorities;       /* n branching priorities or NULL: among the fractional integers only those with the LOWEST
                                  value are candidates (OsiChooseVariable::setupList as kept by
                                  BonChooseVariable::setupList, Branching/BonChooseVariable.cpp:307-584).
                                  bnlp_bb_run_engine uses the problem's own (.nl suffix) priorities when this is NULL */
  int nlp_failure_behavior;    /* a node NLP that ends in an abandoned status (iteration limit, computation error:
                                  OsiTMINLPInterface::isAbandoned, BonOsiTMINLPInterface.cpp:2907-2920) is never read as
                                  infeasible: 0 = stop, status 4 (the reference's `nlp_failure_behavior stop`: newUnsolvedError,
                                  :2639-2644), 1 = fathom that node and go on (`fathom`, pretendFailIsInfeasible_) */
  int number_look_ahead;       /* `number_look_ahead` (BonChooseVariable.cpp:152-158, 930-948): away from the root, stop the
                                  strong-branching loop after that many candidates in a row that did not improve the best
                                  score; 0 = off (the default) */
  const double *incumbent_x;   /* n values or NULL: a known integer-feasible point with objective incumbent_obj (what a primal
                                  heuristic found before the tree search starts: B-Hyb's HeuristicDive*,
                                  src/CbcBonmin/Heuristics/BonHeuristicDive.cpp:67-302, registered in
                                  BonBonminSetup.cpp:300-330).  The search starts with cutoff = incumbent_obj - cutoff_decr,
                                  as CbcModel::setBestSolution does */
  double incumbent_obj;
} bnlp_bb_options;
typedef struct bnlp_bb_result {
  int status;                 /* 0 optimal, 1 infeasible, 2 limit reached, 3 the batch-solve provider returned an error (the
                                 call returns -1), 4 a node NLP stayed unsolved and nlp_failure_behavior = stop */
  int nodes, nlp_solves, sb_batches, sb_children, served_from_sb_cache;
  long long ipm_iterations;
  double best_obj, best_bound, seconds;
  int failed_nodes;           /* node NLPs that ended in an abandoned status */
} bnlp_bb_result;
/* arithmetic provider = any batch solver with bnlp_solve_batch's meaning (tests pass the CPU oracle here) */
int bnlp_bb_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                   const double *gu, const double *x0, bnlp_batch_solve_cb cb, void *ctx, const bnlp_bb_options *o,
                   bnlp_bb_result *res, double *best_x);
/* same, plus the QP-surrogate provider used when variable_selection = 1 (meaning of bnlp_tqp_solve_batch) */
typedef int (*bnlp_tqp_solve_cb)(void *ctx, const double *x_sol, const double *duals_sol, int B, const double *XL,
                                 const double *XU, double *OBJ, int *STATUS, int *ITERS, double *X);
int bnlp_bb_run_cb2(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                    const double *gu, const double *x0, bnlp_batch_solve_cb cb, bnlp_tqp_solve_cb qp_cb, void *ctx,
                    const bnlp_bb_options *o, bnlp_bb_result *res, double *best_x);
/* same, plus a provider for the compact strong-branching form (meaning of bnlp_solve_children, shared start point); NULL =
 * the driver expands the children's bound vectors itself and uses `cb` */
typedef int (*bnlp_children_solve_cb)(void *ctx, int P, const double *parent_XL, const double *parent_XU, int B,
                                      const int *child_parent, const int *child_col, const int *child_way,
                                      const double *child_bound, const double *X0, double *OBJ, int *STATUS, int *ITERS,
                                      double *X);
int bnlp_bb_run_cb3(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                    const double *gu, const double *x0, bnlp_batch_solve_cb cb, bnlp_tqp_solve_cb qp_cb,
                    bnlp_children_solve_cb children_cb, void *ctx, const bnlp_bb_options *o, bnlp_bb_result *res,
                    double *best_x);
/* arithmetic provider = the CUDA engine */
int bnlp_bb_run_engine(bnlp_engine *, const bnlp_problem *, const bnlp_bb_options *o, bnlp_bb_result *res,
                       double *best_x);

/* --- Multi-rank exchange inside the engine (one process per GPU) --------------------------------------
 * The children of a batch are independent: ranks solve contiguous blocks of the child index space and the only
 * communication is an ncclAllGather of per-child {objective, status, iterations} (+ primal point) and an ncclAllReduce(min)
 * of the incumbent bound (what BM_lp_branch.cpp:445-535,746-811 does by message passing in the reference's experimental
 * BCP code).  NCCL is bound at run time (dlopen libnccl.so.2).
 *   bnlp_comm_unique_id        rank 0: 128 bytes to hand to every rank (the host broadcasts them, e.g. torch.distributed)
 *   bnlp_comm_create           collective over the ranks (ncclCommInitRank); the engine must own exactly one device
 *   bnlp_solve_batch_sharded   SAME input arrays on every rank; every rank receives all B results
 *   bnlp_comm_allreduce_min    incumbent bound
 *   bnlp_bb_run_engine_sharded bnlp_bb_run_engine with every batch of children sharded: all ranks run the same search on
 *                              identical results and return the same optimum (checked with the all-reduce) */
typedef struct bnlp_comm bnlp_comm;
int bnlp_comm_unique_id(unsigned char *id128);
bnlp_comm *bnlp_comm_create(bnlp_engine *, int rank, int world, const unsigned char *id128);
void bnlp_comm_free(bnlp_comm *);
int bnlp_comm_stats(const bnlp_comm *, long long *collectives, double *exchange_ms);
int bnlp_comm_allreduce_min(bnlp_comm *, double *value);
int bnlp_solve_batch_sharded(bnlp_comm *, int B, const double *XL, const double *XU, const double *X0, int x0_per_child,
                             double *OBJ, int *STATUS, int *ITERS, double *X);
int bnlp_bb_run_engine_sharded(bnlp_comm *, const bnlp_problem *, const bnlp_bb_options *o, bnlp_bb_result *res, double *best_x);

/* --- NLP diving heuristics (host control flow; the NLPs of a step are one batch) --------------------------------
 * Restates HeuristicDive::solution (src/CbcBonmin/Heuristics/BonHeuristicDive.cpp:67-302) with the selection rules of
 * HeuristicDiveFractional (BonHeuristicDiveFractional.cpp:55-108, rule 0) and HeuristicDiveVectorLength
 * (BonHeuristicDiveVectorLength.cpp:93-159, rule 1).  The dive starts from a solved relaxation (xl/xu: its bounds, x_sol,
 * g_sol, obj_sol); cutoff = incumbent value (1e100 if none).  speculate != 0: the retry ladder of one step (:171-207 — with
 * at-bound fixings, without them, other rounding direction) is solved as ONE batch and replayed; 0: one NLP at a time. */
typedef int (*bnlp_grad_f_cb)(void *ctx, const double *x, double *grad_f);
typedef struct bnlp_dive_options {
  int rule, speculate;
  double integer_tolerance, percentage_to_fix;   /* 0 = reference defaults 1e-6, 0.2 */
  unsigned seed;
} bnlp_dive_options;
typedef struct bnlp_dive_result {
  int found, steps, nlp_solves, batches;
  long long ipm_iterations;
  double obj;
} bnlp_dive_result;
int bnlp_dive_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *gl,
                     const double *gu, const double *x_sol, const double *g_sol, double obj_sol, const int *column_length,
                     double cutoff, bnlp_batch_solve_cb cb, bnlp_grad_f_cb grad_cb, void *ctx, const bnlp_dive_options *o,
                     bnlp_dive_result *res, double *better_x);
int bnlp_dive_run_engine(bnlp_engine *, const bnlp_problem *, const double *xl, const double *xu, const double *x_sol,
                         const double *g_sol, double obj_sol, double cutoff, const bnlp_dive_options *o,
                         bnlp_dive_result *res, double *better_x);

/* --- retry policies around one NLP (host control flow; the restarts are one batch with per-child start points) ------
 * mode 0 = OsiTMINLPInterface::resolveForRobustness(numsolve) (src/Interfaces/BonOsiTMINLPInterface.cpp:989-1093): the
 *   user's start point, then numsolve uniform random points (randomStartingPoint, :1749-1795); the first result that is not
 *   abandoned (isAbandoned, :2907-2920) is kept; *used = its index (0 = user point), -1 if every attempt was abandoned.
 * mode 1 = resolveForCost(numsolve) (:836-986): numsolve random starts; the best proven-optimal objective replaces the
 *   current solution (cur_obj / cur_status; pass cur_status = 0 if there is none) when it improves on it; *used = 1 + index
 *   of the start that won, -1 if the current solution stays.
 * The start callback has bnlp_batch_solve_cb's signature with X0 holding B x n start points. */
typedef struct bnlp_retry_result { int used, num_failed, num_infeasible, nlp_solves; double mean, std_dev; } bnlp_retry_result;
int bnlp_resolve_cb(int mode, int n, int m, const double *xl, const double *xu, const double *x_init, double cur_obj,
                    int cur_status, int numsolve, double max_random_point_radius, unsigned seed,
                    bnlp_batch_solve_cb cb_starts, void *ctx, bnlp_retry_result *res, double *obj, int *status, double *x,
                    double *g, double *duals);
int bnlp_resolve_engine(bnlp_engine *, int mode, const double *xl, const double *xu, const double *x_init, double cur_obj,
                        int cur_status, int numsolve, double max_random_point_radius, unsigned seed, bnlp_retry_result *res,
                        double *obj, int *status, double *x, double *g, double *duals);

/* The host lineariser (B200Linearizer : TMINLP2OsiLP, src/Interfaces/BonTMINLP2OsiLP.hpp:73-83) with the engine behind it:
 * get_oas at B points (cuts flattened to CSR: row_start n_cuts + 1, col/val, lb, ub, is_global) and extract at the
 * first point (dimensions of the loaded LP).  -2 = a capacity is too small. */
int bnlp_linearizer_engine(bnlp_engine *, const bnlp_problem *, int B, const double *X, int getObj, int global, int *n_cuts,
                           int *row_start, int *col, double *val, double *lb, double *ub, int *is_global, int cap_cuts,
                           int cap_nnz, int *lp_cols, int *lp_rows, int *lp_nnz, int *lp_integers);
/* --- B-OA in the host layer (bonmin_b200/host/B200Oa.cpp): the decomposition loop of OACutGenerator2::performOa
 * (src/Algorithms/OaGenerators/BonOACutGenerator2.cpp:121-268; fixIntegers / installCuts, src/Algorithms/BonSolverHelp.cpp:49-174;
 * post_nlp_solve, src/Algorithms/OaGenerators/BonOaDecBase.cpp:237-400) in C++ around the batched engine.  The master MILP
 * is the reference's SubMipSolver (src/Algorithms/OaGenerators/BonSubMipSolver.hpp: optimize(cutoff, loglevel, maxTime),
 * getLastSolution, lowBound) — Cbc there, a callback here: it receives the whole master (initial relaxation + every cut
 * installed so far, CSR by rows, +-DBL_MAX = free side) and the cutoff on the objective, returns 0 with a point (n_cols)
 * and its bound, 1 when the master is infeasible / cut off, < 0 on error.
 * bnlp_oa_run_engine: relaxation, every NLP(y), the initial LP and the linearisations (at the NLP(y) point and, when
 * add_master_point, at the master point: one bnlp_oa_batch call per pass) run on the CUDA engine.
 * bnlp_oa_run_cb: the same loop over caller-supplied arithmetic (tests: the CPU oracle).  cuts callback = B x
 * getOuterApproximation(getObj = 1) with `duals` = the multipliers of the last NLP solve, generated rows only, CSR;
 * extract callback = extractLinearRelaxation + addObjectiveFunction (obj: n + 1 entries, the last one alpha's);
 * both return -2 when a capacity is too small.  0 in an option field means its default (1e-5, 1e-9, 1e10, 200). */
typedef int (*bnlp_master_mip_cb)(void *ctx, int n_cols, int n_rows, const int *row_start, const int *col, const double *val,
                                  const double *row_lb, const double *row_ub, const double *col_lb, const double *col_ub,
                                  const double *obj, double obj_offset, const int *is_integer, double cutoff,
                                  double max_seconds, double *x, double *bound);
typedef int (*bnlp_oa_cuts_cb)(void *ctx, int B, const double *X, const double *duals, int *n_cuts, int *row_start, int *col,
                               double *val, double *lb, double *ub, int *src_row, int cap_cuts, int cap_nnz);
typedef int (*bnlp_oa_extract_cb)(void *ctx, const double *x, const double *duals, int *n_rows, int *has_alpha,
                                  double *obj_offset, int *row_start, int *col, double *val, double *row_lb, double *row_ub,
                                  double *obj, int cap_rows, int cap_nnz);
typedef struct bnlp_oa_run_options { double cutoff_increment, gap_tol, time_limit; int max_passes, no_master_point; } bnlp_oa_run_options;
typedef struct bnlp_oa_run_result {
  int status;                 /* 0 optimal, 1 infeasible, 2 limit reached */
  int passes, nlp_solves, cuts, oa_points;
  double best_obj, lower_bound, seconds, master_seconds;
} bnlp_oa_run_result;
int bnlp_oa_run_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *x0,
                   bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts, bnlp_oa_extract_cb extract, bnlp_master_mip_cb master,
                   void *ctx, const bnlp_oa_run_options *o, bnlp_oa_run_result *res, double *best_x);
int bnlp_oa_run_engine(bnlp_engine *, const bnlp_problem *, bnlp_master_mip_cb master, void *ctx, const bnlp_oa_run_options *o,
                       bnlp_oa_run_result *res, double *best_x);
/* OaFeasibilityChecker::performOa (src/Algorithms/OaGenerators/BonOaFeasChecker.cpp:38-135), the generator B-QG / B-Hyb call
 * when a node LP has an integer solution x_lp (n_cols): fix the integers, solve NLP(y), update incumbent and cutoff, linearise
 * at the NLP point, install the cuts in the node LP (given as CSR rows + columns + objective) and re-solve it through
 * lp_solve (a bnlp_master_mip_cb called with no integer column and no cutoff); repeated while the LP stays feasible below
 * the cutoff, integer, and different on the integers from the NLP point.  Outputs: the installed cuts (CSR, global), the new
 * cutoff, the incumbent if one was found, the last LP point, fathomed = the node can be cut off.  -2: a capacity is too
 * small (res->cuts / res->new_nnz say what is needed).  add_only_violated_oa (reference default "no") is not offered. */
typedef struct bnlp_oa_feas_options { double integer_tolerance, cutoff_increment; int max_passes; } bnlp_oa_feas_options;
typedef struct bnlp_oa_feas_result {
  int status, cuts, nlp_solves, fathomed, has_incumbent, has_x_lp, new_nnz;
  double cutoff, incumbent_obj;
} bnlp_oa_feas_result;
int bnlp_oa_feas_check_cb(int n, int m, const int *var_type, const double *xl, const double *xu, int n_cols, int n_rows,
                          const int *row_start, const int *col, const double *val, const double *row_lb, const double *row_ub,
                          const double *col_lb, const double *col_ub, const double *obj, double obj_offset, const double *x_lp,
                          double cutoff, const bnlp_oa_feas_options *o, bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts,
                          bnlp_master_mip_cb lp_solve, void *ctx, bnlp_oa_feas_result *res, double *incumbent_x, double *x_lp_out,
                          int *new_row_start, int *new_col, double *new_val, double *new_lb, double *new_ub, int *new_src,
                          int cap_cuts, int cap_nnz);
int bnlp_oa_feas_check_engine(bnlp_engine *, const bnlp_problem *, int n_cols, int n_rows, const int *row_start, const int *col,
                              const double *val, const double *row_lb, const double *row_ub, const double *col_lb,
                              const double *col_ub, const double *obj, double obj_offset, const double *x_lp, double cutoff,
                              const bnlp_oa_feas_options *o, bnlp_master_mip_cb lp_solve, void *ctx, bnlp_oa_feas_result *res,
                              double *incumbent_x, double *x_lp_out, int *new_row_start, int *new_col, double *new_val,
                              double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz);
/* OaNlpOptim::generateCuts (src/Algorithms/OaGenerators/BonOaNlpOptim.cpp:63-173): the NLP relaxation of a NODE (integer
 * columns take node_lb / node_ub, n values each), started from x_lp; OA cuts at its solution (CSR out, getObj = 1); x_nlp (n)
 * and duals_out (2n+m) of the solve; has_solution = the NLP point is integer (1e-7).  Returns -3 for the reference's "Unsolved
 * NLP ... exit" (abandoned / iteration limit), -2 when a capacity is too small. */
typedef struct bnlp_nlp_optim_result { int nlp_status, cuts, new_nnz, has_solution; double obj; } bnlp_nlp_optim_result;
int bnlp_oa_nlp_optim_cb(int n, int m, const int *var_type, const double *xl, const double *xu, const double *node_lb,
                         const double *node_ub, const double *x_lp, bnlp_batch_solve_cb solve, bnlp_oa_cuts_cb cuts, void *ctx,
                         bnlp_nlp_optim_result *res, double *x_nlp, double *duals_out, int *new_row_start, int *new_col,
                         double *new_val, double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz);
int bnlp_oa_nlp_optim_engine(bnlp_engine *, const bnlp_problem *, const double *node_lb, const double *node_ub, const double *x_lp,
                             bnlp_nlp_optim_result *res, double *x_nlp, double *duals_out, int *new_row_start, int *new_col,
                             double *new_val, double *new_lb, double *new_ub, int *new_src, int cap_cuts, int cap_nnz);
/* EcpCuts::generateCuts (src/Algorithms/OaGenerators/BonEcpCuts.cpp:46-143) on an LP (CSR rows + columns + objective;
 * is_integer NULL = pure LP): while the LP optimum violates the nonlinear rows (getNonLinearitiesViolation) by more than
 * abs_tol and rel_tol x the first violation, linearise them AT THE LP POINT (duals: the multipliers that pick the side of
 * ranged rows, NULL = zeros), install the cuts, re-solve; at most `rounds` rounds.  Out: every cut generated (CSR; the first
 * res->installed of them went into the LP — the last round's are generated but not installed, :86), the last LP point x_out
 * (n_cols), res->obj = the LP value after the rounds (DBL_MAX and lp_infeasible when the LP became infeasible). */
typedef int (*bnlp_violation_cb)(void *ctx, const double *x, double obj, double *viol);
typedef struct bnlp_ecp_result { int rounds, cuts, installed, new_nnz, lp_infeasible; double violation, obj; } bnlp_ecp_result;
int bnlp_ecp_rounds_cb(int n, int m, int n_cols, int n_rows, const int *row_start, const int *col, const double *val,
                       const double *row_lb, const double *row_ub, const double *col_lb, const double *col_ub, const double *obj,
                       double obj_offset, const int *is_integer, const double *duals, int rounds, double abs_tol, double rel_tol,
                       bnlp_oa_cuts_cb cuts, bnlp_violation_cb violation, bnlp_master_mip_cb lp_solve, void *ctx,
                       bnlp_ecp_result *res, double *x_out, int *new_row_start, int *new_col, double *new_val, double *new_lb,
                       double *new_ub, int *new_src, int cap_cuts, int cap_nnz);
int bnlp_ecp_rounds_engine(bnlp_engine *, const bnlp_problem *, int n_cols, int n_rows, const int *row_start, const int *col,
                           const double *val, const double *row_lb, const double *row_ub, const double *col_lb,
                           const double *col_ub, const double *obj, double obj_offset, const int *is_integer, const double *duals,
                           int rounds, double abs_tol, double rel_tol, bnlp_master_mip_cb lp_solve, void *ctx,
                           bnlp_ecp_result *res, double *x_out, int *new_row_start, int *new_col, double *new_val, double *new_lb,
                           double *new_ub, int *new_src, int cap_cuts, int cap_nnz);
/* Self-test of the host adapter's TNLPSolver contract (options, warm starts, error objects, output levels) on a stub
 * provider; 0 = all expectations hold, otherwise the source line of the first failure (bonmin_b200/host/host_cabi.cpp). */
int bnlp_host_selftest(void);

/* --- on-disk formats either side of the path (host code; bonmin_b200/host/node_files.cpp) -------------------------
 * Failed-node dump = TMINLP2TNLP::outputDiffs (src/Interfaces/BonTMINLP2TNLP.cpp:551-615): writes <prob_name>.bounds
 * ("LO|UP\tindex\tvalue" per bound that differs from the original model, CoinRelFltEq 1e-10) and <prob_name>.start
 * ("n\t2n+m", x_init with 17 digits, then the 2n+m duals when duals_init != NULL). */
int bnlp_write_node_files(const char *prob_name, int n, int m, const double *xl, const double *xu, const double *orig_xl,
                          const double *orig_xu, const double *x_init, const double *duals_init);
/* BoundsReader::readAndApply (src/Interfaces/BonBoundsReader.cpp:52-119): applies the file's LO/UP lines to xl/xu (the
 * caller pre-fills them with the model's bounds).  Returns the number of lines applied, -1 on a malformed file. */
int bnlp_read_bounds_file(const char *file, int n, double *xl, double *xu);
/* StartPointReader::readFile (src/Interfaces/BonStartPointReader.cpp:15-37).  n_duals returns how many dual values the
 * file really holds (outputDiffs omits them when the node had none). */
int bnlp_read_start_file(const char *file, int *n_primals, int *n_duals, double *primals, int cap_primals, double *duals,
                         int cap_duals);
/* Batch replay file: all B children of one bnlp_solve_batch call in the same line format — <prob_name>.children
 * ("BATCH\tn\tm\tB\tx0_per_child", then per child "CHILD\tk" and its LO/UP lines, 17 digits) and <prob_name>.start
 * (the shared or per-child start points).  A GPU/CPU disagreement becomes a two-file repro. */
int bnlp_write_batch_replay(const char *prob_name, int n, int m, int B, const double *XL, const double *XU,
                            const double *orig_xl, const double *orig_xu, const double *X0, int x0_per_child);
int bnlp_read_batch_replay_header(const char *file, int *n, int *m, int *B, int *x0_per_child);
int bnlp_read_batch_replay(const char *file, int n, int B, const double *orig_xl, const double *orig_xu, double *XL,
                           double *XU);
/* AmplTMINLP::finalize_solution (src/Interfaces/Ampl/BonAmplTMINLP.cpp:712-753): message + AMPL solve_result_num for a
 * TMINLP::SolverReturn (0 SUCCESS 3, 1 INFEASIBLE 220, 2 CONTINUOUS_UNBOUNDED 300, 3 LIMIT_EXCEEDED 421/410,
 * 4 USER_INTERRUPT 422/411, 5 MINLP_ERROR 500); finished: 1 "Finished", 0 "Not finished", -1 "Aborted". */
int bnlp_sol_status(int solver_return, int has_x, const char *app_name, char *message, size_t msglen, int *solve_result_num,
                    int *finished);
/* option integers of the first line of a text .nl file (opts[0] = count, then the values) — what the .sol file echoes */
int bnlp_nl_header_options(const char *nl_text, size_t len, int *opts, int cap, double *vbtol);
/* AmplTMINLP::write_solution (:769-794) -> ASL write_sol, text format (ASL is not in the reference tree: published format
 * restated, parity unpinned): message, blank line, "Options" block, n_con, #duals, n_var, #primals, values (shortest
 * round-trip decimals), "objno 0 <solve_result_num>".  x or y may be NULL (count 0 is written). */
int bnlp_write_sol(const char *file, const char *message, const int *ampl_options, double ampl_vbtol, int n_con, int n_var,
                   const double *x, const double *y, int solve_result_num);
/* the non-AMPL fallback of finalize_solution (:758-766): "i\tx_i" lines and a closing -1 */
int bnlp_write_plain_sol(const char *file, int n, const double *x);

#ifdef __cplusplus
}
#endif
#endif
