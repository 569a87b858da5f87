// Internal declarations shared by the host model compiler, the CUDA kernels and the C ABI.
// Product code: nothing here may include or link anything under oracle/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace bnlp {

constexpr double kInf = 1e19;  // Bonmin's infinity (BonOsiTMINLPInterface.cpp:3002-3005)

// One general nonlinear element term  phi(x[vars])  added to function `fn` (0 = objective).
// prog is a postfix program over local variable slots, see include/bnlp.h (BNLP_OP_*).
struct Term {
  int fn;
  std::vector<int> vars;
  std::vector<int> op;       // opcode per instruction
  std::vector<int> iarg;     // local variable slot for BNLP_OP_VAR
  std::vector<double> darg;  // constant for BNLP_OP_CONST / BNLP_OP_POWC
};

// Host-side "flat tape": every function k (0 = objective, 1..m = rows) is
//   F_k(x) = c_k + a_k.x + 1/2 x'Q_k x + sum_t phi_t(x[S_t])
struct HostModel {
  int n = 0, m = 0;
  std::vector<double> xl, xu, gl, gu, x0, c;
  std::vector<int> vtype;
  // linear part, CSR over functions 0..m
  std::vector<int> lin_start, lin_col;
  std::vector<double> lin_val;
  // dense quadratic blocks (packed lower, row-major) and their function index
  std::vector<int> qd_fn;
  std::vector<double> qd;  // qd_fn.size() * n(n+1)/2
  // sparse quadratic triplets (i >= j, Q_ij = Q_ji = v, convention 1/2 x'Qx)
  std::vector<int> qs_fn, qs_i, qs_j;
  std::vector<double> qs_v;
  std::vector<Term> terms;
  // branching information from the .nl variable suffixes (AmplTMINLP::read_priorities, BonAmplTMINLP.cpp:215-252);
  // empty = suffix absent
  std::vector<int> br_priority, br_direction;
  std::vector<double> br_up_ps, br_down_ps;
  // SOS suffixes of the variables (AmplTMINLP::read_sos, BonAmplTMINLP.cpp:254-296): .sosno (set number: > 0 type 1, < 0
  // type 2, 0 = in no set) and .ref (reference weight); empty = suffix absent
  std::vector<int> sos_no;
  std::vector<double> sos_ref;
  // derived structures (TMINLP::eval_jac_g / eval_h sparsity, BonTMINLP.hpp:286-315), 0-based
  std::vector<int> jac_row, jac_col;       // rows 0..m-1 (constraints only), row-major sorted
  std::vector<int> jac_start;              // m+1
  std::vector<int> hes_row, hes_col;       // lower triangle, row >= col, sorted
  std::vector<int> grad_f_cols;            // structural nonzeros of grad f
  std::vector<int> con_nonlinear;          // m: 1 if the row has a quadratic part or a term
  bool obj_nonlinear = false;
  void finalize();                         // builds the derived structures
};

bool parse_nl(const char *text, size_t len, HostModel &out, std::string &err);

// deepest postfix stack of a (validated) term program
int term_stack_depth(const Term &T);
// 0: every term fits the narrow term machine (BNLP_MAX_TERM_VARS / _STACK); 1: the model needs the wide build
// (BNLP_MAX_WIDE_TERM_VARS / _STACK hold for every term); -1 + err: a term fits neither
int term_width_class(const HostModel &M, std::string &err);

}  // namespace bnlp
