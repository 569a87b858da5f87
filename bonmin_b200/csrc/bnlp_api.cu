// C ABI implementation (include/bnlp.h): model upload, batched solve / eval / OA launches.
// Product code — must not reference anything under oracle/.  No CPU fallback: every compute entry point
// needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/bnlp.h"
#include "bnlp_internal.h"

namespace bnlp {
constexpr double kInfD = 1e19;
}
#define BNLP_EVAL_KERNELS
#include "bnlp_eval.cuh"
#include "bnlp_ipm.cuh"
#include "bnlp_ipm_coop.cuh"

namespace bnlp {
bool model_from_desc(const bnlp_tape_desc *d, HostModel &M, std::string &err);
bool model_fp_variant(const HostModel &B, const bnlp_fp_desc &d, HostModel &M, std::string &err);
bool model_cs_variant(const HostModel &B, int constr_index, HostModel &M, std::string &err);
// kernel launchers live in their own translation units (bnlp_launch_ipm.cu, bnlp_launch_coop.cu) so the three big
// kernels compile in parallel
cudaError_t ipm_general_set_smem(size_t smem);
cudaError_t ipm_general_launch(bool wide, int ctas, size_t smem, cudaStream_t stream, const DevModel &M, const IpmArgs &A);
cudaError_t ipm_coop_set_smem(size_t smem);
cudaError_t ipm_coop_launch(bool timers, int ctas, size_t smem, cudaStream_t stream, DevModel &M, IpmArgs &A, CoopBufs &CB);
}

using namespace bnlp;

// wide build of the term machine (bnlp_launch_wide.cu: namespace bnlpw, BNLP_TV = BNLP_MAX_WIDE_TERM_VARS)
cudaError_t bnlpw_ipm_set_smem(size_t smem);
cudaError_t bnlpw_ipm_launch(bool big, int ctas, size_t smem, cudaStream_t stream, const void *M, const void *A);
cudaError_t bnlpw_eval_launch(int nc, size_t smem, cudaStream_t stream, const void *ev, const void *args);
// 256-thread build of the cooperative kernel (bnlp_launch_coop256.cu)
cudaError_t bnlp_coop256_set_smem(size_t smem);
cudaError_t bnlp_coop256_launch(bool timers, int ctas, size_t smem, cudaStream_t stream, void *M, void *A, void *CB);
int bnlp_coop256_max_ctas_per_sm(size_t smem);

static thread_local std::string g_err;
static void set_err(const std::string &s) { g_err = s; }
extern "C" const char *bnlp_last_error(void) { return g_err.c_str(); }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

struct bnlp_problem {
  HostModel M;
};

namespace {

template <class T>
struct DBuf {
  T *p = nullptr;
  size_t cap = 0;
  int reserve(size_t cnt) {
    if (cnt <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, std::max<size_t>(cnt, 1) * sizeof(T)) != cudaSuccess) return -1;
    cap = cnt;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Device {
  int id = 0;
  int sms = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // outer-approximation pipeline: copy stream, per-buffer-set events, double-buffered cut buffers
  cudaStream_t stream2 = nullptr;
  cudaEvent_t oa_comp[2] = {nullptr, nullptr}, oa_copy[2] = {nullptr, nullptr}, oa_k0[2] = {nullptr, nullptr};
  DBuf<double> oaVal[2], oaLB[2], oaUB[2];
  DBuf<int> oaCI[2], oaSR[2], oaRS[2];
  // model
  DBuf<double> c, A, Qp, Qf, row_sgn, row_bnd, gl, gu, xl, xu;
  DBuf<int> qidx, row_fn, fn_row_up, fn_row_lo, eq_fn, fn_eq, dense_rows, sp_seg_start, sp_seg_i, sp_seg_k, sp_ent_row, sp_rows, spc_row;
  DBuf<double> sp_ent_p, lrV, spc_val, sph_rec;
  DBuf<int> fn_lr_start;
  DBuf<double> eq_bnd;
  EvalModelDev ev;   // eval / OA view of the tape
  DBuf<int> e_int;   // packed integer arrays of the eval tape
  DBuf<double> e_dbl;
  // solve buffers
  DBuf<double> dXL, dXU, dX0, dOBJ, dX, dG, dD, dD0, ws, cxw, cW, cxd;
  DBuf<double> xch;          // cooperative kernel: [ws | Wbuf | xd | xw | Qp | Qf] in ONE allocation (one L2 access-policy window covers it)
  size_t xch_q_off = 0;
  const double *xch_q_for = nullptr;   // xch.p the copies of the quadratic blocks behind the exchange buffers were made for (null: stale)
  size_t l2_persist_max = 0, l2_window_max = 0;
  DBuf<int> cidle;
  DBuf<int> dST, dIT, queue;
  DBuf<double> dPXL, dPXU, dCB;   // bnlp_solve_children: parent bounds and the one changed bound of each child
  DBuf<int> dCP, dCC, dCW;
  DBuf<long long> stats;
  // eval buffers
  DBuf<double> eX, eLAM, eF, eG, eGF, eJ, eH, eX2, eDU, eU, eTH, oVal, oLB, oUB;
  DBuf<int> oRS, oCI, oSR;
  DevModel dm{};
};

}  // namespace

struct bnlp_engine {
  const bnlp_problem *prob = nullptr;
  HostModel M;
  std::vector<Device> devs;
  bnlp_options opt{};
  int R = 0, E = 0;
  int nq = 0;
  long long launches = 0;
  long long last_iters = 0, last_fact = 0;
  double last_ms = 0.0;
  bool ipm_supported = true;
  std::string ipm_unsupported_why;
  int h_in_smem = 1;
  int tm_kc = 16;
  size_t smem_bytes = 0;
  bool wide_terms = false;   // some element term couples more than BNLP_MAX_TERM_VARS variables: evaluation and the IPM run the wide build (namespace bnlpw)
  bool coop_ok = false;
  bool coop256 = false;   // small model: 256-thread CTAs, two per SM
  size_t coop_smem_bytes = 0;
  int nqpad = 0, ldd = 0;
  size_t l2_set_aside = 0;   // bytes of L2 reserved for persisting lines (cudaLimitPersistingL2CacheSize), device 0's view
  size_t ws_doubles = 0;
  int nnz_jac = 0, nnz_hes = 0, m_nl = 0;
  bnlp_engine *tqp_engine = nullptr;   // cached engine of the QP strong-branching surrogate (same shape at every node)
};

extern "C" bnlp_problem *bnlp_problem_from_tape(const bnlp_tape_desc *desc, char *err, size_t errlen) {
  auto *p = new bnlp_problem();
  std::string e;
  if (!model_from_desc(desc, p->M, e)) {
    if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
    set_err(e);
    delete p;
    return nullptr;
  }
  return p;
}

extern "C" bnlp_problem *bnlp_problem_from_nl(const char *nl_text, size_t len, char *err, size_t errlen) {
  auto *p = new bnlp_problem();
  std::string e;
  if (!parse_nl(nl_text, len, p->M, e)) {
    if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
    set_err(e);
    delete p;
    return nullptr;
  }
  return p;
}

extern "C" bnlp_problem *bnlp_problem_fp(const bnlp_problem *base, const bnlp_fp_desc *desc, char *err, size_t errlen) {
  std::string e;
  if (!base || !desc) e = "null argument";
  auto *p = new bnlp_problem();
  if (e.empty() && model_fp_variant(base->M, *desc, p->M, e)) return p;
  if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
  set_err(e);
  delete p;
  return nullptr;
}

extern "C" bnlp_problem *bnlp_problem_cut_strengthening(const bnlp_problem *base, int constr_index, char *err, size_t errlen) {
  std::string e;
  if (!base) e = "null argument";
  auto *p = new bnlp_problem();
  if (e.empty() && model_cs_variant(base->M, constr_index, p->M, e)) return p;
  if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
  set_err(e);
  delete p;
  return nullptr;
}

extern "C" int bnlp_problem_branching_info(const bnlp_problem *p, int *priorities, int *directions, double *up_pseudo, double *down_pseudo) {
  if (!p) return -1;
  const HostModel &M = p->M;
  int mask = 0;
  if (!M.br_priority.empty()) {
    mask |= 1;
    if (priorities) for (int i = 0; i < M.n; ++i) priorities[i] = -M.br_priority[i] + 9999;   // BonAmplTMINLP.cpp:231-236
  }
  if (!M.br_direction.empty()) {
    mask |= 2;
    if (directions) std::copy(M.br_direction.begin(), M.br_direction.end(), directions);
  }
  const std::vector<double> &up = !M.br_up_ps.empty() ? M.br_up_ps : M.br_down_ps;     // one given: used for both (:240-241)
  const std::vector<double> &dn = !M.br_down_ps.empty() ? M.br_down_ps : M.br_up_ps;
  if (!up.empty()) {
    mask |= 4 | 8;
    if (up_pseudo) std::copy(up.begin(), up.end(), up_pseudo);
    if (down_pseudo) std::copy(dn.begin(), dn.end(), down_pseudo);
  }
  return mask;
}

extern "C" int bnlp_problem_sos_info(const bnlp_problem *p, int *num, int *num_nz, char *types, int *priorities, int *starts,
                                     int *indices, double *weights) {
  if (!p || !num || !num_nz) return -1;
  const HostModel &M = p->M;
  *num = 0; *num_nz = 0;
  if (M.sos_no.empty()) return 0;
  // sets = the distinct non-zero .sosno values, in increasing order of the value; members by increasing .ref (ties and
  // a missing .ref: by variable index)
  std::vector<std::tuple<int, double, int>> mem;     // (set number, weight, variable)
  for (int j = 0; j < M.n; ++j)
    if (M.sos_no[j] != 0) mem.push_back({M.sos_no[j], M.sos_ref.empty() ? (double)j : M.sos_ref[j], j});
  std::sort(mem.begin(), mem.end());
  int ns = 0;
  for (size_t e = 0; e < mem.size(); ++e) {
    const bool first = e == 0 || std::get<0>(mem[e]) != std::get<0>(mem[e - 1]);
    if (first) {
      if (starts) starts[ns] = (int)e;
      if (types) types[ns] = std::get<0>(mem[e]) > 0 ? 1 : 2;
      if (priorities) priorities[ns] = 0;
      ++ns;
    }
    if (indices) indices[e] = std::get<2>(mem[e]);
    if (weights) weights[e] = std::get<1>(mem[e]);
    // a set inherits the best (largest) .priority of its members, stored like the variables' (9999 - value)
    if (priorities && !M.br_priority.empty()) priorities[ns - 1] = std::max(priorities[ns - 1], M.br_priority[std::get<2>(mem[e])]);
  }
  if (starts) starts[ns] = (int)mem.size();
  if (priorities) for (int s = 0; s < ns; ++s) priorities[s] = 9999 - priorities[s];
  if (weights && starts) {
    // ampl_utils::sos_kludge (src/Interfaces/Ampl/sos_kludge.cpp:29-45): reference weights made strictly increasing
    for (int s = 0; s < ns; ++s) {
      double t = weights[starts[s]];
      for (int e = starts[s] + 1; e < starts[s + 1]; ++e) {
        double t1 = weights[e];
        t += 1e-10;
        if (t1 <= t) weights[e] = t1 = t + 1e-10;
        t = t1;
      }
    }
  }
  *num = ns; *num_nz = (int)mem.size();
  return 0;
}

extern "C" void bnlp_problem_free(bnlp_problem *p) { delete p; }

extern "C" int bnlp_problem_sizes(const bnlp_problem *p, int *n, int *m, int *nnz_jac, int *nnz_hess) {
  if (!p) return -1;
  if (n) *n = p->M.n;
  if (m) *m = p->M.m;
  if (nnz_jac) *nnz_jac = (int)p->M.jac_row.size();
  if (nnz_hess) *nnz_hess = (int)p->M.hes_row.size();
  return 0;
}

extern "C" int bnlp_problem_info(const bnlp_problem *p, int *jac_irow, int *jac_jcol, int *hes_irow, int *hes_jcol,
                                 int *var_type, int *con_nonlinear, double *xl, double *xu, double *gl, double *gu,
                                 double *x0) {
  if (!p) return -1;
  const HostModel &M = p->M;
  auto cp = [](auto *dst, const auto &v) { if (dst && !v.empty()) std::copy(v.begin(), v.end(), dst); };
  cp(jac_irow, M.jac_row); cp(jac_jcol, M.jac_col); cp(hes_irow, M.hes_row); cp(hes_jcol, M.hes_col);
  cp(var_type, M.vtype); cp(con_nonlinear, M.con_nonlinear);
  cp(xl, M.xl); cp(xu, M.xu); cp(gl, M.gl); cp(gu, M.gu); cp(x0, M.x0);
  return 0;
}

static size_t ipm_smem_doubles(int n, int m, int R, int E, int h_in_smem, int n_lr = 0, int tm_kc = IPM_KC) {
  const int N = n + 1;
  const int npad = (N + 2 + 3) & ~3;
  size_t d = 5 * IPM_MAXE + IPM_MAXE * IPM_MAXE;
  if (h_in_smem) d += (size_t)(((N + 2 + E) * (N + 3 + E) / 2) + 1) & ~(size_t)1;
  d += std::max((size_t)(IPM_KC + 1) * npad, h_in_smem ? (size_t)0 : (size_t)tm_kc * ((((size_t)N + 7) & ~(size_t)7) + 4));
  d += (size_t)7 * npad;
  const int Rp = (R + 1) & ~1, Kp = (m + 2) & ~1;
  d += (size_t)10 * Rp + (size_t)8 * Kp + (size_t)IPM_NW * 12;
  d += (size_t)(((N + 1) & ~1) * 2 + 2 * Kp) / 2 + 2;   // flg, fixlist, qblk, qfn (ints)
  if (!h_in_smem) d += 2 * (size_t)((N + 7) & ~7) + 2 + 32 + (size_t)n_lr;  // tile-major H in global memory: the two right-hand sides live here (+ chunk signs, low-rank products)
  return d;
}

static size_t ipm_smem_doubles_coop(int n, int m, int R) {
  const int N = n + 1;
  const int npad = (N + 2 + 3) & ~3;
  const int lds = ((N + 2 + 7) & ~7) + 4;
  size_t d = (size_t)(((N + 2) * (N + 3) / 2) + 1) & ~(size_t)1;
  d += (size_t)(IPM_KC + 1) * lds;
  d += (size_t)7 * npad;
  const int Rp = (R + 1) & ~1, Kp = (m + 2) & ~1;
  d += (size_t)10 * Rp + (size_t)6 * Kp + (size_t)IPM_NW * 12;
  d += (size_t)(((N + 1) & ~1) * 2 + 2 * Kp + Rp) / 2 + 2;   // flg, fixlist, qblk, qfn, drow (ints)
  return d;
}

template <class T>
static int upload(DBuf<T> &b, const std::vector<T> &v) {
  if (b.reserve(v.size())) return -1;
  if (!v.empty() && cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  return 0;
}

int build_eval_tape(const HostModel &M, bool densified, std::vector<int> &ints, std::vector<double> &dbls,
                    EvalModelHostOffsets &off, const std::vector<int> &qidx, int tv_cap);

namespace {
// IPM view of a model: dense A, one packed quadratic block per function (device strides), one-sided rows, equality rows
struct IpmView {
  std::vector<double> A, QpD, Qf, row_sgn, row_bnd, eq_bnd;
  std::vector<int> qidx, row_fn, fn_up, fn_lo, eq_fn, fn_eq;
  int nq = 0, nE = 0;
  size_t ldp = 0, ldq = 0;
  // sparse linear rows of J' Sigma J (DevModel::sp_*): everything else is in dense_rows
  std::vector<int> dense_rows, sp_seg_start, sp_seg_i, sp_seg_k, sp_ent_row;
  std::vector<double> sp_ent_p;
  // the same sparse rows in the layout the cooperative kernel reads (DevModel::spc_* / sph_*)
  std::vector<int> sp_rows, spc_row;
  std::vector<double> spc_val, sph_rec;
  int spc_slots = 0, spc_ld = 0, sph_slots = 0;
  // low-rank quadratic blocks (DevModel::lrV): factor rows per function, CSR over functions 0 .. m
  std::vector<int> fn_lr_start;
  std::vector<double> lrV;
  int n_lr = 0;
};

// Pivoted Cholesky of a symmetric block given as packed lower triangle: Q ~= sum_r v_r v_r' with at most max_rank rows.
// Returns the rank, or -1 if the block is not positive semidefinite of rank <= max_rank to the relative tolerance.
int pivoted_cholesky(const double *Qp, int n, int max_rank, double rel_tol, std::vector<double> &rows) {
  auto q = [&](int i, int j) { return i >= j ? Qp[(size_t)i * (i + 1) / 2 + j] : Qp[(size_t)j * (j + 1) / 2 + i]; };
  std::vector<double> d(n);
  double dmax0 = 0.0;
  for (int i = 0; i < n; ++i) { d[i] = q(i, i); dmax0 = std::max(dmax0, std::fabs(d[i])); }
  rows.clear();
  if (dmax0 == 0.0) return 0;
  int rank = 0;
  for (;;) {
    int piv = 0;
    double dm = -1e300, dmin = 1e300;
    for (int i = 0; i < n; ++i) { if (d[i] > dm) { dm = d[i]; piv = i; } dmin = std::min(dmin, d[i]); }
    if (dmin < -rel_tol * dmax0 * 10) return -1;          // indefinite
    if (dm <= rel_tol * dmax0) break;                      // residual diagonal gone: done (a PSD residual is then gone everywhere)
    if (rank == max_rank) return -1;
    rows.resize((size_t)(rank + 1) * n);
    double *v = rows.data() + (size_t)rank * n;
    const double inv = 1.0 / std::sqrt(dm);
    for (int i = 0; i < n; ++i) {
      double a = q(i, piv);
      for (int r = 0; r < rank; ++r) a -= rows[(size_t)r * n + i] * rows[(size_t)r * n + piv];
      v[i] = a * inv;
    }
    for (int i = 0; i < n; ++i) d[i] -= v[i] * v[i];
    d[piv] = 0.0;
    ++rank;
  }
  return rank;
}
void build_ipm_view(const HostModel &M, IpmView &V) {
  const int n = M.n, m = M.m;
  const size_t P = (size_t)n * (n + 1) / 2;
  V.A.assign((size_t)(m + 1) * n, 0.0);
  for (int k = 0; k <= m; ++k)
    for (int e = M.lin_start[k]; e < M.lin_start[k + 1]; ++e) V.A[(size_t)k * n + M.lin_col[e]] += M.lin_val[e];
  V.qidx.assign(m + 1, -1);
  std::vector<double> Qp;
  auto block_of = [&](int fn) {
    if (V.qidx[fn] < 0) { V.qidx[fn] = (int)(Qp.size() / P); Qp.resize(Qp.size() + P, 0.0); }
    return Qp.data() + (size_t)V.qidx[fn] * P;
  };
  for (size_t q = 0; q < M.qd_fn.size(); ++q) {
    double *dst = block_of(M.qd_fn[q]);
    const double *src = M.qd.data() + q * P;
    for (size_t e = 0; e < P; ++e) dst[e] += src[e];
  }
  const bool densify = n <= 32 * IPM_MAXC;   // small models: every quadratic part becomes a dense block
  for (size_t e = 0; densify && e < M.qs_fn.size(); ++e) {
    double *dst = block_of(M.qs_fn[e]);
    dst[(size_t)M.qs_i[e] * (M.qs_i[e] + 1) / 2 + M.qs_j[e]] += M.qs_v[e];
  }
  V.nq = (int)(Qp.size() / P);
  // device strides: 128-byte aligned packed blocks and full-matrix rows (every request is a whole cache line)
  V.ldp = (P + 15) & ~(size_t)15; V.ldq = ((size_t)n + 15) & ~(size_t)15;
  V.Qf.assign((size_t)V.nq * n * V.ldq, 0.0); V.QpD.assign((size_t)V.nq * V.ldp, 0.0);
  for (int q = 0; q < V.nq; ++q) {
    std::copy(Qp.begin() + (size_t)q * P, Qp.begin() + (size_t)(q + 1) * P, V.QpD.begin() + (size_t)q * V.ldp);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        const double v = Qp[(size_t)q * P + (size_t)i * (i + 1) / 2 + j];
        V.Qf[((size_t)q * n + i) * V.ldq + j] = v;
        V.Qf[((size_t)q * n + j) * V.ldq + i] = v;
      }
  }
  V.fn_up.assign(m + 1, -1); V.fn_lo.assign(m + 1, -1); V.fn_eq.assign(m + 1, -1);
  V.nE = 0;
  for (int i = 0; i < m; ++i) {
    const bool lo = M.gl[i] > -kInf, hi = M.gu[i] < kInf;
    if (lo && hi && M.gl[i] == M.gu[i]) { V.fn_eq[i + 1] = V.nE++; V.eq_fn.push_back(i + 1); V.eq_bnd.push_back(M.gl[i]); continue; }
    if (hi) { V.fn_up[i + 1] = (int)V.row_fn.size(); V.row_fn.push_back(i + 1); V.row_sgn.push_back(1.0); V.row_bnd.push_back(M.gu[i]); }
    if (lo) { V.fn_lo[i + 1] = (int)V.row_fn.size(); V.row_fn.push_back(i + 1); V.row_sgn.push_back(-1.0); V.row_bnd.push_back(-M.gl[i]); }
  }
  // ---- sparse linear rows: a function without quadratic part and element terms whose row of A has few nonzeros
  std::vector<char> has_term(m + 1, 0);
  for (const Term &t : M.terms) has_term[t.fn] = 1;
  struct Ent { int i, k, row; double p; };
  std::vector<Ent> ents;
  // Which rows are worth it: a staged dense row costs ~1.2 k cycles of its rank-16 DMMA update at n = 200 (cooperative kernel,
  // H in shared memory), a sparse row ~1 cycle per touched entry (z (z + 1) / 2 of them for z nonzeros), so models that
  // qualify for the cooperative kernel take rows of up to n / 5 nonzeros (C3: the 40 linear rows with ~20 of 200).  The
  // general kernel, whose H may live in global memory, keeps the short-row limit.
  const bool coop_candidate = n <= 256 && M.terms.empty() && V.nE == 0 && V.nq > 0;
  int sparse_max = coop_candidate ? std::max(IPM_SPARSE_ROW, n / 5) : IPM_SPARSE_ROW;
  if (const char *e = getenv("BNLP_SPARSE_ROW_MAX")) sparse_max = atoi(e);
  for (int r = 0; r < (int)V.row_fn.size(); ++r) {
    const int fn = V.row_fn[r];
    std::vector<int> cols;
    for (int j = 0; j < n; ++j) if (V.A[(size_t)fn * n + j] != 0.0) cols.push_back(j);
    const bool sparse = V.qidx[fn] < 0 && !has_term[fn] && (int)cols.size() <= sparse_max && n > 64;
    if (!sparse) { V.dense_rows.push_back(r); continue; }
    for (size_t a = 0; a < cols.size(); ++a) {
      const double va = V.A[(size_t)fn * n + cols[a]];
      for (size_t b = 0; b <= a; ++b) ents.push_back({cols[a], cols[b], r, va * V.A[(size_t)fn * n + cols[b]]});   // kind 0
      ents.push_back({n, cols[a], r | (1 << 28), va});                                                           // kind 1
    }
    ents.push_back({n, n, r | (2 << 28), 1.0});                                                                   // kind 2
  }
  std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.i != b.i ? a.i < b.i : a.k < b.k; });
  V.sp_seg_start.push_back(0);
  for (size_t e = 0; e < ents.size(); ++e) {
    if (e == 0 || ents[e].i != ents[e - 1].i || ents[e].k != ents[e - 1].k) {
      if (e) V.sp_seg_start.push_back((int)e);
      V.sp_seg_i.push_back(ents[e].i); V.sp_seg_k.push_back(ents[e].k);
    }
    V.sp_ent_row.push_back(ents[e].row); V.sp_ent_p.push_back(ents[e].p);
  }
  if (!ents.empty()) V.sp_seg_start.push_back((int)ents.size());
  // ---- the cooperative kernel's view of the same rows.  Walking the segment lists above costs three DEPENDENT L2 round
  // trips per touched entry (segment -> entry -> row data); measured at C3 that was 84 k cycles per iteration, twice what
  // the two staged chunks it replaced had cost.  So: (1) the transposed lists are slot-major arrays (slot s of column j at
  // [s * spc_ld + j]: consecutive threads read consecutive addresses, the loads of all slots are independent); (2) the
  // touched entries of H are dealt to 512 virtual threads as self-contained 16-byte records {packed index, row | last
  // << 30, a_i a_k}, slot-major again ([s * 512 + v]): a thread streams its records with several loads in flight and adds
  // a run of records that share a target in a fixed order before it touches H once.  The elastic's row is not listed:
  // its entries are sums the column pass forms on the side.
  V.sp_rows.clear(); V.spc_row.clear(); V.spc_val.clear(); V.sph_rec.clear();
  V.spc_slots = V.spc_ld = V.sph_slots = 0;
  {
    std::vector<char> is_dense(V.row_fn.size(), 0);
    for (int r : V.dense_rows) is_dense[r] = 1;
    for (int r = 0; r < (int)V.row_fn.size(); ++r) if (!is_dense[r]) V.sp_rows.push_back(r);
  }
  if (!V.sp_rows.empty()) {
    std::vector<std::vector<std::pair<int, double>>> col(n);
    for (int r : V.sp_rows) {
      const int fn = V.row_fn[r];
      for (int j = 0; j < n; ++j) if (V.A[(size_t)fn * n + j] != 0.0) col[j].push_back({r, V.A[(size_t)fn * n + j]});
    }
    for (int j = 0; j < n; ++j) V.spc_slots = std::max(V.spc_slots, (int)col[j].size());
    V.spc_ld = (n + 31) & ~31;
    V.spc_row.assign((size_t)V.spc_slots * V.spc_ld, -1);
    V.spc_val.assign((size_t)V.spc_slots * V.spc_ld, 0.0);
    for (int j = 0; j < n; ++j)
      for (size_t e = 0; e < col[j].size(); ++e) { V.spc_row[e * V.spc_ld + j] = col[j][e].first; V.spc_val[e * V.spc_ld + j] = col[j][e].second; }
    // records: whole segments (kind 0 only) to the least loaded virtual thread, longest first
    const int VT = 512;
    struct Seg { int start, len, tgt; };
    std::vector<Seg> segs;
    for (size_t sgi = 0; sgi < V.sp_seg_i.size(); ++sgi) {
      if (V.sp_seg_i[sgi] >= n) continue;
      segs.push_back({V.sp_seg_start[sgi], V.sp_seg_start[sgi + 1] - V.sp_seg_start[sgi], V.sp_seg_i[sgi] * (V.sp_seg_i[sgi] + 1) / 2 + V.sp_seg_k[sgi]});
    }
    std::stable_sort(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.len > b.len; });
    std::vector<std::vector<std::pair<long long, double>>> th(VT);   // (tgt | row << 32 | last << 62, p)
    size_t nxt = 0;
    for (const Seg &sg : segs) {
      // longest first + round robin over threads sorted by load: with the lengths sorted a plain cyclic deal is balanced to
      // within one segment
      std::vector<std::pair<long long, double>> &T = th[nxt];
      nxt = (nxt + 1) % VT;
      for (int e = 0; e < sg.len; ++e) {
        const long long row = V.sp_ent_row[sg.start + e] & 0xfffffff;
        const long long last = (e == sg.len - 1) ? 1 : 0;
        T.push_back({(long long)sg.tgt | (row << 32) | (last << 62), V.sp_ent_p[sg.start + e]});
      }
    }
    for (int v = 0; v < VT; ++v) V.sph_slots = std::max(V.sph_slots, (int)th[v].size());
    V.sph_rec.assign((size_t)V.sph_slots * VT * 2, 0.0);
    for (int v = 0; v < VT; ++v)
      for (int sl = 0; sl < V.sph_slots; ++sl) {
        long long key = 0;            // padding: target 0, row 0, coefficient 0, not last: adds nothing and is never flushed
        double pv = 0.0;
        if (sl < (int)th[v].size()) { key = th[v][sl].first; pv = th[v][sl].second; }
        double kd;
        std::memcpy(&kd, &key, 8);
        V.sph_rec[((size_t)sl * VT + v) * 2] = kd;
        V.sph_rec[((size_t)sl * VT + v) * 2 + 1] = pv;
      }
  }
  // ---- low-rank quadratic blocks (used by the large-model path only; the dense copies stay for evaluation / OA)
  V.fn_lr_start.assign(m + 2, 0);
  V.n_lr = 0;
  if (n > 256 && !getenv("BNLP_NO_LOWRANK")) {
    std::vector<double> rows;
    for (int k = 0; k <= m; ++k) {
      V.fn_lr_start[k] = V.n_lr;
      if (V.qidx[k] < 0) continue;
      const int rk = pivoted_cholesky(Qp.data() + (size_t)V.qidx[k] * P, n, IPM_LR_MAX, 1e-14, rows);
      if (rk <= 0) continue;
      V.lrV.insert(V.lrV.end(), rows.begin(), rows.begin() + (size_t)rk * n);
      V.n_lr += rk;
    }
  }
  V.fn_lr_start[m + 1] = V.n_lr;
  for (int k = m; k >= 0; --k) if (V.fn_lr_start[k] > V.fn_lr_start[k + 1]) V.fn_lr_start[k] = V.fn_lr_start[k + 1];
}
}  // namespace

extern "C" bnlp_engine *bnlp_engine_create(const bnlp_problem *prob, const int *device_ids, int ndev,
                                           const bnlp_options *opt) {
  if (!prob) { set_err("null problem"); return nullptr; }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_err("bnlp_engine_create: no CUDA device available (there is no CPU fallback)");
    return nullptr;
  }
  auto *E = new bnlp_engine();
  E->prob = prob;
  E->M = prob->M;
  if (opt) E->opt = *opt;
  if (E->opt.tol <= 0) E->opt.tol = 1e-8;
  if (E->opt.max_iter <= 0) E->opt.max_iter = 3000;
  const HostModel &M = E->M;
  const int n = M.n, m = M.m;
  const size_t P = (size_t)n * (n + 1) / 2;
  E->nnz_jac = (int)M.jac_row.size();
  E->nnz_hes = (int)M.hes_row.size();
  for (int i = 0; i < m; ++i) E->m_nl += M.con_nonlinear[i];

  // ---- IPM view of the model: dense A, one packed quadratic block per function, one-sided rows
  IpmView V;
  build_ipm_view(M, V);
  std::vector<double> &A = V.A, &QpD = V.QpD, &Qf = V.Qf, &row_sgn = V.row_sgn, &row_bnd = V.row_bnd, &eq_bnd = V.eq_bnd;
  std::vector<int> &qidx = V.qidx, &row_fn = V.row_fn, &fn_up = V.fn_up, &fn_lo = V.fn_lo, &eq_fn = V.eq_fn, &fn_eq = V.fn_eq;
  const size_t ldp = V.ldp, ldq = V.ldq;
  const bool densify = n <= 32 * IPM_MAXC;
  E->nq = V.nq;
  E->E = V.nE;
  E->R = (int)row_fn.size();
  if (E->E > IPM_MAXE) { E->ipm_supported = false; E->ipm_unsupported_why = "more than 8 equality rows are not supported by the batched IPM yet"; }
  if (n > 32 * IPM_MAXC) { E->ipm_supported = false; E->ipm_unsupported_why = "n > 512 is not supported by the batched IPM yet"; }

  std::vector<int> ids;
  if (ndev <= 0 || !device_ids) ids.push_back(0);
  else ids.assign(device_ids, device_ids + ndev);
  E->devs.resize(ids.size());

  std::vector<int> e_ints;
  std::vector<double> e_dbls;
  EvalModelHostOffsets eoff;
  {
    std::string werr;
    const int wc = term_width_class(M, werr);
    if (wc < 0) { set_err("bnlp_engine_create: " + werr); delete E; return nullptr; }
    E->wide_terms = wc == 1;
  }
  const int tv_cap = E->wide_terms ? BNLP_MAX_WIDE_TERM_VARS : BNLP_MAX_TERM_VARS;
  build_eval_tape(M, densify, e_ints, e_dbls, eoff, qidx, tv_cap);

  for (size_t d = 0; d < ids.size(); ++d) {
    Device &D = E->devs[d];
    D.id = ids[d];
    if (D.id < 0 || D.id >= count || cudaSetDevice(D.id) != cudaSuccess) {
      set_err("bad device id");
      delete E;
      return nullptr;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, D.id);
    D.sms = prop.multiProcessorCount;
    D.smem_optin = prop.sharedMemPerBlockOptin;
    D.l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    D.l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking);
    cudaEventCreate(&D.ev0);
    cudaEventCreate(&D.ev1);
    int bad = 0;
    bad |= upload(D.c, M.c); bad |= upload(D.A, A); bad |= upload(D.qidx, qidx); bad |= upload(D.Qp, QpD); bad |= upload(D.Qf, Qf);
    bad |= upload(D.row_fn, row_fn); bad |= upload(D.row_sgn, row_sgn); bad |= upload(D.row_bnd, row_bnd);
    bad |= upload(D.fn_row_up, fn_up); bad |= upload(D.fn_row_lo, fn_lo);
    bad |= upload(D.eq_fn, eq_fn); bad |= upload(D.fn_eq, fn_eq); bad |= upload(D.eq_bnd, eq_bnd);
    bad |= upload(D.dense_rows, V.dense_rows); bad |= upload(D.sp_seg_start, V.sp_seg_start); bad |= upload(D.sp_seg_i, V.sp_seg_i);
    bad |= upload(D.sp_seg_k, V.sp_seg_k); bad |= upload(D.sp_ent_row, V.sp_ent_row); bad |= upload(D.sp_ent_p, V.sp_ent_p);
    bad |= upload(D.sp_rows, V.sp_rows); bad |= upload(D.spc_row, V.spc_row); bad |= upload(D.spc_val, V.spc_val); bad |= upload(D.sph_rec, V.sph_rec);
    bad |= upload(D.fn_lr_start, V.fn_lr_start); bad |= upload(D.lrV, V.lrV);
    bad |= upload(D.gl, M.gl); bad |= upload(D.gu, M.gu); bad |= upload(D.xl, M.xl); bad |= upload(D.xu, M.xu);
    bad |= upload(D.e_int, e_ints); bad |= upload(D.e_dbl, e_dbls);
    if (bad) { set_err("device upload failed"); delete E; return nullptr; }
    D.ev = make_eval_model_dev(eoff, D.e_int.p, D.e_dbl.p, D.gl.p, D.gu.p);
    D.ev.qidx = D.qidx.p; D.ev.Qp = D.Qp.p; D.ev.Qf = D.Qf.p; D.ev.ldp = (int)ldp; D.ev.ldq = (int)ldq;
    DevModel &dm = D.dm;
    dm.n = n; dm.m = m; dm.N = n + 1; dm.R = E->R; dm.nq = E->nq; dm.P = (int)P;
    dm.npad = (dm.N + 2 + 3) & ~3;
    dm.ldp = (int)ldp; dm.ldq = (int)ldq;
    dm.c = D.c.p; dm.A = D.A.p; dm.qidx = D.qidx.p; dm.Qp = D.Qp.p; dm.Qf = D.Qf.p;
    dm.row_fn = D.row_fn.p; dm.row_sgn = D.row_sgn.p; dm.row_bnd = D.row_bnd.p;
    dm.fn_row_up = D.fn_row_up.p; dm.fn_row_lo = D.fn_row_lo.p; dm.gl = D.gl.p; dm.gu = D.gu.p;
    dm.n_terms = D.ev.n_terms; dm.term_fn = D.ev.term_fn; dm.term_var_start = D.ev.term_var_start; dm.term_var = D.ev.term_var;
    dm.fn_term_start = D.ev.fn_term_start; dm.fn_term_list = D.ev.fn_term_list; dm.hseg_start = D.ev.hseg_start;
    dm.hseg_target = D.ev.hseg_target; dm.hent_term = D.ev.hent_term; dm.hent_lidx = D.ev.hent_lidx; dm.n_hseg = D.ev.n_hseg;
    dm.term_op_start = D.ev.term_op_start; dm.term_op = D.ev.term_op; dm.term_iarg = D.ev.term_iarg; dm.term_darg = D.ev.term_darg;
    dm.E = E->E; dm.eq_fn = D.eq_fn.p; dm.eq_bnd = D.eq_bnd.p; dm.fn_eq = D.fn_eq.p;
    dm.Rd = (int)V.dense_rows.size(); dm.dense_rows = D.dense_rows.p;
    dm.sp_nseg = (int)V.sp_seg_i.size(); dm.sp_seg_start = D.sp_seg_start.p; dm.sp_seg_i = D.sp_seg_i.p; dm.sp_seg_k = D.sp_seg_k.p;
    dm.sp_ent_row = D.sp_ent_row.p; dm.sp_ent_p = D.sp_ent_p.p;
    dm.Rs = (int)V.sp_rows.size(); dm.sp_rows = D.sp_rows.p; dm.spc_slots = V.spc_slots; dm.spc_ld = V.spc_ld; dm.spc_row = D.spc_row.p;
    dm.spc_val = D.spc_val.p; dm.sph_slots = V.sph_slots; dm.sph_rec = D.sph_rec.p;
    dm.n_lr = V.n_lr; dm.fn_lr_start = D.fn_lr_start.p; dm.lrV = D.lrV.p;
    if (d == 0) {
      const size_t need = ipm_smem_doubles(n, m, E->R, E->E, 1) * sizeof(double);
      E->h_in_smem = need + 1024 <= D.smem_optin ? 1 : 0;
      E->tm_kc = IPM_KC;
      if (!E->h_in_smem && ipm_smem_doubles(n, m, E->R, E->E, 0, V.n_lr, 32) * sizeof(double) + 1024 <= D.smem_optin && !getenv("BNLP_TM_KC16")) E->tm_kc = 32;
      E->smem_bytes = ipm_smem_doubles(n, m, E->R, E->E, E->h_in_smem, V.n_lr, E->tm_kc) * sizeof(double);
      if (E->smem_bytes + 1024 > D.smem_optin) { E->ipm_supported = false; E->ipm_unsupported_why = "iterate vectors do not fit shared memory"; }
      E->ws_doubles = (size_t)2 * (m + 1) * n + (size_t)M.terms.size() * (2 + tv_cap + tv_cap * (tv_cap + 1) / 2) + (E->h_in_smem ? 0 : (size_t)tm_size(tm_layout(n + 3 + E->E)) + 2);   // tile-major H (bnlp_dense_tm.cuh)
      E->ws_doubles = (E->ws_doubles + 1) & ~(size_t)1;
      // cooperative DMMA kernel: needs H in shared memory, n <= 256 and the group weights to fit the staging area
      E->nqpad = (std::max(E->nq, 1) + 3) & ~3;
      E->ldd = ((n + 15) & ~15) + 1;   // >= roundup16(n); stride = 1 (mod 16) doubles: conflict-free permuted-k fragments
      const int lds = ((n + 1 + 2 + 7) & ~7) + 4;
      E->coop_smem_bytes = ipm_smem_doubles_coop(n, m, E->R) * sizeof(double);
      int coopAttr = 0;
      cudaDeviceGetAttribute(&coopAttr, cudaDevAttrCooperativeLaunch, D.id);
      E->coop_ok = E->ipm_supported && coopAttr && n <= 256 && E->nq > 0 && E->E == 0 && M.terms.empty() && E->coop_smem_bytes + 1024 <= D.smem_optin &&
                   (size_t)COOP_RB * (E->nqpad + 4) <= (size_t)(IPM_KC + 1) * lds &&
                   (size_t)COOP_RB * E->ldd <= (size_t)(IPM_KC + 1) * lds + (size_t)(n + 3) * (n + 4) / 2 &&   // U-pass operands live in H + staging
                   !getenv("BNLP_NO_COOP");
    }
    // (the attribute is per kernel, not per engine: several engines with different needs may be alive, so opt in to the maximum)
    // (function attributes are per device: every device opts in; the variant is used only if all of them can)
    {
      const bool ok256 = E->coop_ok && n + 3 <= 128 && !getenv("BNLP_NO_COOP256") &&
                         bnlp_coop256_set_smem(D.smem_optin - 1024) == cudaSuccess && bnlp_coop256_max_ctas_per_sm(E->coop_smem_bytes) >= 2;
      E->coop256 = (d == 0) ? ok256 : (E->coop256 && ok256);
    }
    if (E->coop_ok && ipm_coop_set_smem(D.smem_optin - 1024) != cudaSuccess) {
      E->coop_ok = false;
      cudaGetLastError();
    }
    if (E->ipm_supported && (E->wide_terms ? bnlpw_ipm_set_smem(D.smem_optin - 1024) : ipm_general_set_smem(D.smem_optin - 1024)) != cudaSuccess) {
      set_err(std::string("cudaFuncSetAttribute(ipm_kernel): ") + cudaGetErrorString(cudaGetLastError()));
      delete E;
      return nullptr;
    }
    D.queue.reserve(1);
    D.stats.reserve(12);
  }
  return E;
}

// Same structure, new numbers: refresh the model data of an existing engine in place (used by the QP strong-branching
// surrogate, whose tape has the same shape at every node).  Returns 1 if reloaded, 0 if the structure differs.
int engine_reload_values(bnlp_engine *Q, const HostModel &M) {
  if (Q->M.n != M.n || Q->M.m != M.m || Q->M.terms.size() != M.terms.size() || !M.terms.empty()) return 0;
  IpmView V;
  build_ipm_view(M, V);
  if (V.nq != Q->nq || V.nE != Q->E || (int)V.row_fn.size() != Q->R) return 0;
  for (Device &D : Q->devs) {
    if (cudaSetDevice(D.id) != cudaSuccess) return 0;
    std::vector<int> old_q(M.m + 1), old_r(V.row_fn.size());
    if (cudaMemcpy(old_q.data(), D.qidx.p, old_q.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    if (!old_r.empty() && cudaMemcpy(old_r.data(), D.row_fn.p, old_r.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    if (old_q != V.qidx || old_r != V.row_fn) return 0;
    int bad = 0;
    bad |= upload(D.c, M.c); bad |= upload(D.A, V.A); bad |= upload(D.Qp, V.QpD); bad |= upload(D.Qf, V.Qf);
    D.xch_q_for = nullptr;      // the copies behind the exchange buffers are stale
    bad |= upload(D.row_bnd, V.row_bnd); bad |= upload(D.eq_bnd, V.eq_bnd);
    bad |= upload(D.gl, M.gl); bad |= upload(D.gu, M.gu); bad |= upload(D.xl, M.xl); bad |= upload(D.xu, M.xu);
    // the sparse-row lists depend on which coefficients of A are non-zero: they follow the new numbers
    bad |= upload(D.dense_rows, V.dense_rows); bad |= upload(D.sp_seg_start, V.sp_seg_start); bad |= upload(D.sp_seg_i, V.sp_seg_i);
    bad |= upload(D.sp_seg_k, V.sp_seg_k); bad |= upload(D.sp_ent_row, V.sp_ent_row); bad |= upload(D.sp_ent_p, V.sp_ent_p);
    bad |= upload(D.sp_rows, V.sp_rows); bad |= upload(D.spc_row, V.spc_row); bad |= upload(D.spc_val, V.spc_val); bad |= upload(D.sph_rec, V.sph_rec);
    if (bad) return 0;
    D.dm.A = D.A.p; D.dm.Qp = D.Qp.p; D.dm.Qf = D.Qf.p; D.dm.c = D.c.p; D.dm.row_bnd = D.row_bnd.p; D.dm.eq_bnd = D.eq_bnd.p;
    D.dm.gl = D.gl.p; D.dm.gu = D.gu.p;
    D.dm.Rd = (int)V.dense_rows.size(); D.dm.dense_rows = D.dense_rows.p;
    D.dm.sp_nseg = (int)V.sp_seg_i.size(); D.dm.sp_seg_start = D.sp_seg_start.p; D.dm.sp_seg_i = D.sp_seg_i.p; D.dm.sp_seg_k = D.sp_seg_k.p;
    D.dm.sp_ent_row = D.sp_ent_row.p; D.dm.sp_ent_p = D.sp_ent_p.p;
    D.dm.Rs = (int)V.sp_rows.size(); D.dm.sp_rows = D.sp_rows.p; D.dm.spc_slots = V.spc_slots; D.dm.spc_ld = V.spc_ld; D.dm.spc_row = D.spc_row.p;
    D.dm.spc_val = D.spc_val.p; D.dm.sph_slots = V.sph_slots; D.dm.sph_rec = D.sph_rec.p;
    if (upload(D.fn_lr_start, V.fn_lr_start) || upload(D.lrV, V.lrV)) return 0;
    D.dm.n_lr = V.n_lr; D.dm.fn_lr_start = D.fn_lr_start.p; D.dm.lrV = D.lrV.p;
  }
  Q->M = M;
  return 1;
}

extern "C" void bnlp_engine_free(bnlp_engine *E) {
  if (!E) return;
  if (E->tqp_engine) { bnlp_engine_free(E->tqp_engine); E->tqp_engine = nullptr; }
  for (Device &D : E->devs) {
    cudaSetDevice(D.id);
    for (auto *b : {&D.c, &D.A, &D.Qp, &D.Qf, &D.row_sgn, &D.row_bnd, &D.gl, &D.gu, &D.xl, &D.xu, &D.e_dbl, &D.dXL, &D.dXU,
                    &D.dX0, &D.dOBJ, &D.dX, &D.dG, &D.dD, &D.dD0, &D.ws, &D.cxw, &D.cW, &D.cxd, &D.xch, &D.eq_bnd, &D.eX, &D.eLAM, &D.eF, &D.eG, &D.eGF, &D.eJ, &D.eH,
                    &D.eX2, &D.eDU, &D.eU, &D.eTH, &D.oVal, &D.oLB, &D.oUB, &D.dPXL, &D.dPXU, &D.dCB})
      b->release();
    for (auto *b : {&D.qidx, &D.row_fn, &D.fn_row_up, &D.fn_row_lo, &D.e_int, &D.dST, &D.dIT, &D.queue, &D.cidle, &D.eq_fn, &D.fn_eq, &D.oRS, &D.oCI,
                    &D.oSR, &D.dCP, &D.dCC, &D.dCW, &D.dense_rows, &D.sp_seg_start, &D.sp_seg_i, &D.sp_seg_k, &D.sp_ent_row, &D.sp_rows, &D.spc_row})
      b->release();
    D.stats.release();
    for (int p = 0; p < 2; ++p) { D.oaVal[p].release(); D.oaLB[p].release(); D.oaUB[p].release(); D.oaCI[p].release(); D.oaSR[p].release(); D.oaRS[p].release();
      if (D.oa_comp[p]) cudaEventDestroy(D.oa_comp[p]); if (D.oa_copy[p]) cudaEventDestroy(D.oa_copy[p]); if (D.oa_k0[p]) cudaEventDestroy(D.oa_k0[p]); }
    if (D.stream2) cudaStreamDestroy(D.stream2);
    D.sp_ent_p.release(); D.lrV.release(); D.fn_lr_start.release(); D.spc_val.release(); D.sph_rec.release();
    if (D.stream) cudaStreamDestroy(D.stream);
    if (D.ev0) cudaEventDestroy(D.ev0);
    if (D.ev1) cudaEventDestroy(D.ev1);
  }
  delete E;
}

static int launch_ipm(bnlp_engine *E, Device &D, int B, const double *dXL, const double *dXU, const double *dX0,
                      int x0_per_child, double *dOBJ, int *dST, int *dIT, double *dX, double *dG, double *dD,
                      cudaStream_t stream, const double *dDUAL0 = nullptr, int dual0_per_child = 0) {
  if (!E->ipm_supported) { set_err("bnlp_solve_batch: " + E->ipm_unsupported_why); return -2; }
  if (B <= 0) return 0;
  int ctas = E->opt.ctas_per_device > 0 ? E->opt.ctas_per_device : D.sms;
  if (!E->coop_ok) ctas = std::min(ctas, B);   // the cooperative kernel keeps every SM for its shared passes
  // small models: two 256-thread CTAs per SM interleave two children's latency chains (+50 % at C2) — worth it only when
  // the batch is clearly larger than one child per SM (measured at C2, scripts/dev_256_threshold.py: 149 children 1.82 ms
  // with 512 threads vs 2.07 ms, 200 children 2.70 vs 2.16 ms); a strong-branching batch of one node keeps the
  // 512-thread CTAs
  static const int min_b256 = getenv("BNLP_COOP256_MIN_B") ? atoi(getenv("BNLP_COOP256_MIN_B")) : 0;   // (tuning aid)
  const bool use256 = E->coop256 && E->opt.ctas_per_device <= 0 && B >= (min_b256 > 0 ? min_b256 : D.sms + D.sms / 4);
  if (use256) ctas = 2 * D.sms;
  // cooperative kernel: the per-CTA workspace and the three exchange buffers of the shared passes are carved from ONE block,
  // which is marked PERSISTING in L2 for this launch: they are rewritten every iteration (W_b 162 KB + Q.dx 65 KB + gradients
  // per child and step at C3) and round 1's ncu showed half of that traffic going out to HBM (12 GB per 4096-child launch
  // against 36 MB of algorithmic bytes) although the whole working set is smaller than L2.
  const size_t ws_cnt = (size_t)ctas * E->ws_doubles;
  size_t w_off = 0, xd_off = 0, xw_off = 0, xch_cnt = 0, qp_off = 0, qf_off = 0;
  if (E->coop_ok) {
    auto up = [](size_t v) { return (v + 31) & ~(size_t)31; };   // 256-byte aligned pieces
    w_off = up(ws_cnt); xd_off = w_off + up((size_t)ctas * D.dm.ldp); xw_off = xd_off + up((size_t)ctas * E->ldd);
    xch_cnt = xw_off + up((size_t)ctas * E->nqpad);
    // the quadratic blocks ride behind the exchange buffers (copies of D.Qp / D.Qf, which the evaluation kernels keep using):
    // every step re-reads all of them, and beside 62 MB of persisting lines they were what the rest of L2 kept evicting
    // (4.7 GB of DRAM reads per 4096-child launch against 20 MB of blocks)
    static const bool q_in_window = getenv("BNLP_NO_Q_IN_WINDOW") == nullptr;
    if (q_in_window) {
      qp_off = xch_cnt; qf_off = qp_off + up((size_t)E->nq * D.dm.ldp);
      xch_cnt = qf_off + up((size_t)E->nq * E->M.n * D.dm.ldq);
    }
    if (D.xch.reserve(xch_cnt) || D.cidle.reserve(2)) { set_err("cooperative buffers: allocation failed"); return -1; }
    if (q_in_window && (D.xch_q_for != D.xch.p || D.xch_q_off != qp_off)) {
      CK(cudaMemcpyAsync(D.xch.p + qp_off, D.Qp.p, (size_t)E->nq * D.dm.ldp * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      CK(cudaMemcpyAsync(D.xch.p + qf_off, D.Qf.p, (size_t)E->nq * E->M.n * D.dm.ldq * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      D.xch_q_for = D.xch.p; D.xch_q_off = qp_off;
    }
  } else if (D.ws.reserve(ws_cnt)) { set_err("workspace allocation failed"); return -1; }
  CK(cudaMemsetAsync(D.queue.p, 0, sizeof(int), stream));
  CK(cudaMemsetAsync(D.stats.p, 0, 12 * sizeof(long long), stream));
  IpmArgs a{};
  a.B = B; a.XL = dXL; a.XU = dXU; a.X0 = dX0; a.x0_stride = x0_per_child ? E->M.n : 0;
  a.DUAL0 = dDUAL0; a.dual0_stride = dual0_per_child ? 2 * E->M.n + E->M.m : 0;
  a.OBJ = dOBJ; a.STATUS = dST; a.ITERS = dIT; a.X = dX; a.G = dG; a.DUALS = dD;
  a.queue = D.queue.p; a.ws = E->coop_ok ? D.xch.p : D.ws.p; a.ws_stride = E->ws_doubles; a.h_in_smem = E->h_in_smem; a.tm_kc = E->tm_kc; a.stats = D.stats.p;
  { static const int tma = getenv("BNLP_COOP_TMA") ? atoi(getenv("BNLP_COOP_TMA")) : 0; a.coop_tma = tma; }
  a.phase_cycles = getenv("BNLP_PHASE_TIMERS") ? D.stats.p + 2 : nullptr;
  a.o = IpmOpt();
  a.o.tol = E->opt.tol; a.o.max_iter = E->opt.max_iter;
  if (E->coop_ok) {
    CoopBufs cb{};
    CK(cudaMemsetAsync(D.cidle.p, 0, 2 * sizeof(int), stream));
    cb.xw = D.xch.p + xw_off; cb.Wbuf = D.xch.p + w_off; cb.xd = D.xch.p + xd_off; cb.idle_cnt = D.cidle.p; cb.nqpad = E->nqpad; cb.ldd = E->ldd;
    static const bool no_persist = getenv("BNLP_NO_L2_PERSIST") != nullptr;
    const size_t xbytes = xch_cnt * sizeof(double);
    bool windowed = false;
    if (!no_persist && D.l2_persist_max > 0 && D.l2_window_max > 0) {
      // (the set-aside is a device limit; the window is an attribute of the launching stream and is cleared right after the
      // launch has been enqueued, so a caller's stream is left as it was)
      const size_t want = std::min(xbytes, D.l2_persist_max);
      if (E->l2_set_aside < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) E->l2_set_aside = want;
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.base_ptr = D.xch.p;
      av.accessPolicyWindow.num_bytes = std::min(xbytes, D.l2_window_max);
      // Hit ratio: with EVERY access of the window persisting (ratio 1: 82.7 MB of the 126 MB L2 at C3) the DRAM traffic of a
      // 4096-child launch is lowest (2.4 GB) but the kernel is 5 % SLOWER than with no window at all — what is left of L2 for
      // everything else (code, spills, inputs) is then too small; 0.4 - 0.7 is the fastest setting measured (49.0 - 50.1 k against
      // 47.3 - 47.5 k relaxations/s at ratio 1 and 47.9 - 48.6 k without a window; 7.6 GB of DRAM traffic, 1.4 % of the HBM
      // bandwidth).  BNLP_L2_HIT_RATIO overrides.
      static const double hit_cap = getenv("BNLP_L2_HIT_RATIO") ? atof(getenv("BNLP_L2_HIT_RATIO")) : 0.6;
      av.accessPolicyWindow.hitRatio = (float)std::min(hit_cap, (double)E->l2_set_aside / (double)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = getenv("BNLP_L2_MISS_NORMAL") ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
      windowed = E->l2_set_aside > 0 && cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
      if (getenv("BNLP_L2_INFO"))
        fprintf(stderr, "[bnlp] L2 window %.1f MB (exchange %.1f MB + quadratic blocks %.1f MB), set-aside %.1f MB of max %.1f MB, max window %.1f MB, hit ratio %.2f, %s\n",
                av.accessPolicyWindow.num_bytes / 1e6, (qp_off ? qp_off : xch_cnt) * 8 / 1e6, qp_off ? (xch_cnt - qp_off) * 8 / 1e6 : 0.0,
                E->l2_set_aside / 1e6, D.l2_persist_max / 1e6, D.l2_window_max / 1e6, av.accessPolicyWindow.hitRatio, windowed ? "on" : "off");
      cudaGetLastError();
    }
    cudaError_t le;
    DevModel dmc = D.dm;
    if (qp_off) { dmc.Qp = D.xch.p + qp_off; dmc.Qf = D.xch.p + qf_off; }
    if (use256) le = bnlp_coop256_launch(a.phase_cycles != nullptr, ctas, E->coop_smem_bytes, stream, &dmc, &a, &cb);
    else le = ipm_coop_launch(a.phase_cycles != nullptr, ctas, E->coop_smem_bytes, stream, dmc, a, cb);
    if (windowed) {
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.num_bytes = 0;
      cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
    }
    CK(le);
  } else if (E->wide_terms) CK(bnlpw_ipm_launch(E->M.n > 256, ctas, E->smem_bytes, stream, &D.dm, &a));
  else CK(ipm_general_launch(E->M.n > 256, ctas, E->smem_bytes, stream, D.dm, a));
  E->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int bnlp_solve_batch_device(bnlp_engine *E, int dev_index, int B, const double *dXL, const double *dXU,
                                       const double *dX0, int x0_per_child, double *dOBJ, int *dSTATUS, int *dITERS,
                                       double *dX, double *dG, double *dDUALS, void *stream) {
  if (!E || dev_index < 0 || dev_index >= (int)E->devs.size()) { set_err("bad engine/device"); return -1; }
  Device &D = E->devs[dev_index];
  CK(cudaSetDevice(D.id));
  return launch_ipm(E, D, B, dXL, dXU, dX0, x0_per_child, dOBJ, dSTATUS, dITERS, dX, dG, dDUALS, (cudaStream_t)stream);
}

// Results of a launched batch back to the host, then wait and collect statistics (shared by bnlp_solve_batch and
// bnlp_solve_children).
static int collect_batch(bnlp_engine *E, int B, int per, double *OBJ, int *STATUS, int *ITERS, double *X, double *G, double *DUALS) {
  const int n = E->M.n, m = E->M.m, nd = (int)E->devs.size();
  // results back.  (Separate loop: a device-to-host copy into pageable memory blocks the host until the
  // device's kernel has finished, so issuing it inside the launch loop would run the devices one after the other.)
  for (int d = 0; d < nd; ++d) {
    const int b0 = d * per, nb = std::min(per, B - b0);
    if (nb <= 0) break;
    Device &D = E->devs[d];
    CK(cudaSetDevice(D.id));
    CK(cudaMemcpyAsync(OBJ + b0, D.dOBJ.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, D.stream));
    CK(cudaMemcpyAsync(STATUS + b0, D.dST.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, D.stream));
    CK(cudaMemcpyAsync(ITERS + b0, D.dIT.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, D.stream));
    if (X) CK(cudaMemcpyAsync(X + (size_t)b0 * n, D.dX.p, (size_t)nb * n * 8, cudaMemcpyDeviceToHost, D.stream));
    if (G && m) CK(cudaMemcpyAsync(G + (size_t)b0 * m, D.dG.p, (size_t)nb * m * 8, cudaMemcpyDeviceToHost, D.stream));
    if (DUALS) CK(cudaMemcpyAsync(DUALS + (size_t)b0 * (2 * n + m), D.dD.p, (size_t)nb * (2 * n + m) * 8, cudaMemcpyDeviceToHost, D.stream));
    CK(cudaEventRecord(D.ev1, D.stream));
  }
  // phase 2: wait and collect statistics
  E->last_iters = E->last_fact = 0;
  E->last_ms = 0.0;
  for (int d = 0; d < nd; ++d) {
    const int b0 = d * per;
    if (b0 >= B) break;
    Device &D = E->devs[d];
    CK(cudaSetDevice(D.id));
    CK(cudaStreamSynchronize(D.stream));
    long long st[12] = {0};
    CK(cudaMemcpy(st, D.stats.p, sizeof(st), cudaMemcpyDeviceToHost));
    if (getenv("BNLP_PHASE_TIMERS")) {
      static const char *nm0[10] = {"residuals", "W pass", "J'SJ", "cholesky", "backsolve", "direction", "U pass", "linesearch", "update", "init"};
      static const char *nm1[10] = {"top+J'SJ", "gridsync", "W coop", "cholesky", "solve+dir", "U coop", "linesearch", "update", "addW+fix", "idle/init"};
      const char **nm = E->coop_ok ? nm1 : nm0;
      long long tot = 0;
      for (int q = 0; q < 10; ++q) tot += st[2 + q];
      for (int q = 0; q < 10; ++q) fprintf(stderr, "[bnlp phase] %-10s %6.2f%%  %.0f cyc/iter\n", nm[q], 100.0 * st[2 + q] / (tot ? tot : 1), (double)st[2 + q] / (st[0] ? st[0] : 1));
    }
    E->last_iters += st[0]; E->last_fact += st[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, D.ev0, D.ev1);
    E->last_ms = std::max(E->last_ms, (double)ms);
  }
  return 0;
}



// child b = the bounds of parent cp[b] with one bound of column cc[b] replaced (cw[b] == 0: upper bound, the down branch;
// 1: lower bound, the up branch); X0 rows are copied per parent when x0_per_parent
__global__ void expand_children_kernel(int B, int n, const double *__restrict__ pXL, const double *__restrict__ pXU,
                                       const int *__restrict__ cp, const int *__restrict__ cc, const int *__restrict__ cw,
                                       const double *__restrict__ cb, double *__restrict__ XL, double *__restrict__ XU,
                                       const double *__restrict__ pX0, double *__restrict__ X0) {
  const size_t total = (size_t)B * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / n), j = (int)(e % n);
    const size_t src = (size_t)cp[b] * n + j;
    const bool hit = (j == cc[b]);
    XL[e] = (hit && cw[b] == 1) ? cb[b] : pXL[src];
    XU[e] = (hit && cw[b] == 0) ? cb[b] : pXU[src];
    if (X0) X0[e] = pX0[src];
  }
}

extern "C" int bnlp_solve_children(bnlp_engine *E, int P, const double *parent_XL, const double *parent_XU, int B,
                                   const int *child_parent, const int *child_col, const int *child_way,
                                   const double *child_bound, const double *X0, int x0_per_parent, double *OBJ, int *STATUS,
                                   int *ITERS, double *X, double *G, double *DUALS) {
  if (!E) { set_err("null engine"); return -1; }
  if (B <= 0) return 0;
  if (!E->ipm_supported) { set_err("bnlp_solve_children: " + E->ipm_unsupported_why); return -2; }
  const int n = E->M.n, m = E->M.m, nd = (int)E->devs.size();
  for (int b = 0; b < B; ++b)
    if (child_parent[b] < 0 || child_parent[b] >= P || child_col[b] < 0 || child_col[b] >= n || (child_way[b] | 1) != 1) {
      set_err("bnlp_solve_children: child " + std::to_string(b) + " out of range");
      return -1;
    }
  const int per = (B + nd - 1) / nd;
  for (int d = 0; d < nd; ++d) {
    const int b0 = d * per, nb = std::min(per, B - b0);
    if (nb <= 0) break;
    Device &D = E->devs[d];
    CK(cudaSetDevice(D.id));
    if (D.dXL.reserve((size_t)nb * n) || D.dXU.reserve((size_t)nb * n) || D.dX0.reserve((size_t)(x0_per_parent ? nb : 1) * n) ||
        D.dOBJ.reserve(nb) || D.dST.reserve(nb) || D.dIT.reserve(nb) || D.dX.reserve((size_t)nb * n) ||
        D.dG.reserve((size_t)nb * std::max(m, 1)) || D.dD.reserve((size_t)nb * (2 * n + m)) ||
        D.dPXL.reserve((size_t)P * n * (x0_per_parent ? 2 : 1)) || D.dPXU.reserve((size_t)P * n) || D.dCB.reserve(nb) ||
        D.dCP.reserve(nb) || D.dCC.reserve(nb) || D.dCW.reserve(nb)) {
      set_err("device allocation failed");
      return -1;
    }
    CK(cudaEventRecord(D.ev0, D.stream));
    CK(cudaMemcpyAsync(D.dPXL.p, parent_XL, (size_t)P * n * 8, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dPXU.p, parent_XU, (size_t)P * n * 8, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dCP.p, child_parent + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dCC.p, child_col + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dCW.p, child_way + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dCB.p, child_bound + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, D.stream));
    double *pX0 = nullptr;
    if (x0_per_parent) {   // parent start points ride behind the parent lower bounds
      pX0 = D.dPXL.p + (size_t)P * n;
      CK(cudaMemcpyAsync(pX0, X0, (size_t)P * n * 8, cudaMemcpyHostToDevice, D.stream));
    } else {
      CK(cudaMemcpyAsync(D.dX0.p, X0, (size_t)n * 8, cudaMemcpyHostToDevice, D.stream));
    }
    const int blocks = (int)std::min<size_t>(((size_t)nb * n + 255) / 256, (size_t)D.sms * 8);
    expand_children_kernel<<<blocks, 256, 0, D.stream>>>(nb, n, D.dPXL.p, D.dPXU.p, D.dCP.p, D.dCC.p, D.dCW.p, D.dCB.p, D.dXL.p,
                                                         D.dXU.p, pX0, x0_per_parent ? D.dX0.p : nullptr);
    CK(cudaGetLastError());
    E->launches++;
    int rc = launch_ipm(E, D, nb, D.dXL.p, D.dXU.p, D.dX0.p, x0_per_parent, D.dOBJ.p, D.dST.p, D.dIT.p,
                        X ? D.dX.p : nullptr, G ? D.dG.p : nullptr, DUALS ? D.dD.p : nullptr, D.stream);
    if (rc) return rc;
  }
  return collect_batch(E, B, per, OBJ, STATUS, ITERS, X, G, DUALS);
}

extern "C" int bnlp_solve_batch_warm(bnlp_engine *E, int B, const double *XL, const double *XU, const double *X0,
                                     int x0_per_child, const double *DUAL0, int dual0_per_child, double *OBJ, int *STATUS,
                                     int *ITERS, double *X, double *G, double *DUALS) {
  if (!E) { set_err("null engine"); return -1; }
  if (B <= 0) return 0;
  const int n = E->M.n, m = E->M.m, nd = (int)E->devs.size();
  const int nmul = 2 * n + m;
  const int per = (B + nd - 1) / nd;  // contiguous child shards, one per device
  // phase 1: enqueue copies + kernels on every device
  for (int d = 0; d < nd; ++d) {
    const int b0 = d * per, nb = std::min(per, B - b0);
    if (nb <= 0) break;
    Device &D = E->devs[d];
    CK(cudaSetDevice(D.id));
    if (D.dXL.reserve((size_t)nb * n) || D.dXU.reserve((size_t)nb * n) || D.dX0.reserve((size_t)(x0_per_child ? nb : 1) * n) ||
        D.dOBJ.reserve(nb) || D.dST.reserve(nb) || D.dIT.reserve(nb) || D.dX.reserve((size_t)nb * n) ||
        D.dG.reserve((size_t)nb * std::max(m, 1)) || D.dD.reserve((size_t)nb * nmul) ||
        (DUAL0 && D.dD0.reserve((size_t)(dual0_per_child ? nb : 1) * nmul))) {
      set_err("device allocation failed");
      return -1;
    }
    CK(cudaEventRecord(D.ev0, D.stream));
    CK(cudaMemcpyAsync(D.dXL.p, XL + (size_t)b0 * n, (size_t)nb * n * 8, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dXU.p, XU + (size_t)b0 * n, (size_t)nb * n * 8, cudaMemcpyHostToDevice, D.stream));
    CK(cudaMemcpyAsync(D.dX0.p, X0 + (x0_per_child ? (size_t)b0 * n : 0), (size_t)(x0_per_child ? nb : 1) * n * 8,
                       cudaMemcpyHostToDevice, D.stream));
    if (DUAL0)
      CK(cudaMemcpyAsync(D.dD0.p, DUAL0 + (dual0_per_child ? (size_t)b0 * nmul : 0), (size_t)(dual0_per_child ? nb : 1) * nmul * 8,
                         cudaMemcpyHostToDevice, D.stream));
    int rc = launch_ipm(E, D, nb, D.dXL.p, D.dXU.p, D.dX0.p, x0_per_child, D.dOBJ.p, D.dST.p, D.dIT.p,
                        X ? D.dX.p : nullptr, G ? D.dG.p : nullptr, DUALS ? D.dD.p : nullptr, D.stream,
                        DUAL0 ? D.dD0.p : nullptr, dual0_per_child);
    if (rc) return rc;
  }
  return collect_batch(E, B, per, OBJ, STATUS, ITERS, X, G, DUALS);
}

extern "C" int bnlp_solve_batch(bnlp_engine *E, int B, const double *XL, const double *XU, const double *X0,
                                int x0_per_child, double *OBJ, int *STATUS, int *ITERS, double *X, double *G,
                                double *DUALS) {
  return bnlp_solve_batch_warm(E, B, XL, XU, X0, x0_per_child, nullptr, 0, OBJ, STATUS, ITERS, X, G, DUALS);
}

// Page-lock a caller-owned host buffer that is passed to the batch calls again and again (cut buffers, bound arrays):
// copies to and from it then run as DMA at PCIe speed instead of through the driver's staging buffer.
extern "C" int bnlp_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return -1;
  CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return 0;
}
extern "C" int bnlp_host_unregister(void *p) {
  if (!p) return -1;
  CK(cudaHostUnregister(p));
  return 0;
}

extern "C" int bnlp_engine_stats(const bnlp_engine *E, long long *iterations, long long *factorisations, double *device_ms) {
  if (!E) return -1;
  if (iterations) *iterations = E->last_iters;
  if (factorisations) *factorisations = E->last_fact;
  if (device_ms) *device_ms = E->last_ms;
  return 0;
}
extern "C" long long bnlp_engine_launches(const bnlp_engine *E) { return E ? E->launches : 0; }
extern "C" int bnlp_engine_sizes(const bnlp_engine *E, int *n, int *m) {
  if (!E) return -1;
  if (n) *n = E->M.n;
  if (m) *m = E->M.m;
  return 0;
}

// ------------------------------------------------------------------------------------------------ FP64 peak probe
__global__ void fp64_peak_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" double bnlp_measure_fp64_peak(bnlp_engine *E, int dev_index) {
  if (!E || dev_index < 0 || dev_index >= (int)E->devs.size()) return -1.0;
  Device &D = E->devs[dev_index];
  if (cudaSetDevice(D.id) != cudaSuccess) return -1.0;
  const int blocks = D.sms * 4, threads = 512, iters = 1 << 16;
  double *out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, D.stream);
    fp64_peak_kernel<<<blocks, threads, 0, D.stream>>>(out, iters);
    cudaEventRecord(e1, D.stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    E->launches++;
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

#include "bnlp_api_eval.inc"
#include "bnlp_api_comm.inc"
