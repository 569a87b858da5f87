// Batched tape evaluation and outer-approximation kernels.
//
// eval_kernel   : B x {eval_f, eval_grad_f, eval_g, eval_jac_g values, eval_h values}
//                 replaces the per-point callbacks of Bonmin::TMINLP (src/Interfaces/BonTMINLP.hpp:268-315) that the
//                 reference forwards to ASL (src/Interfaces/Ampl/BonAmplTMINLP.cpp:629-661).
// oa_rows_kernel: B x OsiTMINLPInterface::getOuterApproximation (src/Interfaces/BonOsiTMINLPInterface.cpp:1924-2107,
//                 cleanNnz :1881-1919) on the values produced by eval_kernel; oa_scan/oa_compact turn the
//                 fixed-capacity rows into CSR.
// One CTA per point for evaluation, one warp per (point, cut) for the OA assembly.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

#include "../../include/bnlp.h"

// The device headers are written in `namespace BNLP_NS` (default: bnlp) so that a translation unit can compile a second
// copy of them with other compile-time parameters (bnlp_launch_coop256.cu: BNLP_NT = 256 in namespace bnlp256).
#ifndef BNLP_NS
#define BNLP_NS bnlp
#endif
namespace BNLP_NS {

constexpr int EV_NT = 256;
// (BNLP_TV / BNLP_TS: the wide build of the term machine — bnlp_launch_wide.cu compiles these headers a second time in
// namespace bnlpw with BNLP_MAX_WIDE_TERM_VARS / _STACK)
#ifndef BNLP_TV
#define BNLP_TV BNLP_MAX_TERM_VARS
#endif
#ifndef BNLP_TS
#define BNLP_TS BNLP_MAX_TERM_STACK
#endif
constexpr int TV = BNLP_TV;                            // max variables of a term
constexpr int TH = TV * (TV + 1) / 2;                  // packed local Hessian
constexpr int TS = BNLP_TS;

struct EvalModelHostOffsets {
  int n = 0, m = 0, nnz_jac = 0, nnz_hes = 0, n_terms = 0, nqs = 0, mnl = 0, nlin = 0, nfq = 0, P = 0;
  // offsets into the packed int buffer
  int lin_start = 0, lin_col = 0, lin_jpos = 0, qs_fn = 0, qs_i = 0, qs_j = 0, qs_jpos_i = 0, qs_jpos_j = 0, qs_hpos = 0,
      term_fn = 0, term_var_start = 0, term_var = 0, term_op_start = 0, term_op = 0, term_iarg = 0, term_jpos = 0,
      term_hpos = 0, jac_start = 0, jac_col = 0, hes_row = 0, hes_col = 0, nl_rows = 0, fq_fn = 0, fq_dense = 0,
      row_cap_start = 0;
  // fixed-order accumulation of element terms (IPM kernel): terms grouped by function, Hessian entries grouped by target
  int fn_term_start = 0, fn_term_list = 0, hseg_start = 0, hseg_target = 0, hent_term = 0, hent_lidx = 0, n_hseg = 0;
  int hseg_hpos = 0;     // position of every Hessian-entry group in the sparse Hessian structure (eval_kernel)
  // offsets into the packed double buffer
  int c = 0, lin_val = 0, qs_v = 0, term_darg = 0;
};

struct EvalModelDev {
  int n, m, nnz_jac, nnz_hes, n_terms, nqs, mnl, nlin, nfq, P;
  const int *lin_start, *lin_col, *lin_jpos;
  const int *qs_fn, *qs_i, *qs_j, *qs_jpos_i, *qs_jpos_j, *qs_hpos;
  const int *term_fn, *term_var_start, *term_var, *term_op_start, *term_op, *term_iarg, *term_jpos, *term_hpos;
  const int *jac_start, *jac_col, *hes_row, *hes_col, *nl_rows;
  const int *fq_fn;        // functions that own a dense quadratic block (nfq)
  const int *fq_dense;     // 1 if the block came from an explicit dense block (full Jacobian row structure)
  const int *row_cap_start;  // mnl+2 : fixed-capacity offsets of the OA rows of one point
  const int *fn_term_start, *fn_term_list;                       // m+2, n_terms : terms of each function, in tape order
  const int *hseg_start, *hseg_target, *hent_term, *hent_lidx;   // Hessian entries grouped by packed target tri(i)+j
  const int *hseg_hpos;    // n_hseg : position of the group's target in the sparse Hessian structure
  int n_hseg;
  const double *c, *lin_val, *qs_v, *term_darg;
  const double *gl, *gu;
  // dense quadratic blocks shared with the IPM model
  const int *qidx;         // m+1
  const double *Qp, *Qf;
  int ldp, ldq;            // strides of the packed blocks / full-matrix rows
};

inline EvalModelDev make_eval_model_dev(const EvalModelHostOffsets &o, const int *I, const double *D, const double *gl,
                                        const double *gu) {
  EvalModelDev e{};
  e.n = o.n; e.m = o.m; e.nnz_jac = o.nnz_jac; e.nnz_hes = o.nnz_hes; e.n_terms = o.n_terms; e.nqs = o.nqs;
  e.mnl = o.mnl; e.nlin = o.nlin; e.nfq = o.nfq; e.P = o.P;
  e.lin_start = I + o.lin_start; e.lin_col = I + o.lin_col; e.lin_jpos = I + o.lin_jpos;
  e.qs_fn = I + o.qs_fn; e.qs_i = I + o.qs_i; e.qs_j = I + o.qs_j; e.qs_jpos_i = I + o.qs_jpos_i;
  e.qs_jpos_j = I + o.qs_jpos_j; e.qs_hpos = I + o.qs_hpos;
  e.term_fn = I + o.term_fn; e.term_var_start = I + o.term_var_start; e.term_var = I + o.term_var;
  e.term_op_start = I + o.term_op_start; e.term_op = I + o.term_op; e.term_iarg = I + o.term_iarg;
  e.term_jpos = I + o.term_jpos; e.term_hpos = I + o.term_hpos;
  e.jac_start = I + o.jac_start; e.jac_col = I + o.jac_col; e.hes_row = I + o.hes_row; e.hes_col = I + o.hes_col;
  e.nl_rows = I + o.nl_rows; e.fq_fn = I + o.fq_fn; e.fq_dense = I + o.fq_dense; e.row_cap_start = I + o.row_cap_start;
  e.fn_term_start = I + o.fn_term_start; e.fn_term_list = I + o.fn_term_list; e.hseg_start = I + o.hseg_start;
  e.hseg_target = I + o.hseg_target; e.hent_term = I + o.hent_term; e.hent_lidx = I + o.hent_lidx; e.n_hseg = o.n_hseg;
  e.hseg_hpos = I + o.hseg_hpos;
  e.c = D + o.c; e.lin_val = D + o.lin_val; e.qs_v = D + o.qs_v; e.term_darg = D + o.term_darg;
  e.gl = gl; e.gu = gu;
  return e;
}

// ---- second-order forward-mode evaluation of one element term phi(x[vars]) by a stack machine:
// every stack entry carries value, gradient (nv) and packed lower Hessian (nv(nv+1)/2) w.r.t. the local variables.
struct HDual {
  double v;
  double g[TV];
  double h[TH];
};

__device__ inline int hidx(int a, int b) { return a >= b ? (a * (a + 1) / 2 + b) : (b * (b + 1) / 2 + a); }

__device__ inline void hd_unary(HDual &a, int nv, double f, double f1, double f2) {
  // z = f(a):  z' = f1 a',  z'' = f2 a' a'^T + f1 a''
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j <= i; ++j) a.h[i * (i + 1) / 2 + j] = f2 * a.g[i] * a.g[j] + f1 * a.h[i * (i + 1) / 2 + j];
  for (int i = 0; i < nv; ++i) a.g[i] *= f1;
  a.v = f;
}

template <class MM>
__device__ inline void eval_term(const MM &M, int t, const double *x, HDual *st, HDual &out, int order) {
  const int v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
  const int p0 = M.term_op_start[t], p1 = M.term_op_start[t + 1];
  const int nh = order >= 2 ? nv * (nv + 1) / 2 : 0;
  int sp = 0;
  for (int p = p0; p < p1; ++p) {
    const int op = M.term_op[p];
    if (op == BNLP_OP_CONST || op == BNLP_OP_VAR) {
      HDual &z = st[sp++];
      for (int i = 0; i < nv; ++i) z.g[i] = 0.0;
      for (int i = 0; i < nh; ++i) z.h[i] = 0.0;
      if (op == BNLP_OP_CONST) z.v = M.term_darg[p];
      else { const int s = M.term_iarg[p]; z.v = x[M.term_var[v0 + s]]; z.g[s] = 1.0; }
      continue;
    }
    if (op == BNLP_OP_ADD || op == BNLP_OP_SUB || op == BNLP_OP_MUL || op == BNLP_OP_DIV || op == BNLP_OP_POW) {
      HDual &b = st[sp - 1];
      HDual &a = st[sp - 2];
      --sp;
      if (op == BNLP_OP_ADD || op == BNLP_OP_SUB) {
        const double sgn = (op == BNLP_OP_ADD) ? 1.0 : -1.0;
        a.v += sgn * b.v;
        for (int i = 0; i < nv; ++i) a.g[i] += sgn * b.g[i];
        for (int i = 0; i < nh; ++i) a.h[i] += sgn * b.h[i];
      } else if (op == BNLP_OP_MUL) {
        for (int i = 0; i < nv; ++i)
          for (int j = 0; j <= i; ++j) {
            const int e = i * (i + 1) / 2 + j;
            if (e < nh) a.h[e] = a.v * b.h[e] + b.v * a.h[e] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
          }
        for (int i = 0; i < nv; ++i) a.g[i] = a.v * b.g[i] + b.v * a.g[i];
        a.v *= b.v;
      } else if (op == BNLP_OP_DIV) {
        // a / b = a * (1/b)
        const double r = 1.0 / b.v;
        hd_unary(b, nv, r, -r * r, 2.0 * r * r * r);
        for (int i = 0; i < nv; ++i)
          for (int j = 0; j <= i; ++j) {
            const int e = i * (i + 1) / 2 + j;
            if (e < nh) a.h[e] = a.v * b.h[e] + b.v * a.h[e] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
          }
        for (int i = 0; i < nv; ++i) a.g[i] = a.v * b.g[i] + b.v * a.g[i];
        a.v *= b.v;
      } else {  // a ^ b = exp(b log a)
        const double la = log(a.v);
        hd_unary(a, nv, la, 1.0 / a.v, -1.0 / (a.v * a.v));
        for (int i = 0; i < nv; ++i)
          for (int j = 0; j <= i; ++j) {
            const int e = i * (i + 1) / 2 + j;
            if (e < nh) a.h[e] = a.v * b.h[e] + b.v * a.h[e] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
          }
        for (int i = 0; i < nv; ++i) a.g[i] = a.v * b.g[i] + b.v * a.g[i];
        a.v *= b.v;
        const double ex = exp(a.v);
        hd_unary(a, nv, ex, ex, ex);
      }
      continue;
    }
    HDual &a = st[sp - 1];
    const double v = a.v;
    switch (op) {
      case BNLP_OP_NEG: hd_unary(a, nv, -v, -1.0, 0.0); break;
      case BNLP_OP_SQR: hd_unary(a, nv, v * v, 2.0 * v, 2.0); break;
      case BNLP_OP_POWC: {
        const double p_ = M.term_darg[p];
        const double f = pow(v, p_);
        const double f1 = (p_ == 0.0) ? 0.0 : p_ * pow(v, p_ - 1.0);
        const double f2 = (p_ == 0.0 || p_ == 1.0) ? 0.0 : p_ * (p_ - 1.0) * pow(v, p_ - 2.0);
        hd_unary(a, nv, f, f1, f2);
        break;
      }
      case BNLP_OP_EXP: { const double e = exp(v); hd_unary(a, nv, e, e, e); break; }
      case BNLP_OP_LOG: hd_unary(a, nv, log(v), 1.0 / v, -1.0 / (v * v)); break;
      case BNLP_OP_SQRT: { const double s = sqrt(v); hd_unary(a, nv, s, 0.5 / s, -0.25 / (s * v)); break; }
      case BNLP_OP_SIN: hd_unary(a, nv, sin(v), cos(v), -sin(v)); break;
      case BNLP_OP_COS: hd_unary(a, nv, cos(v), -sin(v), -cos(v)); break;
      case BNLP_OP_ABS: hd_unary(a, nv, fabs(v), v >= 0 ? 1.0 : -1.0, 0.0); break;
      case BNLP_OP_ATAN: { const double q = 1.0 / (1.0 + v * v); hd_unary(a, nv, atan(v), q, -2.0 * v * q * q); break; }
      case BNLP_OP_ASIN: { const double q = 1.0 / sqrt(1.0 - v * v); hd_unary(a, nv, asin(v), q, v * q * q * q); break; }
      case BNLP_OP_ACOS: { const double q = 1.0 / sqrt(1.0 - v * v); hd_unary(a, nv, acos(v), -q, -v * q * q * q); break; }
      default: break;
    }
  }
  out = st[0];
}

struct EvalArgs {
  int B;
  const double *X;      // B x n
  const double *LAM;    // B x m or null
  double sigma;
  double *F, *G, *GRADF, *JVAL, *HVAL;   // any may be null
  const double *U;      // B x nfq x n : u = Q x of every dense quadratic block, from quad_batch_kernel (null if nfq == 0)
  double *TH_scratch;   // B x n_terms x TH : local Hessians of the element terms (needed when HVAL is asked for and the model has terms)
};

__device__ inline int jac_pos(const EvalModelDev &M, int fn, int col, int dense) {
  // position of (row fn-1, col) in the row-major Jacobian structure
  const int s = M.jac_start[fn - 1], e = M.jac_start[fn];
  if (dense) return s + col;
  int lo = s, hi = e - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, cc = M.jac_col[mid];
    if (cc == col) return mid;
    if (cc < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

#ifdef BNLP_EVAL_KERNELS   // the __global__ entry points are compiled by bnlp_api.cu only
// ---- u_{b,q} = Q_q x_b for EVERY dense quadratic block and a whole batch of points, as a dense contraction on the FP64
// tensor pipe (mma.sync m8n8k4 f64): C[8 matrix rows x 8 points] += Qf[8 x 4] * X'[4 x 8].  CTA (s, g) stages the PB
// points of group g in shared memory (row stride = 1 mod 16 doubles: conflict-free fragments) and serves slice s of the
// (block, 8-row tile) list, so a block element crosses L2 -> SM once per PB points instead of once per point (the
// per-point version re-read all blocks for every point: 200 MB per point at C4).  Every global request is a whole
// 128-byte line: lane (lr, lc) loads sector lc of {row r0+lr, columns 16g .. 16g+15}; k-step s contracts over the
// columns {16g + 4lc + s}, a permutation of the summation index that the B fragments follow.
constexpr int QB_NT = 512;
__device__ __forceinline__ void dmma_ev(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int PB>
__global__ void __launch_bounds__(QB_NT, 1) quad_batch_kernel(EvalModelDev M, int B, const double *X, double *U, int ldd) {
  extern __shared__ double Dsm[];   // [PB][ldd]
  const int n = M.n, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = QB_NT / 32, NP = PB / 8;
  const int b0 = blockIdx.y * PB;
  for (int e = tid; e < PB * ldd; e += QB_NT) {
    const int b = b0 + e / ldd, j = e % ldd;
    Dsm[e] = (b < B && j < n) ? X[(size_t)b * n + j] : 0.0;
  }
  __syncthreads();
  const int tpb = (n + 7) >> 3, ntile = M.nfq * tpb;
  const int t0 = (int)(((long long)ntile * blockIdx.x) / gridDim.x), t1 = (int)(((long long)ntile * (blockIdx.x + 1)) / gridDim.x);
  const int lr = lane >> 2, lc = lane & 3;
  const int ngrp = (n + 15) >> 4;
  for (int t = t0 + wid; t < t1; t += NW) {
    const int qi = t / tpb, r0 = (t - qi * tpb) << 3;
    const int k = M.fq_fn[qi];
    const int r = r0 + lr;
    const bool rv = r < n;
    const double2 *Qrow = reinterpret_cast<const double2 *>(M.Qf + ((size_t)M.qidx[k] * n + min(r, n - 1)) * M.ldq) + 2 * lc;
    double acc[NP][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p][0] = acc[p][1] = 0.0;
    double2 a0 = __ldg(Qrow), a1 = __ldg(Qrow + 1);
    for (int g = 0; g < ngrp; ++g) {
      const int gn = min(g + 1, ngrp - 1);
      const double2 n0 = __ldg(Qrow + 8 * gn), n1 = __ldg(Qrow + 8 * gn + 1);   // next group in flight under the DMMAs
      const int cb = 16 * g + 4 * lc;
      const double av[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const double a = rv ? av[s4] : 0.0;
#pragma unroll
        for (int p = 0; p < NP; ++p) dmma_ev(acc[p][0], acc[p][1], a, Dsm[(8 * p + lr) * ldd + cb + s4]);
      }
      a0 = n0; a1 = n1;
    }
    // C[row lr][point 8p + 2lc, + 1]: for a fixed (p, half) the eight lanes lr = 0..7 write 64 contiguous bytes
    if (rv) {
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int b = b0 + 8 * p + 2 * lc + hh;
          if (b < B) U[((size_t)b * M.nfq + qi) * n + r] = acc[p][hh];
        }
    }
  }
}

__global__ void __launch_bounds__(EV_NT) eval_kernel(EvalModelDev M, EvalArgs A) {
  extern __shared__ double fv[];   // m+1 function values
  const int b = blockIdx.x, tid = threadIdx.x, n = M.n, m = M.m;
  const int w = tid >> 5, lane = tid & 31, NWARP = EV_NT / 32;
  const double *x = A.X + (size_t)b * n;
  const double *lam = A.LAM ? A.LAM + (size_t)b * m : nullptr;
  double *gf = A.GRADF ? A.GRADF + (size_t)b * n : nullptr;
  double *jv = A.JVAL ? A.JVAL + (size_t)b * M.nnz_jac : nullptr;
  double *hv = A.HVAL ? A.HVAL + (size_t)b * M.nnz_hes : nullptr;
  for (int k = tid; k <= m; k += EV_NT) fv[k] = M.c[k];
  if (gf) for (int j = tid; j < n; j += EV_NT) gf[j] = 0.0;
  if (jv) for (int e = tid; e < M.nnz_jac; e += EV_NT) jv[e] = 0.0;
  __syncthreads();
  // ---- linear part: one warp per function (plain stores: CSR columns are unique per row)
  for (int k = w; k <= m; k += NWARP) {
    double acc = 0.0;
    for (int e = M.lin_start[k] + lane; e < M.lin_start[k + 1]; e += 32) {
      const double a = M.lin_val[e];
      const int j = M.lin_col[e];
      acc = fma(a, x[j], acc);
      if (k == 0) { if (gf) gf[j] = a; }
      else if (jv) jv[M.lin_jpos[e]] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) fv[k] += acc;
  }
  __syncthreads();
  // ---- dense quadratic blocks: u = Q x comes from quad_batch_kernel (one contraction for the whole batch); one warp per
  // function adds 1/2 x.u to the value and u to the gradient / Jacobian row
  for (int qi = w; qi < M.nfq; qi += NWARP) {
    const int k = M.fq_fn[qi], dense = M.fq_dense[qi];
    const double *u = A.U + ((size_t)b * M.nfq + qi) * n;
    double vacc = 0.0;
    for (int r = lane; r < n; r += 32) {
      const double ur = u[r];
      vacc = fma(0.5 * x[r], ur, vacc);
      if (k == 0) { if (gf) gf[r] += ur; }
      else if (jv) {
        const int p = jac_pos(M, k, r, dense);
        if (p >= 0) jv[p] += ur;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vacc += __shfl_xor_sync(0xffffffffu, vacc, o);
    if (lane == 0) fv[k] += vacc;
  }
  __syncthreads();
  // ---- sparse quadratic triplets (kept sparse only when the model was not densified)
  for (int e = tid; e < M.nqs; e += EV_NT) {
    const int k = M.qs_fn[e], i = M.qs_i[e], j = M.qs_j[e];
    const double v = M.qs_v[e];
    if (i == j) {
      atomicAdd(&fv[k], 0.5 * v * x[i] * x[i]);
      if (k == 0) { if (gf) atomicAdd(&gf[i], v * x[i]); }
      else if (jv) atomicAdd(&jv[M.qs_jpos_i[e]], v * x[i]);
    } else {
      atomicAdd(&fv[k], v * x[i] * x[j]);
      if (k == 0) { if (gf) { atomicAdd(&gf[i], v * x[j]); atomicAdd(&gf[j], v * x[i]); } }
      else if (jv) { atomicAdd(&jv[M.qs_jpos_i[e]], v * x[j]); atomicAdd(&jv[M.qs_jpos_j[e]], v * x[i]); }
    }
  }
  // ---- Hessian: sigma Q_0 + sum lambda_i Q_i at the structural positions
  if (hv) {
    for (int e = tid; e < M.nnz_hes; e += EV_NT) {
      double a = 0.0;
      if (M.nfq > 0) {
        const size_t pe = (size_t)M.hes_row[e] * (M.hes_row[e] + 1) / 2 + M.hes_col[e];
        for (int qi = 0; qi < M.nfq; ++qi) {
          const int k = M.fq_fn[qi];
          const double wgt = (k == 0) ? A.sigma : (lam ? lam[k - 1] : 0.0);
          if (wgt != 0.0) a = fma(wgt, M.Qp[(size_t)M.qidx[k] * M.ldp + pe], a);
        }
      }
      hv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < M.nqs; e += EV_NT) {
      const int k = M.qs_fn[e];
      const double wgt = (k == 0) ? A.sigma : (lam ? lam[k - 1] : 0.0);
      if (wgt != 0.0) atomicAdd(&hv[M.qs_hpos[e]], wgt * M.qs_v[e]);
    }
  }
  __syncthreads();
  // ---- general nonlinear element terms: one thread per FUNCTION walks its terms in tape order, so values, gradients and
  // Jacobian entries are accumulated in a fixed order; the Hessian values go through per-position lists (second pass)
  if (M.n_terms > 0) {
    HDual st[TS];
    HDual r;
    for (int k = tid; k <= m; k += EV_NT) {
      const double wgt = (k == 0) ? A.sigma : (lam ? lam[k - 1] : 0.0);
      double fk = 0.0;
      for (int e = M.fn_term_start[k]; e < M.fn_term_start[k + 1]; ++e) {
        const int t = M.fn_term_list[e];
        eval_term(M, t, x, st, r, hv ? 2 : 1);
        const int v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
        fk += r.v;
        for (int a = 0; a < nv; ++a) {
          if (k == 0) { if (gf) gf[M.term_var[v0 + a]] += r.g[a]; }
          else if (jv) jv[M.term_jpos[v0 + a]] += r.g[a];
        }
        if (hv) {      // the local Hessian is kept; the entries of different functions that meet in one position are summed below
          double *th = A.TH_scratch + ((size_t)b * M.n_terms + t) * TH;
          for (int a = 0; a < nv * (nv + 1) / 2; ++a) th[a] = r.h[a];
        }
      }
      fv[k] += fk;
    }
    if (hv) {
      // one thread per touched position of the Hessian, contributions in the fixed order of the host-built lists: the
      // values are bit-reproducible (round 1 used FP64 atomics here)
      __syncthreads();
      for (int sg = tid; sg < M.n_hseg; sg += EV_NT) {
        double acc = 0.0;
        for (int e = M.hseg_start[sg]; e < M.hseg_start[sg + 1]; ++e) {
          const int t = M.hent_term[e], k = M.term_fn[t];
          const double wgt = (k == 0) ? A.sigma : (lam ? lam[k - 1] : 0.0);
          if (wgt != 0.0) acc = fma(wgt, A.TH_scratch[((size_t)b * M.n_terms + t) * TH + M.hent_lidx[e]], acc);
        }
        hv[M.hseg_hpos[sg]] += acc;
      }
    }
  }
  __syncthreads();
  if (A.F && tid == 0) A.F[b] = fv[0];
  if (A.G) for (int i = tid; i < m; i += EV_NT) A.G[(size_t)b * m + i] = fv[i + 1];
}

// ------------------------------------------------------------------------------------------------ outer approximation
struct OaArgs {
  int B, getObj, use_duals, has_x2, skip_non_violated, obj_nonlinear;
  const double *X, *DUALS, *XL, *XU, *X2;   // X2: B x (n+1) when getObj cut is filtered, else B x n is enough
  int x2_stride;
  double theta;
  bnlp_oa_tols tol;
  const double *F, *G, *GRADF, *JVAL;       // from eval_kernel
  // fixed-capacity temporaries
  int *tcol; double *tval; int *cnt;        // B x cap, B x cap, B x (mnl+1)
  double *LB, *UB; int *src_row;            // B x (mnl+1)
  int cap;                                  // capacity per point = row_cap_start[mnl+1]
};

// cleanNnz (BonOsiTMINLPInterface.cpp:1881-1919); returns keep flag, accumulates the lb/ub corrections
__device__ inline bool clean_nnz(double value, double colLower, double colUpper, double rowLower, double rowUpper,
                                 double colsol, double &lbadd, double &ubadd, double tiny, double veryTiny, double infty) {
  if (fabs(value) >= tiny) return true;
  if (fabs(value) < veryTiny) return false;
  const bool colUpBounded = colUpper < infty, colLoBounded = colLower > -infty;
  const bool rowNotLoBounded = rowLower <= -infty, rowNotUpBounded = rowUpper >= infty;
  const bool pos = value > 0;
  if (colUpBounded && pos && rowNotUpBounded) { lbadd += value * (colsol - colUpper); return false; }
  if (colUpBounded && !pos && rowNotLoBounded) { ubadd += value * (colsol - colUpper); return false; }
  if (colLoBounded && !pos && rowNotUpBounded) { lbadd += value * (colsol - colLower); return false; }
  if (colLoBounded && pos && rowNotLoBounded) { ubadd += value * (colsol - colLower); return false; }
  return true;
}

// one warp per (point, slot); slot < mnl: nonlinear row, slot == mnl: objective cut
__global__ void oa_rows_kernel(EvalModelDev M, OaArgs A) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int slots = M.mnl + 1;
  if (gw >= A.B * slots) return;
  const int b = gw / slots, slot = gw % slots;
  const int n = M.n, m = M.m;
  const double DMAX = DBL_MAX, infty_ = A.tol.infty;
  const double *x = A.X + (size_t)b * n;
  int *tcol = A.tcol + (size_t)b * A.cap + M.row_cap_start[slot];
  double *tval = A.tval + (size_t)b * A.cap + M.row_cap_start[slot];
  const size_t o = (size_t)b * slots + slot;
  double lb, ub, lbadd = 0.0, ubadd = 0.0;
  int kept = 0;
  if (slot < M.mnl) {
    const int i = M.nl_rows[slot];
    const double g = A.G[(size_t)b * m + i], rl = M.gl[i], ru = M.gu[i];
    lb = (rl > -infty_) ? rl - g : -DMAX;
    ub = (ru < infty_) ? ru - g : DMAX;
    if (A.use_duals && rl > -DMAX && ru < DMAX) {   // :1968-1974 (two different infinities)
      const double du = A.DUALS[(size_t)b * (2 * n + m) + 2 * n + i];
      if (du >= 0) lb = -DMAX;
      if (du <= 0) ub = DMAX;
    }
    const int s = M.jac_start[i], e = M.jac_start[i + 1];
    const double *jv = A.JVAL + (size_t)b * M.nnz_jac;
    for (int p0 = s; p0 < e; p0 += 32) {
      const int p = p0 + lane;
      bool keep = false;
      double v = 0.0;
      int col = 0;
      if (p < e) {
        v = jv[p]; col = M.jac_col[p];
        keep = clean_nnz(v, A.XL[col], A.XU[col], rl, ru, x[col], lbadd, ubadd, A.tol.tiny, A.tol.very_tiny, infty_);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = kept + __popc(mask & ((1u << lane) - 1));
        tcol[pos] = col; tval[pos] = v;
      }
      kept += __popc(mask);
    }
    __syncwarp();
    // kept coefficients move v*x to both sides (:1995-1998); separate accumulator, applied only to finite sides
    double kx = 0.0;
    for (int p = lane; p < kept; p += 32) kx = fma(tval[p], x[tcol[p]], kx);
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) {
      kx += __shfl_xor_sync(0xffffffffu, kx, q);
      lbadd += __shfl_xor_sync(0xffffffffu, lbadd, q);
      ubadd += __shfl_xor_sync(0xffffffffu, ubadd, q);
    }
    lb += lbadd; ub += ubadd;
    if (lb > -DMAX) lb += kx;
    if (ub < DMAX) ub += kx;
    bool gen = true;
    if (A.has_x2 && A.skip_non_violated) {   // reference: only when oaHandler_->logLevel() > 0
      const double *x2 = A.X2 + (size_t)b * A.x2_stride;
      double rhs = 0.0;
      for (int p = lane; p < kept; p += 32) rhs = fma(tval[p], x2[tcol[p]], rhs);
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) rhs += __shfl_xor_sync(0xffffffffu, rhs, q);
      const double viol = fmax(fmax(0.0, rhs - ub), lb - rhs);
      if (viol < A.theta) gen = false;
    }
    // rhs relaxation: the lb branch of the reference is dead code
    if (ub < DMAX) ub += A.tol.rhs_relax * fmax(fabs(ub), 1.0);
    if (lane == 0) {
      A.LB[o] = lb; A.UB[o] = ub;
      A.src_row[o] = gen ? i : -1;
      A.cnt[o] = gen ? kept : 0;
    }
  } else {
    const bool want = A.getObj == 2 || (A.getObj && A.obj_nonlinear);
    if (!want) {
      if (lane == 0) { A.LB[o] = 0.0; A.UB[o] = 0.0; A.src_row[o] = -1; A.cnt[o] = 0; }
      return;
    }
    const double f = A.F[b];
    const double *gf = A.GRADF + (size_t)b * n;
    lb = -f; ub = -f;
    for (int p0 = 0; p0 < n; p0 += 32) {
      const int p = p0 + lane;
      bool keep = false;
      double v = 0.0;
      if (p < n) {
        v = gf[p];
        keep = clean_nnz(v, A.XL[p], A.XU[p], -DMAX, 0.0, x[p], lbadd, ubadd, A.tol.tiny, 1e-15, infty_);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = kept + __popc(mask & ((1u << lane) - 1));
        tcol[pos] = p; tval[pos] = v;
      }
      kept += __popc(mask);
    }
    __syncwarp();
    double kx = 0.0;
    for (int p = lane; p < kept; p += 32) kx = fma(tval[p], x[tcol[p]], kx);
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) {
      kx += __shfl_xor_sync(0xffffffffu, kx, q);
      lbadd += __shfl_xor_sync(0xffffffffu, lbadd, q);
      ubadd += __shfl_xor_sync(0xffffffffu, ubadd, q);
    }
    ub += ubadd + kx;
    if (lane == 0) { tcol[kept] = n; tval[kept] = -1.0; }   // the alpha column (:2076)
    __syncwarp();
    kept += 1;
    bool gen = true;
    if (A.has_x2) {
      const double *x2 = A.X2 + (size_t)b * A.x2_stride;
      double rhs = 0.0;
      for (int p = lane; p < kept; p += 32) {
        const int cidx = tcol[p];
        rhs = fma(tval[p], (cidx < A.x2_stride) ? x2[cidx] : 0.0, rhs);
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) rhs += __shfl_xor_sync(0xffffffffu, rhs, q);
      if (fmax(0.0, rhs - ub) < A.theta) gen = false;
    }
    if (lane == 0) {
      A.LB[o] = -DMAX; A.UB[o] = ub;
      A.src_row[o] = gen ? m : -1;
      A.cnt[o] = gen ? kept : 0;
    }
  }
}

// ---- violation probes: OsiTMINLPInterface::getConstraintsViolation / getNonLinearitiesViolation
// (BonOsiTMINLPInterface.cpp:1839-1876): max over NONLINEAR rows of the bound violation (bounds beyond +-1e10 ignored);
// with OBJ != null the result is max(norm, f(x) - obj).  One warp per point.
__global__ void violation_kernel(EvalModelDev M, int B, const double *G, const double *F, const double *OBJ, double *VIOL) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B) return;
  double norm = 0.0;
  for (int s = lane; s < M.mnl; s += 32) {
    const int i = M.nl_rows[s];
    const double g = G[(size_t)gw * M.m + i];
    double rv = 0.0;
    if (M.gl[i] > -1e10) rv = fmax(0.0, M.gl[i] - g);
    if (M.gu[i] < 1e10) rv = fmax(rv, g - M.gu[i]);
    norm = fmax(norm, rv);
  }
#pragma unroll
  for (int q = 16; q > 0; q >>= 1) norm = fmax(norm, __shfl_xor_sync(0xffffffffu, norm, q));
  if (lane == 0) {
    if (OBJ) { const double d = F[gw] - OBJ[gw]; if (d > norm) norm = d; }
    VIOL[gw] = norm;
  }
}

// ---- Benders cut: OsiTMINLPInterface::getBendersCut (BonOsiTMINLPInterface.cpp:2111-2202): the lambda-weighted
// aggregate of the linearisations of the nonlinear rows with |lambda| > 1e-6 (+ the objective cut when the objective is
// nonlinear and the point is a proven optimum).  One CTA per point; the dense coefficient vector lives in `cut` (B x (n+1)).
struct BendersArgs {
  int B, with_obj;
  const double *X, *DUALS, *XL, *XU, *F, *G, *GRADF, *JVAL;
  bnlp_oa_tols tol;
  double *cut;          // B x (n+1), zeroed by the kernel
  int *col_idx; double *val; int *row_len; double *UB;   // outputs: fixed capacity n+1 per point
};
__global__ void benders_kernel(EvalModelDev M, BendersArgs A) {
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, n = M.n, m = M.m;
  __shared__ double s_ub[2];   // [0] ub accumulated, [1] scratch
  __shared__ int s_cnt;
  const double *x = A.X + (size_t)b * n;
  const double *lam = A.DUALS + (size_t)b * (2 * n + m) + 2 * n;
  double *cut = A.cut + (size_t)b * (n + 1);
  for (int j = tid; j <= n; j += nt) cut[j] = 0.0;
  if (tid == 0) { s_ub[0] = 0.0; s_cnt = 0; }
  __syncthreads();
  double ub_p = 0.0, lb_dummy = 0.0;
  for (int s = tid; s < M.mnl; s += nt) {
    const int i = M.nl_rows[s];
    const double l = lam[i];
    if (fabs(l) > 1e-6) ub_p += l * ((l < 0 ? M.gl[i] : M.gu[i]) - A.G[(size_t)b * m + i]);
  }
  for (int s = 0; s < M.mnl; ++s) {
    const int i = M.nl_rows[s];
    const double l = lam[i];
    if (!(fabs(l) > 1e-6)) continue;
    for (int p = M.jac_start[i] + tid; p < M.jac_start[i + 1]; p += nt) {
      const int col = M.jac_col[p];
      const double coeff = l * A.JVAL[(size_t)b * M.nnz_jac + p];
      if (clean_nnz(coeff, A.XL[col], A.XU[col], M.gl[i], M.gu[i], x[col], lb_dummy, ub_p, A.tol.tiny, A.tol.very_tiny, A.tol.infty)) {
        atomicAdd(&cut[col], coeff);
        ub_p += coeff * x[col];
      }
    }
  }
  atomicAdd(&s_ub[0], ub_p);
  __syncthreads();
  if (A.with_obj) {
    __syncthreads();
    if (tid == 0) s_ub[0] = -A.F[b];   // the reference overwrites ub here (:2173)
    __syncthreads();
    double up = 0.0;
    for (int j = tid; j < n; j += nt) {
      const double c = A.GRADF[(size_t)b * n + j];
      if (clean_nnz(c, A.XL[j], A.XU[j], -DBL_MAX, 0.0, x[j], lb_dummy, up, A.tol.tiny, 1e-15, A.tol.infty)) {
        cut[j] += c;   // column j is owned by this thread in this phase
        up += c * x[j];
      }
    }
    atomicAdd(&s_ub[0], up);
    __syncthreads();
  }
  // sparse vector: (n, -1) first when the objective cut is included, then the columns with |coef| > 1e-20 in order
  if (tid == 0) {
    int *ci = A.col_idx + (size_t)b * (n + 1);
    double *cv = A.val + (size_t)b * (n + 1);
    int c = 0;
    if (A.with_obj) { ci[c] = n; cv[c] = -1.0; ++c; }
    for (int j = 0; j < n; ++j)
      if (fabs(cut[j]) > 1e-20) { ci[c] = j; cv[c] = cut[j]; ++c; }
    A.row_len[b] = c;
    A.UB[b] = s_ub[0];
  }
}

// exclusive scan of the row counts into CSR row_start (single CTA; rows = B*(mnl+1))
__global__ void oa_scan_kernel(const int *cnt, int *row_start, int rows) {
  __shared__ int part[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < rows; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < rows) ? cnt[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < rows) row_start[i] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_start[rows] = carry;
}

// copy the kept entries of every row from the fixed-capacity temporaries into the CSR arrays (one warp per row)
__global__ void oa_compact_kernel(EvalModelDev M, OaArgs A, const int *row_start, int *col_idx, double *val) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int slots = M.mnl + 1;
  if (gw >= A.B * slots) return;
  const int b = gw / slots, slot = gw % slots;
  const int cntv = A.cnt[gw];
  const int *tcol = A.tcol + (size_t)b * A.cap + M.row_cap_start[slot];
  const double *tval = A.tval + (size_t)b * A.cap + M.row_cap_start[slot];
  const int dst = row_start[gw];
  for (int p = lane; p < cntv; p += 32) { col_idx[dst + p] = tcol[p]; val[dst + p] = tval[p]; }
}

#endif  // BNLP_EVAL_KERNELS

}  // namespace BNLP_NS
