// Batched primal-dual interior-point kernel: ONE CTA PER CHILD RELAXATION, persistent CTAs pulling
// children from an atomic work queue (variable iteration counts never hold an SM hostage).
//
// Replaces, for a whole batch, what Bonmin reaches one child at a time through
//   TNLPSolver::OptimizeTNLP            /root/reference/src/Interfaces/BonTNLPSolver.hpp:127
//   IpoptSolver::OptimizeTNLP           /root/reference/src/Interfaces/Ipopt/BonIpoptSolver.cpp:134-176
// under Bonmin's Ipopt option set (BonIpoptSolver.cpp:265-297) and fills the TMINLP2TNLP result
// fields (BonTMINLP2TNLP.cpp:474-509).  Algorithm: DESIGN.md §4 (condensed form
// H = W + Sigma_x + J' Sigma_s J, probing mu oracle, filter line search, exact phase-1 for infeasibility).
//
// Data layout per CTA
//   shared memory : packed lower triangle of the augmented condensed matrix [H; rhs0'; rhs1'] ((N+2)(N+3)/2 f64,
//                   N = n+1 incl. the phase-1 elastic), all iterate vectors, a 17 x Npad staging area
//   global (L2)   : raw gradients grad[(m+1) x n] and directional products uq[(m+1) x n] of this child
//   shared by all : the model (c, A dense, packed quadratic blocks Qp) — read-only, L2 resident
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "bnlp_eval.cuh"

namespace BNLP_NS {

#ifndef BNLP_NT
#define BNLP_NT 512
#endif
constexpr int IPM_NT = BNLP_NT;       // threads per CTA (512; the cooperative kernel is also built with 256 for two CTAs per SM)
constexpr int IPM_NW = IPM_NT / 32;   // warps per CTA
constexpr int IPM_KC = 16;            // rank of one staged SYRK / trailing update
constexpr int IPM_NB = 8;             // Cholesky panel width
constexpr int IPM_MAXE = 8;           // equality rows supported by the Schur-complement path
constexpr int TREC = 2 + TV + TH;     // per-term record: value, local gradient, packed local Hessian, trial value
constexpr int IPM_MAXC = 16;          // column accumulators per lane in the symmetric mat-vec: n <= 512

// TNLPSolver::ReturnStatus (BonTNLPSolver.hpp:29-44)
enum { ST_iterationLimit = -3, ST_computationError = -2, ST_solvedOptimal = 1, ST_solvedOptimalTol = 2,
       ST_provenInfeasible = 3 };

struct IpmOpt {
  double tol = 1e-8, acceptable_tol = 1e-6;
  int max_iter = 3000, acceptable_iter = 15;
  double dual_inf_tol = 1.0, constr_viol_tol = 1e-4, compl_inf_tol = 1e-4;
  double bound_relax = 1e-8, bound_push = 0.01, bound_frac = 0.01;
  double warm_push = 1e-3;   // warm_start_{bound,slack_bound,mult_bound}_push / _frac of Ipopt (warm_start_init_point yes)
  double tau_min = 0.99, sigma_min = 1e-6, sigma_max = 100, mu_min = 1e-11, mu_max_fact = 1e3, mu_max = 1e5;
  double s_max = 100, kappa_sigma = 1e10;
  double gamma_theta = 1e-4, gamma_phi = 1e-8, eta_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, delta = 1.0;
  double alpha_min_frac = 0.05, alpha_red = 0.5;
  double theta_max_fact = 1e4, theta_min_fact = 1e-4;
  double scal_max_grad = 100.0;
  double infeas_ytol = 1e8, infeas_ctol = 1e-3, infeas_tol = 1e-7;
  int max_phase1 = 8;
  double mu_safeguard = 1e-4;
};

struct DevModel {
  int n, m, N, R, nq, P, npad;
  int ldp, ldq;                 // strides: packed block (multiple of 16 doubles), full-matrix row (multiple of 16)
  const double *c, *A;          // (m+1), (m+1) x n
  const int *qidx;              // m+1 : packed block index or -1
  const double *Qp;             // nq x P packed lower triangles (Hessian accumulation)
  const double *Qf;             // nq x n x n full symmetric copies (mat-vecs)
  const int *row_fn;            // R one-sided rows  sgn * F_fn(x) <= bnd
  const double *row_sgn, *row_bnd;
  const int *fn_row_up, *fn_row_lo;  // m+1 : row index of the upper / lower side of a function, or -1
  int E;                        // equality rows (g_L == g_U): handled through a Schur complement, E <= IPM_MAXE
  const int *eq_fn;             // E : function index
  const double *eq_bnd;         // E : right-hand side
  const int *fn_eq;             // m+1 : equality index of a function, or -1
  // general nonlinear element terms (postfix programs, include/bnlp.h); evaluated by eval_term (bnlp_eval.cuh)
  int n_terms;
  const int *term_fn, *term_var_start, *term_var, *term_op_start, *term_op, *term_iarg;
  const double *term_darg;
  // fixed-order accumulation of the terms (no floating-point atomics: results do not depend on thread timing)
  const int *fn_term_start, *fn_term_list, *hseg_start, *hseg_target, *hent_term, *hent_lidx;
  int n_hseg;
  const double *gl, *gu;        // m
  // J' Sigma J of SPARSE linear rows (a linear function with <= IPM_SPARSE_ROW nonzeros, e.g. the linking rows w_j <= z_j of
  // C5): instead of a dense rank-1 update per row, the few entries of H they touch are listed per TARGET entry (i, k), so
  // one thread sums a target's contributions in a fixed order (no atomics).  Entry e of target s: row sp_ent_row[e] & 0xfffffff,
  // kind sp_ent_row[e] >> 28 (0: p = a_i a_k, both columns < n; 1: p = a_k, i = the phase-1 elastic; 2: i = k = the elastic).
  // Only the rows of dense_rows[0 .. Rd) go through the staged DMMA updates.
  int Rd, sp_nseg;
  const int *dense_rows, *sp_seg_start, *sp_seg_i, *sp_seg_k, *sp_ent_row;
  const double *sp_ent_p;
  // the same rows as the cooperative kernel reads them (built in build_ipm_view, explained there): sp_rows[0 .. Rs) = the
  // sparse rows; column lists, slot-major: slot s of column j at [s * spc_ld + j] (row or -1, coefficient); touched entries
  // of H as 16-byte records {packed index | row << 32 | last << 62, a_i a_k} dealt to 512 virtual threads: [s * 512 + v]
  int Rs, spc_slots, spc_ld, sph_slots;
  const int *sp_rows, *spc_row;
  const double *spc_val, *sph_rec;
  // LOW-RANK quadratic parts (large-model path only): Q_k = sum_r v_r v_r' over the factor rows fn_lr_start[k] ..
  // fn_lr_start[k+1] of lrV (n_lr x n), found by a pivoted Cholesky when the model is loaded (a sum of squares of linear forms
  // arrives as a dense n x n block from the tape builder).  u_k = Q_k d costs 2 r n FMAs instead of n^2 and no block is
  // streamed (C4: 100 blocks of 500 x 500 = 200 MB per pass and CTA); their part of the Hessian, sum_k w_k Q_k, is a
  // rank-(sum r) update done with the staged DMMA rows of J' Sigma J.  n_lr = 0: every block is dense.
  int n_lr;
  const int *fn_lr_start;       // m+2
  const double *lrV;
};
constexpr int IPM_SPARSE_ROW = 8;
constexpr int IPM_LR_MAX = 16;     // a quadratic block of rank <= 16 is kept as factor rows in the large-model path

struct IpmArgs {
  int B;
  const double *XL, *XU, *X0;   // device: B x n, B x n, (x0_stride ? B x n : n)
  int x0_stride;
  // warm start (IpoptSolver::ReOptimizeTNLP after setWarmStart, BonIpoptSolver.cpp:369-419): multipliers [z_L | z_U | lambda]
  // the solve starts from, (dual0_stride ? B x (2n+m) : 2n+m), or null = cold start (unit multipliers)
  const double *DUAL0;
  int dual0_stride;
  double *OBJ; int *STATUS; int *ITERS;
  double *X, *G, *DUALS;        // B x n, B x m, B x (2n+m); any may be null
  int *queue;                   // work-queue counter (zeroed by the host)
  double *ws; size_t ws_stride; // per-CTA global workspace (doubles)
  int h_in_smem;
  int coop_tma;                 // cooperative kernel: stage the Jacobian rows with bulk copies (TMA); measured slower, off by default
  int tm_kc;                    // tile-major path: staged rows per rank-K update of H (16 or 32: the host picks what shared memory allows)
  long long *stats;             // [0] total iterations, [1] factorisations (optional)
  long long *phase_cycles;      // 10 counters or null
  IpmOpt o;
};

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// FP64 tensor-core MMA (SASS DMMA.8x8x4): C[8x8] += A[8x4] B[4x8]; lane (lr = lane / 4, lc = lane % 4) holds A[lr][lc],
// B[lc][lr], C[lr][2 lc], C[lr][2 lc + 1]
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// streaming read of shared read-only model data (no L1 allocation: one CTA per SM never re-reads it from L1)
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// same load, but pinned where it is written (volatile): used for explicit software prefetch
__device__ __forceinline__ double ld_stream_v(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// ---- block reductions: v[0..NS) summed, v[NS..NS+NM) maxed; result broadcast to every thread.
// Stage 1: warp shuffles, lane 0 of every warp leaves its K partials in red[w*K + k].  Stage 2: thread 16k + q picks up
// partial q of value k and the 16-lane groups finish with four more shuffles (all K values in parallel; the old version
// had every thread walk all 16 x K partials).  Two barriers.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce(double *v, double *red) {
  constexpr int K = NS + NM;
  static_assert(K * IPM_NW <= IPM_NT && (IPM_NW == 16 || IPM_NW == 8), "stage 2 uses one thread per (value, warp) pair");
  __shared__ double red_out[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double t = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = (k < NS) ? v[k] + t : fmax(v[k], t);
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) red[w * K + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < ((K * IPM_NW + 31) & ~31)) {      // whole warps take part in the shuffles
    const int k = threadIdx.x / IPM_NW, q = threadIdx.x % IPM_NW;
    double a = (k < K) ? red[q * K + k] : 0.0;
#pragma unroll
    for (int o = IPM_NW / 2; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, a, o);
      a = (k < NS) ? a + t : fmax(a, t);
    }
    if (q == 0 && k < K) red_out[k] = a;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = red_out[k];
}

// ---- H[i,k] += sign * sum_c M[c][i] * M[c][k]   for lo <= k <= i < hi  (packed lower triangle, 4x4 register tiles)
// M is staged in shared memory as K rows of length ld (entries outside [lo,hi) must be zero-padded to a multiple of 4).
__device__ __forceinline__ void syrk_tiles(double *H, int lo, int hi, const double *M, int ld, int K, double sign) {
  const int b0 = lo >> 2, b1 = (hi + 3) >> 2;     // tile block range
  const int nb = b1 - b0;
  const int ntiles = (nb * (nb + 1)) >> 1;
  for (int t = threadIdx.x; t < ntiles; t += IPM_NT) {
    // linear lower-triangular index -> (bi, bk)
    int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (tri(bi + 1) <= t) ++bi;
    while (tri(bi) > t) --bi;
    const int bk = t - tri(bi);
    const int i0 = (b0 + bi) << 2, k0 = (b0 + bk) << 2;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int c = 0; c < K; ++c) {
      const double *Mc = M + c * ld;
      const double2 a01 = *reinterpret_cast<const double2 *>(Mc + i0);
      const double2 a23 = *reinterpret_cast<const double2 *>(Mc + i0 + 2);
      const double2 b01 = *reinterpret_cast<const double2 *>(Mc + k0);
      const double2 b23 = *reinterpret_cast<const double2 *>(Mc + k0 + 2);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y};
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + a;
      if (i < lo || i >= hi) continue;
      double *Hi = H + tri(i);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int k = k0 + b;
        if (k >= lo && k <= i) Hi[k] += sign * acc[a][b];
      }
    }
  }
}

// ---- blocked right-looking Cholesky of rows/cols [0,N) of the packed matrix; rows [N, Nrows) are carried along
// (they hold right-hand sides, so the forward substitution L w = rhs happens inside the factorisation).
// Per 8-column panel: (1) one thread factors the 8x8 diagonal block in registers, (2) every other row of the panel is
// solved against it by its own thread (TRSM) and staged, (3) the trailing matrix gets a rank-8 update from register
// tiles.  The DIAGONAL of L is stored as its RECIPROCAL.  Returns false (block-uniform) on a non-positive pivot.
__device__ __forceinline__ bool chol_packed(double *H, int N, int Nrows, double *stage, int ld, int *fail) {
  for (int j0 = 0; j0 < N; j0 += IPM_NB) {
    const int nb = min(IPM_NB, N - j0);
    __syncthreads();
    if (threadIdx.x == 0) {
      double a[IPM_NB][IPM_NB];
#pragma unroll
      for (int i = 0; i < IPM_NB; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) a[i][k] = (i < nb) ? H[tri(j0 + i) + j0 + k] : (i == k ? 1.0 : 0.0);
      bool ok = true;
#pragma unroll
      for (int j = 0; j < IPM_NB; ++j) {
        double d = a[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-a[j][k], a[j][k], d);
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double inv = rsqrt(d);
        a[j][j] = inv;   // reciprocal of the diagonal entry
#pragma unroll
        for (int i = j + 1; i < IPM_NB; ++i) {
          double v = a[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) v = fma(-a[i][k], a[j][k], v);
          a[i][j] = v * inv;
        }
      }
#pragma unroll
      for (int i = 0; i < IPM_NB; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k)
          if (i < nb) H[tri(j0 + i) + j0 + k] = a[i][k];
      *fail = ok ? 0 : 1;
    }
    __syncthreads();
    if (*fail) return false;
    const int lo = j0 + nb;
    if (lo >= Nrows) break;
    const int lo4 = lo & ~3, hi4 = (Nrows + 3) & ~3;
    // TRSM: row i of the panel, x L_D' = a  (L_D strictly lower from shared memory, diagonal as reciprocal)
    for (int i = lo4 + threadIdx.x; i < hi4; i += IPM_NT) {
      double xr[IPM_NB];
      if (i >= lo && i < Nrows) {
        double *Hi = H + tri(i) + j0;
#pragma unroll
        for (int c = 0; c < IPM_NB; ++c) xr[c] = (c < nb) ? Hi[c] : 0.0;
#pragma unroll
        for (int c = 0; c < IPM_NB; ++c) {
          if (c < nb) {
            const double *Lc = H + tri(j0 + c) + j0;
            double v = xr[c];
#pragma unroll
            for (int k = 0; k < c; ++k) v = fma(-xr[k], Lc[k], v);
            xr[c] = v * Lc[c];
          }
        }
#pragma unroll
        for (int c = 0; c < IPM_NB; ++c)
          if (c < nb) Hi[c] = xr[c];
      } else {
#pragma unroll
        for (int c = 0; c < IPM_NB; ++c) xr[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < IPM_NB; ++c)
        if (c < nb) stage[c * ld + i] = xr[c];
    }
    __syncthreads();
    syrk_tiles(H, lo, Nrows, stage, ld, nb, -1.0);
  }
  __syncthreads();
  return true;
}

// ---- backward substitution L' x = w for the two right-hand sides stored in rows N and N+1 (warps 0 and 1),
// blocked by the same 8-row panels (diagonal of L holds reciprocals)
__device__ __forceinline__ void backsolve2(double *H, int N) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < 2) {
    double *b = H + tri(N + w);
    const int nblk = (N + IPM_NB - 1) / IPM_NB;
    for (int blk = nblk - 1; blk >= 0; --blk) {
      const int i0 = blk * IPM_NB, nb = min(IPM_NB, N - i0);
      // solve the diagonal block (every lane redundantly; all reads are broadcasts)
      double xr[IPM_NB];
#pragma unroll
      for (int c = 0; c < IPM_NB; ++c) xr[c] = (c < nb) ? b[i0 + c] : 0.0;
#pragma unroll
      for (int c = IPM_NB - 1; c >= 0; --c) {
        if (c < nb) {
          double v = xr[c];
#pragma unroll
          for (int cc = c + 1; cc < IPM_NB; ++cc)
            if (cc < nb) v = fma(-H[tri(i0 + cc) + i0 + c], xr[cc], v);
          xr[c] = v * H[tri(i0 + c) + i0 + c];
        }
      }
      __syncwarp();
      if (lane < nb) b[i0 + lane] = (lane == 0) ? xr[0] : (lane == 1) ? xr[1] : (lane == 2) ? xr[2] : (lane == 3) ? xr[3]
                                  : (lane == 4) ? xr[4] : (lane == 5) ? xr[5] : (lane == 6) ? xr[6] : xr[7];
      // b[k] -= sum_c L[i0+c][k] x_c   for k < i0
      for (int k = lane; k < i0; k += 32) {
        double v = b[k];
#pragma unroll
        for (int c = 0; c < IPM_NB; ++c)
          if (c < nb) v = fma(-H[tri(i0 + c) + k], xr[c], v);
        b[k] = v;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

// ---- u_k = Q_k d for every quadratic block, from the FULL (n x n, row-major) copies: one warp per row, four rows
// in flight per warp (aligned, fully coalesced 8-byte loads; lane l owns columns l + 32c and keeps d there in
// registers).  uq[k][0..n) goes to global; du[k] = d.u_k is reduced afterwards by one warp per block.
template <int MAXC>
__device__ __forceinline__ void pass_U(const DevModel &M, const double *d, double *uq, double *du, const int *qfn,
                                       int nqf) {
  const int n = M.n, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nc = (n + 31) >> 5;
  double dreg[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) dreg[c] = (c < nc && lane + 32 * c < n) ? d[lane + 32 * c] : 0.0;
  const int nrows = nqf * n;
  for (int g = w * 4; g < nrows; g += IPM_NW * 4) {
    const double *Qr[4];
    double *out[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rho = min(g + u, nrows - 1);
      const int qi = rho / n, r = rho - qi * n;
      const int k = qfn[qi];
      Qr[u] = M.Qf + ((size_t)M.qidx[k] * n + r) * M.ldq;
      out[u] = uq + (size_t)k * n + r;
    }
    double rs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c4 = 0; c4 < MAXC; c4 += 4) {
      if (c4 < nc) {
        double qv[4][4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int j = min(lane + 32 * (c4 + cc), n - 1);
#pragma unroll
          for (int u = 0; u < 4; ++u) qv[cc][u] = ld_stream(Qr[u] + j);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int u = 0; u < 4; ++u) rs[u] = fma(qv[cc][u], dreg[c4 + cc], rs[u]);   // dreg is 0 beyond n
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rs[u] += __shfl_xor_sync(0xffffffffu, rs[u], o);
    }
    if (lane < 4 && g + lane < nrows) *((lane == 0) ? out[0] : (lane == 1) ? out[1] : (lane == 2) ? out[2] : out[3]) =
        (lane == 0) ? rs[0] : (lane == 1) ? rs[1] : (lane == 2) ? rs[2] : rs[3];
  }
  __syncthreads();
  for (int qi = w; qi < nqf; qi += IPM_NW) {
    const int k = qfn[qi];
    const double *uk = uq + (size_t)k * n;
    double part = 0.0;
    for (int j = lane; j < n; j += 32) part = fma(d[j], uk[j], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) du[k] = part;
  }
  __syncthreads();
}

// A function without a quadratic part and without element terms is LINEAR: its gradient is the constant row of M.A, which
// every CTA shares.  Reading it there instead of from a per-CTA copy keeps the copies of the linear rows out of L2 (C5: 201
// of 204 functions are linear; the per-CTA gradient block is 650 KB, and 148 of them plus the H blocks do not fit L2).
__device__ __forceinline__ bool fn_is_linear(const DevModel &M, int k) {
  return M.qidx[k] < 0 && (M.n_terms == 0 || M.fn_term_start[k + 1] == M.fn_term_start[k]);
}
__device__ __forceinline__ bool fn_is_lowrank(const DevModel &M, int k, bool large_path) {
  return large_path && M.n_lr > 0 && M.fn_lr_start[k + 1] > M.fn_lr_start[k];
}
// u_k = Q_k d and du[k] = d'Q_k d for the low-rank blocks: s_r = v_r . d (one warp per factor row, kept in lrs), then
// u_k[j] = sum_r s_r v_r[j]
__device__ __forceinline__ void pass_U_lowrank(const DevModel &M, const double *d, double *uq, double *du, double *lrs) {
  const int n = M.n, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = w; r < M.n_lr; r += IPM_NW) {
    const double *vr = M.lrV + (size_t)r * n;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(vr[j], d[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lrs[r] = acc;
  }
  __syncthreads();
  for (int k = w; k <= M.m; k += IPM_NW) {
    const int r0 = M.fn_lr_start[k], r1 = M.fn_lr_start[k + 1];
    if (r1 <= r0) continue;
    double dd = 0.0;
    for (int r = r0; r < r1; ++r) dd = fma(lrs[r], lrs[r], dd);
    for (int j = lane; j < n; j += 32) {
      double u = 0.0;
      for (int r = r0; r < r1; ++r) u = fma(lrs[r], M.lrV[(size_t)r * n + j], u);
      uq[(size_t)k * n + j] = u;
    }
    if (lane == 0) du[k] = dd;
  }
  __syncthreads();
}
__device__ __forceinline__ const double *grow(const DevModel &M, const double *grad, int k) {
  return fn_is_linear(M, k) ? M.A + (size_t)k * M.n : grad + (size_t)k * M.n;
}

// out[k] = base_k . d  for k = 0..m  (base = raw gradients, one warp per function)
__device__ __forceinline__ void dots_grad(const DevModel &M, const double *grad, const double *d0, const double *d1,
                                          double *g0, double *g1) {
  const int n = M.n, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = w; k <= M.m; k += IPM_NW) {
    const double *g = grow(M, grad, k);
    double a = 0.0, b = 0.0;
    for (int j = lane; j < n; j += 32) {
      const double gv = g[j];
      a = fma(gv, d0[j], a);
      b = fma(gv, d1[j], b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { g0[k] = a; g1[k] = b; }
  }
  __syncthreads();
}

enum { F_FIXED = 1, F_HASL = 2, F_HASU = 4 };

}  // namespace BNLP_NS
#include "bnlp_dense_tm.cuh"
namespace BNLP_NS {

// packed lower-triangular index e = tri(i) + k  ->  (i, k)
__device__ __forceinline__ void packed_ik(int e, int &i, int &k) {
  i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
  while (tri(i + 1) <= e) ++i;
  while (tri(i) > e) --i;
  k = e - tri(i);
}

// optional per-phase cycle counters (IpmArgs::phase_cycles != nullptr): 0 residuals 1 W pass 2 J'SJ 3 Cholesky
// 4 back-substitution 5 direction/mu 6 U pass 7 line search 8 update 9 init
#define PH_MARK(idx)                                               \
  do {                                                             \
    if (A.phase_cycles) {                                          \
      const long long t_ = clock64();                              \
      ph[idx] += t_ - ph_t0;                                       \
      ph_t0 = t_;                                                  \
    }                                                              \
  } while (0)

template <int MAXC>
__global__ void __launch_bounds__(IPM_NT, 512 / IPM_NT) ipm_kernel(DevModel M, IpmArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int n = M.n, m = M.m, N = M.N, R = M.R, npad = M.npad;
  const int tid = threadIdx.x;
  const IpmOpt &o = A.o;
  // ---- carve shared memory
  double *sp = smem;
  const int E = M.E;
  const int Haug = tri(N + 2 + E);
  double *H;
  double *ws = A.ws + (size_t)blockIdx.x * A.ws_stride;
  double *grad = ws;                              // (m+1) x n
  double *uq = grad + (size_t)(m + 1) * n;        // (m+1) x n
  const int NTm = M.n_terms;
  double *tgl = uq + (size_t)(m + 1) * n;          // per term: value, local gradient (TV), packed local Hessian (TH)
  // Large models (H does not fit shared memory): H lives in this CTA's global workspace (L2 / L1 resident) in the TILE-MAJOR
  // layout of bnlp_dense_tm.cuh and is assembled, factorised and solved with the DMMA routines there (left-looking
  // Cholesky: every tile is read-modify-written once; the packed scalar path re-read the whole trailing matrix from L2 per
  // 8-column panel and ran at ~2 % of the FP64 peak).  The two right-hand sides then live in their own shared-memory vectors.
  const bool TM = !A.h_in_smem;
  const TmLayout TL = tm_layout(N + 2 + E);
  const int nb8 = (N + 7) & ~7, lds_tm = nb8 + 4;
  if (A.h_in_smem) { H = sp; sp += (Haug + 1) & ~1; } else { H = tgl + (((size_t)NTm * TREC + 1) & ~(size_t)1); }   // 16-byte aligned tiles
  const int KT = TM ? A.tm_kc : IPM_KC;       // rows per staged update (tile-major path: 32 halves the read-modify-writes of H)
  double *stage = sp; sp += max((IPM_KC + 1) * npad, TM ? KT * lds_tm : 0);
  double *x = sp; sp += npad;
  double *zL = sp; sp += npad;
  double *zU = sp; sp += npad;
  double *Lr = sp; sp += npad;
  double *Ur = sp; sp += npad;
  double *dx = sp; sp += npad;
  double *Sx = sp; sp += npad;
  const int Rp = (R + 1) & ~1, Kp = (m + 2) & ~1;
  double *s = sp; sp += Rp;
  double *v = sp; sp += Rp;
  double *rc = sp; sp += Rp;
  double *ds = sp; sp += Rp;
  double *dv = sp; sp += Rp;
  double *Ss = sp; sp += Rp;
  double *bb = sp; sp += Rp;
  double *fsc = sp; sp += Rp;       // sc_r * sgn_r
  double *jd0 = sp; sp += Rp;       // J d0, J d1 per row
  double *jd1 = sp; sp += Rp;
  double *val = sp; sp += Kp;
  double *gd0 = sp; sp += Kp;
  double *gd1 = sp; sp += Kp;
  double *du = sp; sp += Kp;
  double *wq = sp; sp += 2 * Kp;
  double *tval = sp; sp += Kp;       // element-term values per function at x / at a trial point
  double *tvt = sp; sp += Kp;
  double *red = sp; sp += IPM_NW * 12;
  double *ye = sp; sp += IPM_MAXE;       // equality multipliers, residuals, scales, step parts, Schur complement
  double *ce = sp; sp += IPM_MAXE;
  double *esc = sp; sp += IPM_MAXE;
  double *de0 = sp; sp += IPM_MAXE;
  double *de1 = sp; sp += IPM_MAXE;
  double *Ssc = sp; sp += IPM_MAXE * IPM_MAXE;
  int *flg = reinterpret_cast<int *>(sp);  // N ints
  int *fixlist = flg + ((N + 1) & ~1);     // N ints
  int *qblk = fixlist + ((N + 1) & ~1);    // m+1 ints
  int *qfn = qblk + Kp;                    // m+1 ints: functions that own a quadratic block
  double *qwt = wq + Kp;                   // m+1 weights of the compact list
  __shared__ int s_child, s_nqa, s_nfix, s_fail;
  __shared__ int s_ctr[2];
  // the two right-hand sides / solutions: the augmented rows N, N+1 of the packed H, or (tile-major H) their own vectors
  double *d0 = H + tri(N), *d1 = H + tri(N + 1);
  double *lrs = nullptr, *sgn16 = nullptr;      // low-rank path: s_r = v_r . d per factor row; signs of a staged chunk
  if (TM) {
    double *dv2 = reinterpret_cast<double *>(qfn + Kp + (Kp & 1));
    d0 = dv2; d1 = dv2 + nb8;
    sgn16 = d1 + nb8;
    lrs = sgn16 + 32;
  }
  __shared__ int s_nqf;
  if (tid == 0) {
    int c = 0;
    for (int k = 0; k <= m; ++k)
      if (M.qidx[k] >= 0 && !fn_is_lowrank(M, k, TM)) qfn[c++] = k;
    s_nqf = c;
  }
#pragma unroll 1
  for (int k = tid; k <= m; k += IPM_NT) du[k] = 0.0;
  __syncthreads();
  const int nqf = s_nqf;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_child = atomicAdd(A.queue, 1);
    __syncthreads();
    const int child = s_child;
    if (child >= A.B) break;
    long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t0 = A.phase_cycles ? clock64() : 0;
    const double *xl = A.XL + (size_t)child * n, *xu = A.XU + (size_t)child * n;
    const double *x0 = A.X0 + (size_t)child * A.x0_stride;

    // ------------------------------------------------------------------ bounds, fixed variables, start point
    double acc2[2] = {0.0, 0.0};  // [0] sum inconsistent, [1] max width
    double badw = 0.0;
#pragma unroll 1
    for (int j = tid; j < N; j += IPM_NT) {
      int f = 0;
      double l, u, xx;
      if (j < n) {
        l = xl[j]; u = xu[j];
        if (l > -kInfD) f |= F_HASL;
        if (u < kInfD) f |= F_HASU;
        if (u < l) { acc2[0] += 1.0; badw = fmax(badw, l - u); }
        if (u <= l) f |= F_FIXED;
        acc2[1] = fmax(acc2[1], u - l);
        xx = x0[j];
        double lr = l, ur = u;
        if (f & F_FIXED) { xx = l; }
        else {
          if (f & F_HASL) lr = l - o.bound_relax * fmax(1.0, fabs(l));
          if (f & F_HASU) ur = u + o.bound_relax * fmax(1.0, fabs(u));
          const double bp = A.DUAL0 ? o.warm_push : o.bound_push, bf = A.DUAL0 ? o.warm_push : o.bound_frac;   // warm start: 1e-3
          if ((f & F_HASL) && (f & F_HASU)) {
            const double pl = fmin(bp * fmax(1.0, fabs(lr)), bf * (ur - lr));
            const double pu = fmin(bp * fmax(1.0, fabs(ur)), bf * (ur - lr));
            xx = fmin(fmax(xx, lr + pl), ur - pu);
          } else if (f & F_HASL) xx = fmax(xx, lr + bp * fmax(1.0, fabs(lr)));
          else if (f & F_HASU) xx = fmin(xx, ur - bp * fmax(1.0, fabs(ur)));
        }
        Lr[j] = lr; Ur[j] = ur;
      } else {  // elastic t: fixed at 0 outside phase 1
        f = F_HASL | F_FIXED; xx = 0.0; Lr[j] = -o.bound_relax; Ur[j] = kInfD;
      }
      x[j] = xx; flg[j] = f; zL[j] = 0.0; zU[j] = 0.0;
    }
    {
      double r2[3] = {acc2[0], acc2[1], badw};
      block_reduce<1, 2>(r2, red);
      acc2[0] = r2[0]; acc2[1] = r2[1]; badw = r2[2];
    }
    const bool inconsistent = acc2[0] > 0.0;
    const bool all_fixed = !inconsistent && acc2[1] <= 1e-5;
    if (inconsistent) {  // evaluate at the start point clipped into the (crossed) box
#pragma unroll 1
      for (int j = tid; j < n; j += IPM_NT) {
        const double lo = fmin(xl[j], xu[j]), hi = fmax(xl[j], xu[j]);
        x[j] = fmin(fmax(x0[j], lo), hi);
      }
    } else if (all_fixed) {
#pragma unroll 1
      for (int j = tid; j < n; j += IPM_NT) x[j] = xl[j];
    }
    __syncthreads();

    // ------------------------------------------------------------------ full evaluation at x: grad_k = a_k + Q_k x
    pass_U<MAXC>(M, x, uq, du, qfn, nqf);
    if (TM && M.n_lr > 0) pass_U_lowrank(M, x, uq, du, lrs);
    {
      const int w = tid >> 5, lane = tid & 31;
      for (int k = w; k <= m; k += IPM_NW) {
        const double *a = M.A + (size_t)k * n;
        double *g = grad + (size_t)k * n;
        const bool hq = M.qidx[k] >= 0;
        const bool lin = fn_is_linear(M, k);      // gradient = the shared row of M.A: no per-CTA copy is written
        const double *u = uq + (size_t)k * n;
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) {
          const double aj = a[j], uj = hq ? u[j] : 0.0;
          if (!lin) g[j] = aj + uj;
          acc = fma(x[j], aj + 0.5 * uj, acc);
        }
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
        if (lane == 0) val[k] = M.c[k] + acc;
      }
    }
    __syncthreads();

    // ---- element terms: evaluate at x, add values/gradients to val/grad (sgn = +1) or remove the stored ones (-1)
    // Accumulation order is fixed: thread k walks the terms of function k in tape order (M.fn_term_list).
    auto terms_remove = [&]() {
      if (NTm == 0) return;
#pragma unroll 1
      for (int k = tid; k <= m; k += IPM_NT) {
        for (int e = M.fn_term_start[k]; e < M.fn_term_start[k + 1]; ++e) {
          const int t = M.fn_term_list[e], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
          const double *rec = tgl + (size_t)t * TREC;
          for (int a = 0; a < nv; ++a) grad[(size_t)k * n + M.term_var[v0 + a]] -= rec[1 + a];
        }
        val[k] -= tval[k];
      }
      __syncthreads();
    };
    auto terms_add = [&]() {
      if (NTm == 0) return;
      HDual st[TS];
      HDual r;
      for (int t = tid; t < NTm; t += IPM_NT) {
        eval_term(M, t, x, st, r, 2);
        const int v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
        double *rec = tgl + (size_t)t * TREC;
        rec[0] = r.v;
        for (int a = 0; a < nv; ++a) rec[1 + a] = r.g[a];
        for (int a = 0; a < nv * (nv + 1) / 2; ++a) rec[1 + TV + a] = r.h[a];
      }
      __syncthreads();
#pragma unroll 1
      for (int k = tid; k <= m; k += IPM_NT) {
        double tv = 0.0;
        for (int e = M.fn_term_start[k]; e < M.fn_term_start[k + 1]; ++e) {
          const int t = M.fn_term_list[e], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
          const double *rec = tgl + (size_t)t * TREC;
          tv += rec[0];
          for (int a = 0; a < nv; ++a) grad[(size_t)k * n + M.term_var[v0 + a]] += rec[1 + a];
        }
        tval[k] = tv;
        val[k] += tv;
      }
      __syncthreads();
    };
    terms_add();

    int status = 0;
    int iters = 0;
    double objv = 0.0;
    bool with_duals = false;
    double sf = 1.0;
    bool done = false;

    if (inconsistent || all_fixed) {
      // all-variables-fixed short cut (BonTNLPSolver.cpp:78-161) / trivially infeasible crossed bounds
      double t[1] = {0.0};
      double infe = 0.0;
#pragma unroll 1
      for (int i = tid; i < m; i += IPM_NT) {
        const double g = val[i + 1];
        if (M.gu[i] < kInfD) t[0] = fmax(t[0], g - M.gu[i]);
        if (M.gl[i] > -kInfD) t[0] = fmax(t[0], M.gl[i] - g);
        if (g - M.gl[i] < -1e-7 || g - M.gu[i] > 1e-7) infe = 1.0;
      }
      double r2[2] = {t[0], infe};
      block_reduce<0, 2>(r2, red);
      if (inconsistent) { status = ST_provenInfeasible; objv = fmax(badw, r2[0]); }
      else { status = (r2[1] > 0.0) ? ST_provenInfeasible : ST_solvedOptimal; objv = val[0]; }
      done = true;
    }

    bool phase1 = false;
    int phase1_entries = 0, iters_main = 0, acceptable_cnt = 0;
    double theta_init = -1.0, mu_max = -1.0, delta_w_last = 0.0;
    double th_hist[4] = {0, 0, 0, 0}, a_last0 = 1.0, a_last1 = 1.0;
    long long nfact = 0;

    if (!done) {
      // ---------------------------------------------------------------- gradient-based scaling, rows, slacks
      {
        const int w = tid >> 5, lane = tid & 31;
        for (int k = w; k <= m; k += IPM_NW) {
          const double *g = grow(M, grad, k);
          double gm = 0.0;
          for (int j = lane; j < n; j += 32)
            if (!(flg[j] & F_FIXED)) gm = fmax(gm, fabs(g[j]));
#pragma unroll
          for (int q = 16; q > 0; q >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, q));
          if (lane == 0) wq[k] = (gm > o.scal_max_grad) ? o.scal_max_grad / gm : 1.0;
        }
      }
      __syncthreads();
      sf = wq[0];
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        const int fn = M.row_fn[r];
        const double sc = wq[fn], sg = M.row_sgn[r];
        const double b = sc * M.row_bnd[r];
        fsc[r] = sc * sg;
        bb[r] = b + o.bound_relax * fmax(1.0, fabs(b));
      }
#pragma unroll 1
      for (int e = tid; e < E; e += IPM_NT) esc[e] = wq[M.eq_fn[e]];
      __syncthreads();
    }

    // helpers (all block-uniform)
    // warm (first initialisation only): the caller's multipliers [z_L | z_U | lambda] (unscaled) start the solve, each
    // pushed to at least warm_push, slacks pushed by warm_push; a phase switch restarts from unit multipliers
    auto init_slacks_mults = [&](bool warm) {
      const double *d0w = (warm && A.DUAL0) ? A.DUAL0 + (size_t)child * A.dual0_stride : nullptr;
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        const int fn = M.row_fn[r];
        const double h = fsc[r] * val[fn] - (phase1 ? x[n] : 0.0);
        s[r] = fmax(bb[r] - h, (d0w ? o.warm_push : o.bound_push) * fmax(1.0, fabs(bb[r])));
        v[r] = d0w ? fmax(o.warm_push, fmax(0.0, (fsc[r] > 0 ? 1.0 : -1.0) * d0w[2 * n + fn - 1]) * sf / fabs(fsc[r])) : 1.0;
      }
#pragma unroll 1
      for (int j = tid; j < N; j += IPM_NT) {
        const int f = flg[j];
        const bool w = d0w && j < n;
        zL[j] = (!(f & F_FIXED) && (f & F_HASL)) ? (w ? fmax(o.warm_push, d0w[j] * sf) : 1.0) : 0.0;
        zU[j] = (!(f & F_FIXED) && (f & F_HASU)) ? (w ? fmax(o.warm_push, d0w[n + j] * sf) : 1.0) : 0.0;
      }
#pragma unroll 1
      for (int e = tid; e < E; e += IPM_NT) ye[e] = d0w ? d0w[2 * n + M.eq_fn[e] - 1] * sf / esc[e] : 0.0;
      __syncthreads();
    };
    if (!done) init_slacks_mults(true);
    PH_MARK(9);

    while (!done) {
      const double of = phase1 ? 0.0 : sf;
      const double tcur = phase1 ? x[n] : 0.0;
      // ---------------------------------------------------------------- residuals and error measures
      // sweep 1 over rows: rc, Ss, sums
      double a9[9];  // sums: th, sum_v, compl_sum(rows); max: compl_inf, cviol, cviol_unsc, vmax, viol_max ; sum viol_l1
      double th_p = 0, sumv_p = 0, cs_p = 0, vl1_p = 0, ci_p = 0, cv_p = 0, cvu_p = 0, vmax_p = 0, viol_p = -1e300;
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        const double hraw = fsc[r] * val[M.row_fn[r]];
        const double res = hraw - tcur + s[r] - bb[r];
        rc[r] = res;
        Ss[r] = v[r] / s[r];
        th_p += fabs(res);
        sumv_p += v[r];
        cs_p += s[r] * v[r];
        vl1_p += fmax(0.0, hraw - bb[r]);
        ci_p = fmax(ci_p, s[r] * v[r]);
        cv_p = fmax(cv_p, fabs(res));
        cvu_p = fmax(cvu_p, fabs(res) / fabs(fsc[r]));
        vmax_p = fmax(vmax_p, v[r]);
        viol_p = fmax(viol_p, hraw - bb[r]);
      }
      double sumye_p = 0.0;
#pragma unroll 1
      for (int e = tid; e < E; e += IPM_NT) {
        const double res = esc[e] * (val[M.eq_fn[e]] - M.eq_bnd[e]);
        ce[e] = res;
        th_p += fabs(res);
        cv_p = fmax(cv_p, fabs(res));
        cvu_p = fmax(cvu_p, fabs(res) / esc[e]);
        sumye_p += fabs(ye[e]);
      }
      sumv_p += sumye_p;   // enters s_d only (sum of all constraint multipliers)
      __syncthreads();
      // sweep 2 over variables: gradient of the Lagrangian (J' v), rhs vectors, Sigma_x
      double di_p = 0, sumz_p = 0, nz_p = 0;
#pragma unroll 1
      for (int j = tid; j < N; j += IPM_NT) {
        const int f = flg[j];
        double r0 = 0.0, r1 = 0.0, sx = 0.0;
        if (!(f & F_FIXED)) {
          double g0 = (j < n) ? of * grow(M, grad, 0)[j] : (phase1 ? 1.0 : 0.0);
          if (j < n)
            for (int e = 0; e < E; ++e) g0 = fma(esc[e] * grow(M, grad, M.eq_fn[e])[j], ye[e], g0);
          double jv = 0.0, a0 = 0.0, a1 = 0.0;
          if (j < n) {
            for (int r = 0; r < R; ++r) {
              const double Jrj = fsc[r] * grow(M, grad, M.row_fn[r])[j];
              jv = fma(Jrj, v[r], jv);
              a0 = fma(Jrj, Ss[r] * rc[r], a0);
              a1 += Jrj / s[r];
            }
          } else {
            for (int r = 0; r < R; ++r) { jv -= v[r]; a0 -= Ss[r] * rc[r]; a1 -= 1.0 / s[r]; }
          }
          const double gl = g0 + jv;
          di_p = fmax(di_p, fabs(gl - zL[j] + zU[j]));
          r0 = -g0 - a0;
          r1 = -a1;
          if (f & F_HASL) {
            const double sl = x[j] - Lr[j];
            ci_p = fmax(ci_p, sl * zL[j]); cs_p += sl * zL[j]; sumz_p += zL[j]; nz_p += 1.0;
            sx += zL[j] / sl; r1 += 1.0 / sl;
          }
          if (f & F_HASU) {
            const double su = Ur[j] - x[j];
            ci_p = fmax(ci_p, su * zU[j]); cs_p += su * zU[j]; sumz_p += zU[j]; nz_p += 1.0;
            sx += zU[j] / su; r1 -= 1.0 / su;
          }
        }
        Sx[j] = sx;
        d0[j] = r0;   // right-hand sides go straight into the augmented rows
        d1[j] = r1;
      }
      a9[0] = th_p; a9[1] = sumv_p - sumye_p; a9[2] = cs_p; a9[3] = vl1_p; a9[4] = sumz_p; a9[5] = nz_p;
      double m6[6] = {ci_p, cv_p, cvu_p, vmax_p, viol_p, di_p};
      double sum_ye = 0.0;
      if (E > 0) {
        double t1[1] = {sumye_p};
        block_reduce<1, 0>(t1, red);
        sum_ye = t1[0];
      }
      {
        double all[12] = {a9[0], a9[1], a9[2], a9[3], a9[4], a9[5], m6[0], m6[1], m6[2], m6[3], m6[4], m6[5]};
        block_reduce<6, 6>(all, red);
        a9[0] = all[0]; a9[1] = all[1]; a9[2] = all[2]; a9[3] = all[3]; a9[4] = all[4]; a9[5] = all[5];
        m6[0] = all[6]; m6[1] = all[7]; m6[2] = all[8]; m6[3] = all[9]; m6[4] = all[10]; m6[5] = all[11];
      }
      const double th = a9[0], sum_v = a9[1], vl1 = a9[3];
      const double nz = a9[5] + (double)R, sum_z = a9[4] + sum_v;
      const double compl_inf = m6[0], cviol = m6[1], cviol_unsc = m6[2], ymax = m6[3], viol_max = m6[4], dual_inf = m6[5];
      const double avg_compl = nz > 0 ? a9[2] / nz : 0.0;
      if (theta_init < 0) theta_init = th;
      const double s_d = fmax(o.s_max, (sum_v + sum_ye + sum_z) / fmax(1.0, (double)R + (double)E + nz)) / o.s_max;
      const double s_c = fmax(o.s_max, sum_z / fmax(1.0, nz)) / o.s_max;
      const double E0 = fmax(fmax(dual_inf / s_d, cviol), compl_inf / s_c);
      if (mu_max < 0) mu_max = fmin(o.mu_max, o.mu_max_fact * fmax(avg_compl, 1e-300));

      // ---------------------------------------------------------------- termination tests
      const double osc = phase1 ? 1.0 : sf;
      const bool conv = E0 <= o.tol && dual_inf / osc <= o.dual_inf_tol && cviol_unsc <= o.constr_viol_tol &&
                        compl_inf / osc <= o.compl_inf_tol;
      const bool acc = E0 <= o.acceptable_tol && cviol_unsc <= 1e-2 && compl_inf / osc <= 1e-2;
      acceptable_cnt = acc ? acceptable_cnt + 1 : 0;
      bool to_main = false, to_p1 = false;
      if (phase1) {
        if (vl1 <= 0.0) to_main = true;
        else if (conv || acceptable_cnt >= o.acceptable_iter) {
          // (an only "acceptable" phase-1 optimum proves infeasibility only beyond what its own error E0 could explain)
          if (viol_max > (conv ? o.infeas_tol : fmax(o.infeas_tol, 10.0 * E0))) { status = ST_provenInfeasible; done = true; break; }
          to_main = true;
        }
      } else {
        if (conv) { status = ST_solvedOptimal; objv = val[0]; with_duals = true; done = true; break; }
        if (acceptable_cnt >= o.acceptable_iter) { status = ST_solvedOptimalTol; objv = val[0]; with_duals = true; done = true; break; }
      }
      if (!to_main && iters >= o.max_iter) { status = ST_iterationLimit; objv = val[0]; with_duals = true; done = true; break; }
      if (!to_main && !phase1) {
        th_hist[iters_main & 3] = vl1;
        const double vold = th_hist[(iters_main + 1) & 3];
        bool stall = iters_main >= 3 && vold > 0.0 && vl1 > 0.9 * vold && a_last0 < 0.3 && a_last1 < 0.3;
        iters_main++;
        if (phase1_entries > 0) stall = false;
        if (th > o.infeas_ctol && (ymax > o.infeas_ytol || stall) && vl1 > 0.0) {
          if (phase1_entries >= o.max_phase1) { status = ST_iterationLimit; objv = val[0]; with_duals = true; done = true; break; }
          to_p1 = true;
        }
      }
      if (to_main) {
        __syncthreads();
        phase1 = false; iters_main = 0;
        if (tid == 0) { flg[n] |= F_FIXED; x[n] = 0.0; }
        __syncthreads();
        init_slacks_mults(false);
        theta_init = -1; mu_max = -1; acceptable_cnt = 0;
        ++iters;
        continue;
      }
      auto enter_phase1 = [&]() {
        __syncthreads();
        phase1 = true; phase1_entries++;
        const double vm = fmax(0.0, viol_max), T = fmax(1.0, vm);
        if (tid == 0) { flg[n] &= ~F_FIXED; Lr[n] = -T; x[n] = 1.1 * vm + 0.1 * T; }
        __syncthreads();
        init_slacks_mults(false);
        theta_init = -1; mu_max = -1; acceptable_cnt = 0;
      };
      if (to_p1) { enter_phase1(); ++iters; continue; }

      PH_MARK(0);
      // ---------------------------------------------------------------- condensed matrix + factorisation
#pragma unroll 1
      for (int k = tid; k <= m; k += IPM_NT) {
        double w = (k == 0) ? of : 0.0;
        const int ru = M.fn_row_up[k], rl = M.fn_row_lo[k];
        if (ru >= 0) w += v[ru] * fsc[ru];
        if (rl >= 0) w += v[rl] * fsc[rl];
        const int ei = M.fn_eq[k];
        if (ei >= 0) w += ye[ei] * esc[ei];
        wq[k] = w;
      }
      __syncthreads();
      if (tid == 0) {  // compact list of quadratic blocks with a non-zero weight
        int c = 0;
        for (int k = 0; k <= m; ++k)
          if (M.qidx[k] >= 0 && wq[k] != 0.0 && !fn_is_lowrank(M, k, TM)) { qblk[c] = M.qidx[k]; qwt[c] = wq[k]; ++c; }
        s_nqa = c;
        int f = 0;
        for (int j = 0; j < N; ++j)
          if (flg[j] & F_FIXED) fixlist[f++] = j;
        s_nfix = f;
      }
      __syncthreads();
      const int nqa = s_nqa, nfix = s_nfix;
      double delta_w = 0.0;
      bool fact_ok = false;
      for (int attempt = 0; attempt < 45; ++attempt) {
        // W pass: H[i,k] = sum_q wq[q] Qp_q[i,k] (+ diagonal), one warp per row, fixed rows/cols -> identity
        {
          const int P = M.P;
          for (int base = 0; base < P; base += 4 * IPM_NT) {
            int idx[4];
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int u = 0; u < 4; ++u) idx[u] = min(base + u * IPM_NT + tid, P - 1);
            int qi = 0;
            for (; qi + 4 <= nqa; qi += 4) {
              double qv[4][4];
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const double *Qb = M.Qp + (size_t)qblk[qi + a] * M.ldp;
#pragma unroll
                for (int u = 0; u < 4; ++u) qv[a][u] = ld_stream(Qb + idx[u]);
              }
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const double wgt = qwt[qi + a];
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = fma(wgt, qv[a][u], acc[u]);
              }
            }
            for (; qi < nqa; ++qi) {
              const double *Qb = M.Qp + (size_t)qblk[qi] * M.ldp;
              const double wgt = qwt[qi];
#pragma unroll
              for (int u = 0; u < 4; ++u) acc[u] = fma(wgt, ld_stream(Qb + idx[u]), acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = base + u * IPM_NT + tid;
              if (e < P) {
                if (TM) { int i_, k_; packed_ik(e, i_, k_); H[tm_idx(TL, i_, k_)] = acc[u]; }
                else H[e] = acc[u];
              }
            }
          }
          // elastic row (tile-major: also the corner of the carried rows, so that nothing uninitialised is ever loaded)
          if (TM) {
#pragma unroll 1
            for (int k = tid; k <= n; k += IPM_NT) H[tm_idx(TL, n, k)] = 0.0;
            const int nx = 2 + E;
            for (int q = tid; q < nx * nx; q += IPM_NT)
              if (q % nx <= q / nx) H[tm_idx(TL, N + q / nx, N + q % nx)] = 0.0;
          } else {
#pragma unroll 1
            for (int k = tid; k <= n; k += IPM_NT) H[tri(n) + k] = 0.0;
          }
        }
        __syncthreads();
        if (NTm > 0) {   // + sum_t w_fn(t) * Hessian of term t (stored at the last evaluation); one thread per target entry
          for (int sg = tid; sg < M.n_hseg; sg += IPM_NT) {
            double acc = 0.0;
            for (int e = M.hseg_start[sg]; e < M.hseg_start[sg + 1]; ++e) {
              const int t = M.hent_term[e];
              acc = fma(wq[M.term_fn[t]], tgl[(size_t)t * TREC + 1 + TV + M.hent_lidx[e]], acc);
            }
            if (TM) { int i_, k_; packed_ik(M.hseg_target[sg], i_, k_); H[tm_idx(TL, i_, k_)] += acc; }
            else H[M.hseg_target[sg]] += acc;
          }
          __syncthreads();
        }
        // fixed variables: identity rows/columns; then the diagonal Sigma_x + delta_w
        for (int e = tid; e < nfix * N; e += IPM_NT) {
          const int j = fixlist[e / N], t = e % N;
          if (t < j) H[TM ? tm_idx(TL, j, t) : tri(j) + t] = 0.0;
          else if (t > j) H[TM ? tm_idx(TL, t, j) : tri(t) + j] = 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          double *hjj = H + (TM ? tm_idx(TL, j, j) : tri(j) + j);
          if (flg[j] & F_FIXED) *hjj = 1.0;
          else *hjj += Sx[j] + delta_w;
        }
        __syncthreads();
        PH_MARK(1);
        // J' Sigma_s J through staged rank-KC updates: stage[c][j] = sqrt(Ss_r) J[r][j]
        if (TM) {
          // rank-KT DMMA updates of the tile-major H from staged rows (row stride lds_tm = 4 mod 8: conflict-free fragments)
          for (int r0 = 0; r0 < M.Rd; r0 += KT) {
            const int kc = min(KT, M.Rd - r0);
            for (int c = tid >> 5; c < KT; c += IPM_NW) {
              const int r = M.dense_rows[min(r0 + c, M.Rd - 1)];
              const double sq = (c < kc) ? sqrt(Ss[r]) : 0.0, f = sq * fsc[r];
              const double *g = grow(M, grad, M.row_fn[r]);
              for (int j = tid & 31; j < lds_tm; j += 32) {
                double vv = 0.0;
                if (c < kc && j < N && !(flg[j] & F_FIXED)) vv = (j < n) ? f * g[j] : -sq;
                stage[c * lds_tm + j] = vv;
              }
            }
            if (tid == 0) s_ctr[0] = 0;
            __syncthreads();
            tm_syrk_stage<false>(H, TL, (N + 7) >> 3, stage, lds_tm, KT, s_ctr);
            __syncthreads();
          }
          // low-rank quadratic parts: sum_k w_k sum_r v_r v_r' as staged rows sqrt|w_k| v_r with the sign of w_k
          for (int r0 = 0; r0 < M.n_lr; r0 += KT) {
            const int kc = min(KT, M.n_lr - r0);
            for (int c = tid >> 5; c < KT; c += IPM_NW) {
              const int r = min(r0 + c, M.n_lr - 1);
              // owner function of factor row r (fn_lr_start is monotone: a short scan from the previous owner would do; rows
              // per function are few, so a binary search over m+1 functions is cheap enough)
              int lo = 0, hi = m + 1;
              while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (M.fn_lr_start[mid] <= r) lo = mid; else hi = mid; }
              const double wk = (c < kc) ? wq[lo] : 0.0;
              const double sw = sqrt(fabs(wk));
              const double *vr = M.lrV + (size_t)r * n;
              if ((tid & 31) == 0) sgn16[c] = wk < 0.0 ? -1.0 : 1.0;
              for (int j = tid & 31; j < lds_tm; j += 32)
                stage[c * lds_tm + j] = (j < n && !(flg[j] & F_FIXED)) ? sw * vr[j] : 0.0;
            }
            if (tid == 0) s_ctr[0] = 0;
            __syncthreads();
            tm_syrk_stage<false>(H, TL, (N + 7) >> 3, stage, lds_tm, KT, s_ctr, sgn16);
            __syncthreads();
          }
          // the carried rows: the two right-hand sides and the equality rows A_e (after the factorisation Y_e = L^-1 A_e')
#pragma unroll 1
          for (int j = tid; j < N; j += IPM_NT) { H[tm_idx(TL, N, j)] = d0[j]; H[tm_idx(TL, N + 1, j)] = d1[j]; }
          for (int e = 0; e < E; ++e) {
            const double *g = grow(M, grad, M.eq_fn[e]);
#pragma unroll 1
            for (int j = tid; j < N; j += IPM_NT) H[tm_idx(TL, N + 2 + e, j)] = (j < n && !(flg[j] & F_FIXED)) ? esc[e] * g[j] : 0.0;
          }
        } else {
        for (int r0 = 0; r0 < M.Rd; r0 += IPM_KC) {
          const int kc = min(IPM_KC, M.Rd - r0);
          for (int c = tid >> 5; c < kc; c += IPM_NW) {   // one warp per staged row
            const int r = M.dense_rows[r0 + c];
            const double sq = sqrt(Ss[r]), f = sq * fsc[r];
            const double *g = grow(M, grad, M.row_fn[r]);
            for (int j = tid & 31; j < npad; j += 32) {
              double vv = 0.0;
              if (j < N && !(flg[j] & F_FIXED)) vv = (j < n) ? f * g[j] : -sq;
              stage[c * npad + j] = vv;
            }
          }
          __syncthreads();
          syrk_tiles(H, 0, N, stage, npad, kc, 1.0);
          __syncthreads();
        }
        // equality rows A_e ride along as extra rows: after the factorisation they hold Y_e = L^-1 A_e'
        for (int e = 0; e < E; ++e) {
          double *Ye = H + tri(N + 2 + e);
          const double *g = grow(M, grad, M.eq_fn[e]);
#pragma unroll 1
          for (int j = tid; j < N; j += IPM_NT) Ye[j] = (j < n && !(flg[j] & F_FIXED)) ? esc[e] * g[j] : 0.0;
        }
        }
        // sparse linear rows: one thread per touched entry of H (fixed summation order)
        for (int sg = tid; sg < M.sp_nseg; sg += IPM_NT) {
          const int i = M.sp_seg_i[sg], k = M.sp_seg_k[sg];
          if ((flg[i] | flg[k]) & F_FIXED) continue;
          double acc = 0.0;
          for (int e = M.sp_seg_start[sg]; e < M.sp_seg_start[sg + 1]; ++e) {
            const int r = M.sp_ent_row[e] & 0xfffffff, kind = M.sp_ent_row[e] >> 28;
            const double t = (kind == 0) ? Ss[r] * fsc[r] * fsc[r] : (kind == 1) ? -Ss[r] * fsc[r] : Ss[r];
            acc = fma(t, M.sp_ent_p[e], acc);
          }
          H[TM ? tm_idx(TL, i, k) : tri(i) + k] += acc;
        }
        __syncthreads();
        ++nfact;
        PH_MARK(2);
        if (TM ? chol_tm_ll(H, TL, N, N + 2 + E, &s_fail) : chol_packed(H, N, N + 2 + E, stage, npad, &s_fail)) { fact_ok = true; break; }
        // inertia correction (first_hessian_perturbation 1e-4, x100 then x8, /3 between iterations)
        if (delta_w == 0.0) delta_w = (delta_w_last == 0.0) ? 1e-4 : fmax(1e-20, delta_w_last / 3.0);
        else delta_w *= (delta_w_last == 0.0 ? 100.0 : 8.0);
        if (delta_w > 1e20) break;
        __syncthreads();
        // the failed factorisation destroyed the right-hand sides: rebuild them
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          double r0 = 0.0, r1 = 0.0;
          if (!(f & F_FIXED)) {
            double g0 = (j < n) ? of * grow(M, grad, 0)[j] : (phase1 ? 1.0 : 0.0);
            if (j < n)
              for (int e = 0; e < E; ++e) g0 = fma(esc[e] * grow(M, grad, M.eq_fn[e])[j], ye[e], g0);
            double a0 = 0.0, a1 = 0.0;
            if (j < n) {
              for (int r = 0; r < R; ++r) {
                const double Jrj = fsc[r] * grow(M, grad, M.row_fn[r])[j];
                a0 = fma(Jrj, Ss[r] * rc[r], a0);
                a1 += Jrj / s[r];
              }
            } else {
              for (int r = 0; r < R; ++r) { a0 -= Ss[r] * rc[r]; a1 -= 1.0 / s[r]; }
            }
            r0 = -g0 - a0; r1 = -a1;
            if (f & F_HASL) r1 += 1.0 / (x[j] - Lr[j]);
            if (f & F_HASU) r1 -= 1.0 / (Ur[j] - x[j]);
          }
          d0[j] = r0; d1[j] = r1;
        }
        __syncthreads();
      }
      if (!fact_ok) { status = ST_computationError; objv = val[0]; with_duals = true; done = true; break; }
      if (delta_w > 0) delta_w_last = delta_w;
      PH_MARK(3);
      if (E > 0) {
        // S = Y Y' (+ tiny regularisation), t0 = Y w0 + c_e, t1 = Y w1: one warp per entry
        const int w = tid >> 5, lane = tid & 31;
        for (int p = w; p < E * E + 2 * E; p += IPM_NW) {
          const double *ra, *rb;
          int ia, ib;     // rows of the carried block: Y_a, and Y_b or a forward-substituted right-hand side
          if (p < E * E) { ia = N + 2 + p / E; ib = N + 2 + p % E; }
          else { const int q = p - E * E; ia = N + 2 + (q >> 1); ib = N + (q & 1); }
          ra = H + tri(ia); rb = (ib >= N + 2) ? H + tri(ib) : (ib == N ? d0 : d1);
          double acc = 0.0;
          if (TM) { for (int j = lane; j < N; j += 32) acc = fma(H[tm_idx(TL, ia, j)], H[tm_idx(TL, ib, j)], acc); }
          else
          for (int j = lane; j < N; j += 32) acc = fma(ra[j], rb[j], acc);
#pragma unroll
          for (int q = 16; q > 0; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
          if (lane == 0) {
            if (p < E * E) Ssc[p] = acc;
            else { const int q = p - E * E; if (q & 1) de1[q >> 1] = acc; else de0[q >> 1] = acc + ce[q >> 1]; }
          }
        }
        __syncthreads();
        if (tid == 0) {
          bool ok = true;
          for (int a = 0; a < E; ++a) Ssc[a * E + a] += 1e-14 * (1.0 + Ssc[a * E + a]);
          for (int j = 0; j < E; ++j) {     // Cholesky of the E x E Schur complement
            double d = Ssc[j * E + j];
            for (int k = 0; k < j; ++k) d -= Ssc[j * E + k] * Ssc[j * E + k];
            if (!(d > 0.0)) { ok = false; break; }
            d = sqrt(d);
            Ssc[j * E + j] = d;
            for (int i = j + 1; i < E; ++i) {
              double vv = Ssc[i * E + j];
              for (int k = 0; k < j; ++k) vv -= Ssc[i * E + k] * Ssc[j * E + k];
              Ssc[i * E + j] = vv / d;
            }
          }
          if (ok) {
            for (int rhs = 0; rhs < 2; ++rhs) {
              double *t = rhs ? de1 : de0;
              for (int i = 0; i < E; ++i) { double vv = t[i]; for (int k = 0; k < i; ++k) vv -= Ssc[i * E + k] * t[k]; t[i] = vv / Ssc[i * E + i]; }
              for (int i = E - 1; i >= 0; --i) { double vv = t[i]; for (int k = i + 1; k < E; ++k) vv -= Ssc[k * E + i] * t[k]; t[i] = vv / Ssc[i * E + i]; }
            }
          }
          s_fail = ok ? 0 : 1;
        }
        __syncthreads();
        if (s_fail) { status = ST_computationError; objv = val[0]; with_duals = true; done = true; break; }
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          double *p0 = TM ? H + tm_idx(TL, N, j) : d0 + j, *p1 = TM ? H + tm_idx(TL, N + 1, j) : d1 + j;
          double w0 = *p0, w1 = *p1;
          for (int e = 0; e < E; ++e) { const double yv = H[TM ? tm_idx(TL, N + 2 + e, j) : tri(N + 2 + e) + j]; w0 = fma(-yv, de0[e], w0); w1 = fma(-yv, de1[e], w1); }
          *p0 = w0; *p1 = w1;
        }
        __syncthreads();
      }
#ifdef BNLP_OLD_BACKSOLVE_TM
      if (TM) backsolve_tm<true>(H, TL, N, d0, d1);   // rows N, N+1 of H -> d0, d1 = H^-1 rhs0, H^-1 rhs1
#else
      if (TM) backsolve32_tm(H, TL, N, d0, d1, red);  // rows N, N+1 of H -> d0, d1 = H^-1 rhs0, H^-1 rhs1
#endif
      else backsolve2(H, N);   // d0 = H^-1 rhs0, d1 = H^-1 rhs1
      PH_MARK(4);

      // ---------------------------------------------------------------- J d0, J d1 per row
      dots_grad(M, grad, d0, d1, gd0, gd1);
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        const int fn = M.row_fn[r];
        jd0[r] = fsc[r] * gd0[fn] - d0[n];
        jd1[r] = fsc[r] * gd1[fn] - d1[n];
      }
      __syncthreads();

      // direction for a given mu, evaluated on the fly:  dx = d0 + mu d1
      // fraction-to-the-boundary + (optional) complementarity at the trial point
      auto ftb = [&](double mu, double tau, double &ap, double &ad) {
        double mn[2] = {-1.0, -1.0};  // max of -alpha  => min alpha
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          const double dxj = d0[j] + mu * d1[j];
          if (f & F_HASL) {
            const double sl = x[j] - Lr[j];
            if (dxj < 0) mn[0] = fmax(mn[0], -(-tau * sl / dxj));
            const double dz = mu / sl - zL[j] - zL[j] / sl * dxj;
            if (dz < 0) mn[1] = fmax(mn[1], -(-tau * zL[j] / dz));
          }
          if (f & F_HASU) {
            const double su = Ur[j] - x[j];
            if (dxj > 0) mn[0] = fmax(mn[0], -(tau * su / dxj));
            const double dz = mu / su - zU[j] + zU[j] / su * dxj;
            if (dz < 0) mn[1] = fmax(mn[1], -(-tau * zU[j] / dz));
          }
        }
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          const double dsr = -rc[r] - (jd0[r] + mu * jd1[r]);
          const double dvr = mu / s[r] - v[r] - Ss[r] * dsr;
          if (dsr < 0) mn[0] = fmax(mn[0], -(-tau * s[r] / dsr));
          if (dvr < 0) mn[1] = fmax(mn[1], -(-tau * v[r] / dvr));
        }
        block_reduce<0, 2>(mn, red);
        ap = fmin(1.0, -mn[0]); ad = fmin(1.0, -mn[1]);
      };
      // ---------------------------------------------------------------- probing oracle: sigma = (mu_aff/mu)^3
      double mu;
      {
        double ap, ad;
        ftb(0.0, 1.0, ap, ad);
        double t[1] = {0.0};
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          const double dxj = d0[j];
          if (f & F_HASL) {
            const double sl = x[j] - Lr[j];
            const double dz = -zL[j] - zL[j] / sl * dxj;
            t[0] += (sl + ap * dxj) * (zL[j] + ad * dz);
          }
          if (f & F_HASU) {
            const double su = Ur[j] - x[j];
            const double dz = -zU[j] + zU[j] / su * dxj;
            t[0] += (su - ap * dxj) * (zU[j] + ad * dz);
          }
        }
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          const double dsr = -rc[r] - jd0[r];
          const double dvr = -v[r] - Ss[r] * dsr;
          t[0] += (s[r] + ap * dsr) * (v[r] + ad * dvr);
        }
        block_reduce<1, 0>(t, red);
        const double mu_aff = nz > 0 ? t[0] / nz : 0.0;
        double sigma = avg_compl > 0 ? pow(mu_aff / avg_compl, 3.0) : o.sigma_min;
        sigma = fmin(fmax(sigma, o.sigma_min), o.sigma_max);
        mu = fmin(fmax(sigma * avg_compl, o.mu_min), mu_max);
        mu = fmax(mu, fmin(avg_compl, o.mu_safeguard * fmax(cviol, dual_inf / s_d)));
      }
      const double tau = fmax(o.tau_min, 1.0 - mu);
      double ap_max, ad;
      ftb(mu, tau, ap_max, ad);
      // final direction
#pragma unroll 1
      for (int j = tid; j < N; j += IPM_NT) dx[j] = (flg[j] & F_FIXED) ? 0.0 : d0[j] + mu * d1[j];
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        ds[r] = -rc[r] - (jd0[r] + mu * jd1[r]);
        dv[r] = mu / s[r] - v[r] - Ss[r] * ds[r];
      }
#pragma unroll 1
      for (int k = tid; k <= m; k += IPM_NT) gd0[k] = gd0[k] + mu * gd1[k];   // gd = grad_k . dx
      __syncthreads();
      const double gd_obj_total = gd0[0];   // directional derivative of f uses the total gradient
      if (NTm > 0) {   // the closed-form trial values use the quadratic part only: remove the terms' share of grad.dx
        __syncthreads();
#pragma unroll 1
        for (int kf = tid; kf <= m; kf += IPM_NT) {
          double acc = 0.0;
          for (int e = M.fn_term_start[kf]; e < M.fn_term_start[kf + 1]; ++e) {
            const int t = M.fn_term_list[e], v0 = M.term_var_start[t], nv = M.term_var_start[t + 1] - v0;
            const double *rec = tgl + (size_t)t * TREC;
            for (int a = 0; a < nv; ++a) acc = fma(rec[1 + a], dx[M.term_var[v0 + a]], acc);
          }
          gd0[kf] -= acc;
        }
        __syncthreads();
      }

      PH_MARK(5);
      // ---------------------------------------------------------------- u_k = Q_k dx, du_k = dx.u_k
      pass_U<MAXC>(M, dx, uq, du, qfn, nqf);
      if (TM && M.n_lr > 0) pass_U_lowrank(M, dx, uq, du, lrs);

      PH_MARK(6);
      // ---------------------------------------------------------------- filter line search
      const double dxt = phase1 ? dx[n] : 0.0;
      auto merit = [&](double a, double &theta, double &phi) {
        double t2[2] = {0.0, 0.0};
        const double t = tcur + a * dxt;
        if (NTm > 0) {   // element terms at the trial point (values only); Sx is free here and holds x + a dx
#pragma unroll 1
          for (int j = tid; j < n; j += IPM_NT) Sx[j] = x[j] + a * dx[j];
          __syncthreads();
          HDual st[TS];
          HDual r;
          for (int tt = tid; tt < NTm; tt += IPM_NT) {
            eval_term(M, tt, Sx, st, r, 0);
            tgl[(size_t)tt * TREC + TREC - 1] = r.v;       // trial value, summed per function in tape order below
          }
          __syncthreads();
#pragma unroll 1
          for (int k = tid; k <= m; k += IPM_NT) {
            double tv = -tval[k];
            for (int e = M.fn_term_start[k]; e < M.fn_term_start[k + 1]; ++e) tv += tgl[(size_t)M.fn_term_list[e] * TREC + TREC - 1];
            tvt[k] = tv;
          }
          __syncthreads();
        }
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          const int fn = M.row_fn[r];
          const double hv = fsc[r] * (val[fn] + a * gd0[fn] + 0.5 * a * a * du[fn] + (NTm > 0 ? tvt[fn] : 0.0)) - t;
          const double sv = s[r] + a * ds[r];
          t2[0] += fabs(hv + sv - bb[r]);
          t2[1] -= mu * log(sv);
        }
#pragma unroll 1
        for (int e = tid; e < E; e += IPM_NT) {
          const int fn = M.eq_fn[e];
          t2[0] += fabs(esc[e] * (val[fn] + a * gd0[fn] + 0.5 * a * a * du[fn] + (NTm > 0 ? tvt[fn] : 0.0) - M.eq_bnd[e]));
        }
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          const double xv = x[j] + a * dx[j];
          if (f & F_HASL) t2[1] -= mu * log(xv - Lr[j]);
          if (f & F_HASU) t2[1] -= mu * log(Ur[j] - xv);
        }
        block_reduce<2, 0>(t2, red);
        theta = t2[0];
        phi = t2[1] + (phase1 ? t : sf * (val[0] + a * gd0[0] + 0.5 * a * a * du[0] + (NTm > 0 ? tvt[0] : 0.0)));
      };
      double theta0, phi0;
      merit(0.0, theta0, phi0);
      double dphi;
      {
        double t[1] = {0.0};
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) t[0] -= mu * ds[r] / s[r];
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          if (f & F_HASL) t[0] -= mu * dx[j] / (x[j] - Lr[j]);
          if (f & F_HASU) t[0] += mu * dx[j] / (Ur[j] - x[j]);
        }
        block_reduce<1, 0>(t, red);
        dphi = t[0] + (phase1 ? dxt : sf * gd_obj_total);
      }
      const double theta_min = o.theta_min_fact * fmax(1.0, theta_init);
      const double theta_max = o.theta_max_fact * fmax(1.0, theta_init);
      double amin = o.gamma_theta;
      if (dphi < 0) {
        amin = fmin(amin, o.gamma_phi * theta0 / (-dphi));
        if (theta0 <= theta_min) amin = fmin(amin, o.delta * pow(theta0, o.s_theta) / pow(-dphi, o.s_phi));
      }
      amin *= o.alpha_min_frac;
      double alpha = ap_max;
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        double th_t, ph_t;
        merit(alpha, th_t, ph_t);
        const double fuzz = 10.0 * 2.220446049250313e-16 * fabs(phi0);
        const bool sw = dphi < 0 && theta0 <= theta_min && alpha * pow(-dphi, o.s_phi) > o.delta * pow(theta0, o.s_theta);
        bool ok;
        if (sw) ok = ph_t - phi0 <= o.eta_phi * alpha * dphi + fuzz;
        else ok = (th_t <= (1 - o.gamma_theta) * theta0) || (ph_t - phi0 <= -o.gamma_phi * theta0 + fuzz);
        if (!(th_t <= theta_max)) ok = false;
        if (!isfinite(ph_t)) ok = false;
        if (ok) { accepted = true; break; }
        alpha *= o.alpha_red;
        if (alpha < amin) break;
      }
      if (!accepted) {
        if (phase1 || phase1_entries >= o.max_phase1) { status = ST_iterationLimit; objv = val[0]; with_duals = true; done = true; break; }
        enter_phase1();
        ++iters;
        continue;
      }
      a_last1 = a_last0; a_last0 = alpha;

      PH_MARK(7);
      // ---------------------------------------------------------------- accept the step
      __syncthreads();
      terms_remove();
#pragma unroll 1
      for (int j = tid; j < N; j += IPM_NT) {
        const int f = flg[j];
        if (f & F_FIXED) continue;
        const double dxj = dx[j];
        const double xo = x[j];
        const double xn = xo + alpha * dxj;
        if (f & F_HASL) {
          const double sl = xo - Lr[j];
          const double dz = mu / sl - zL[j] - zL[j] / sl * dxj;
          const double sn = xn - Lr[j];
          zL[j] = fmin(fmax(zL[j] + ad * dz, mu / (o.kappa_sigma * sn)), o.kappa_sigma * mu / sn);
        }
        if (f & F_HASU) {
          const double su = Ur[j] - xo;
          const double dz = mu / su - zU[j] + zU[j] / su * dxj;
          const double sn = Ur[j] - xn;
          zU[j] = fmin(fmax(zU[j] + ad * dz, mu / (o.kappa_sigma * sn)), o.kappa_sigma * mu / sn);
        }
        x[j] = xn;
      }
#pragma unroll 1
      for (int r = tid; r < R; r += IPM_NT) {
        const double sn = s[r] + alpha * ds[r];
        s[r] = sn;
        v[r] = fmin(fmax(v[r] + ad * dv[r], mu / (o.kappa_sigma * sn)), o.kappa_sigma * mu / sn);
      }
#pragma unroll 1
      for (int e = tid; e < E; e += IPM_NT) ye[e] += alpha * (de0[e] + mu * de1[e]);
#pragma unroll 1
      for (int k = tid; k <= m; k += IPM_NT) val[k] += alpha * gd0[k] + 0.5 * alpha * alpha * du[k];
      for (int k = 0; k <= m; ++k) {
        if (M.qidx[k] < 0) continue;
        double *g = grad + (size_t)k * n;
        const double *u = uq + (size_t)k * n;
#pragma unroll 1
        for (int j = tid; j < n; j += IPM_NT) g[j] += alpha * u[j];
      }
      __syncthreads();
      terms_add();
      ++iters;
      PH_MARK(8);
    }

    // ------------------------------------------------------------------ results (TMINLP2TNLP layout)
    __syncthreads();
    if (status == ST_provenInfeasible && !(inconsistent || all_fixed)) {
      double t[1] = {0.0};
#pragma unroll 1
      for (int i = tid; i < m; i += IPM_NT) {
        const double g = val[i + 1];
        if (M.gu[i] < kInfD) t[0] = fmax(t[0], g - M.gu[i]);
        if (M.gl[i] > -kInfD) t[0] = fmax(t[0], M.gl[i] - g);
      }
      block_reduce<0, 1>(t, red);
      objv = t[0];   // max-norm violation, BonTMINLP2TNLP.cpp:503-505
    }
    if (tid == 0) {
      A.OBJ[child] = objv; A.STATUS[child] = status; A.ITERS[child] = iters;
      if (A.phase_cycles)
        for (int q = 0; q < 10; ++q) atomicAdd((unsigned long long *)&A.phase_cycles[q], (unsigned long long)ph[q]);
      if (A.stats) { atomicAdd((unsigned long long *)&A.stats[0], (unsigned long long)iters);
                     atomicAdd((unsigned long long *)&A.stats[1], (unsigned long long)nfact); }
    }
    if (A.X) for (int j = tid; j < n; j += IPM_NT) A.X[(size_t)child * n + j] = all_fixed ? xl[j] : fmin(fmax(x[j], fmin(xl[j], xu[j])), fmax(xl[j], xu[j]));
    if (A.G) for (int i = tid; i < m; i += IPM_NT) A.G[(size_t)child * m + i] = val[i + 1];
    if (A.DUALS) {
      double *D = A.DUALS + (size_t)child * (2 * n + m);
#pragma unroll 1
      for (int j = tid; j < n; j += IPM_NT) {
        D[j] = with_duals ? zL[j] / sf : 0.0;
        D[n + j] = with_duals ? zU[j] / sf : 0.0;
      }
#pragma unroll 1
      for (int i = tid; i < m; i += IPM_NT) {
        double lam = 0.0;
        if (with_duals) {
          const int ru = M.fn_row_up[i + 1], rl = M.fn_row_lo[i + 1];
          if (ru >= 0) lam += v[ru] * fsc[ru] / sf;
          if (rl >= 0) lam += v[rl] * fsc[rl] / sf;
          const int ei = M.fn_eq[i + 1];
          if (ei >= 0) lam += ye[ei] * esc[ei] / sf;
        }
        D[2 * n + i] = lam;
      }
    }
  }
}

}  // namespace BNLP_NS
