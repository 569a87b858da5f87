// Cooperative (grid-lockstep) batched interior-point kernel — the fast path for n <= 256 models whose condensed
// matrix fits shared memory.
//
// Same algorithm and same per-child arithmetic as ipm_kernel (bnlp_ipm.cuh), different execution shape:
//   * one persistent CTA per SM owns ONE child relaxation at a time (its condensed matrix H and all iterate vectors stay
//     in shared memory for the whole solve; children are pulled from an atomic queue);
//   * the two contractions with the SHARED problem data, which dominated the first kernel because every CTA re-read
//     all quadratic blocks from L2 (6.6 MB + 13.1 MB per iteration at C3), are done COOPERATIVELY by the whole grid for
//     all resident children at once, as real dense contractions on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA):
//         W pass : W_b = sum_q y_{b,q} Q_q      C[16 children x 8 elements]  += Y[16 x 4] * Qp[4 x 8]
//         U pass : u_{b,q} = Q_q dx_b           C[8 rows x 16 children]      += Qf[8 x 4] * DX[4 x 16]
//     CTA (g, s) serves child group g (16 children) on slice s of the packed elements / matrix rows, so each quadratic
//     block element crosses L2->SM once per 16 children instead of once per child;
//   * four grid barriers per step: publish weights | W pass | factor+direction, publish dx | U pass | line search.
//   * the n x n work that stays per child (J' Sigma J, Cholesky trailing updates) also runs on DMMA from shared memory.
#pragma once
#include <cooperative_groups.h>

#include "bnlp_ipm.cuh"

namespace bnlp {
namespace cg = cooperative_groups;

constexpr int COOP_RB = 16;   // children per cooperative group (two m8 tiles)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- H[i,k] += sign * sum_c M[c][i] M[c][k], lo <= k <= i < hi, 8x8 tiles on the FP64 tensor pipe.
// M: K rows (K multiple of 4) of length ld in shared memory, zero outside [lo,hi) within the touched tile range.
__device__ __forceinline__ void syrk_dmma(double *H, int lo, int hi, const double *M, int ld, int K, double sign) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = lo >> 3, nb = ((hi + 7) >> 3) - b0;
  const int ntiles = (nb * (nb + 1)) >> 1;
  const int lr = lane >> 2, lc = lane & 3;
  for (int t = w; t < ntiles; t += IPM_NW) {
    int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (tri(bi + 1) <= t) ++bi;
    while (tri(bi) > t) --bi;
    const int bk = t - tri(bi);
    const int i0 = (b0 + bi) << 3, k0 = (b0 + bk) << 3;
    double c0 = 0.0, c1 = 0.0;
    const double *Ma = M + lc * ld + i0 + lr, *Mb = M + lc * ld + k0 + lr;
    for (int c = 0; c < K; c += 4) {
      dmma(c0, c1, Ma[c * ld], Mb[c * ld]);
    }
    const int i = i0 + lr, k = k0 + 2 * lc;
    if (i >= lo && i < hi) {
      double *Hi = H + tri(i);
      if (k >= lo && k <= i) Hi[k] += sign * c0;
      if (k + 1 >= lo && k + 1 <= i) Hi[k + 1] += sign * c1;
    }
  }
}

// ---- blocked Cholesky, 8-column panels; rows [N, Nrows) carry right-hand sides (forward substitution for free).
// After the call the 8x8 DIAGONAL BLOCKS of H hold the INVERSES of the diagonal blocks of L (lower triangular), the
// off-diagonal part holds L.  Returns false (block-uniform) on a non-positive pivot.
__device__ __forceinline__ bool chol_dmma(double *H, int N, int Nrows, double *stage, int ld, int *fail) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j0 = 0; j0 < N; j0 += 8) {
    const int nb = min(8, N - j0);
    __syncthreads();
    if (w == 0) {
      // lane r (< 8) owns row r of the diagonal block
      double a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = (lane < nb && k <= lane) ? H[tri(j0 + lane) + j0 + k] : ((lane == k) ? 1.0 : 0.0);
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double rs = rsqrt(d);
        const double colj = (lane > j) ? a[j] * rs : ((lane == j) ? d * rs : 0.0);
        if (lane >= j) a[j] = colj;
#pragma unroll
        for (int k = j + 1; k < 8; ++k) {
          const double ck = __shfl_sync(0xffffffffu, colj, k);
          if (lane >= k) a[k] = fma(-colj, ck, a[k]);
        }
      }
      // inverse X = L^-1 (row r on lane r): X[r][c] = (delta_rc - sum_{k<r} L[r][k] X[k][c]) / L[r][r]
      double x[8], acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { acc[c] = (lane == c) ? 1.0 : 0.0; x[c] = 0.0; }
      double mydiag = 1.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (lane == k) mydiag = a[k];
      const double myrinv = 1.0 / mydiag;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // row k is final once all rows < k have been pushed: finalise on lane k
        if (lane == k) {
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = (c <= k) ? acc[c] * myrinv : 0.0;
        }
#pragma unroll
        for (int c = 0; c <= k; ++c) {
          const double xkc = __shfl_sync(0xffffffffu, x[c], k);
          if (lane > k) acc[c] = fma(-a[k], xkc, acc[c]);
        }
      }
      if (lane < nb) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= lane) H[tri(j0 + lane) + j0 + c] = x[c];
      }
      if (lane == 0) *fail = ok ? 0 : 1;
    }
    __syncthreads();
    if (*fail) return false;
    const int lo = j0 + nb;
    if (lo >= Nrows) break;
    const int lo8 = lo & ~7, hi8 = (Nrows + 7) & ~7;
    // TRSM as a multiply with the inverse block: x_c = sum_{k<=c} a_k X[c][k]; also stage the panel for the update
    for (int i = lo8 + threadIdx.x; i < hi8; i += IPM_NT) {
      double xr[8];
      if (i >= lo && i < Nrows) {
        double *Hi = H + tri(i) + j0;
        double ar[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) ar[c] = (c < nb) ? Hi[c] : 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double v = 0.0;
          if (c < nb) {
            const double *Xc = H + tri(j0 + c) + j0;
#pragma unroll
            for (int k = 0; k <= c; ++k) v = fma(ar[k], Xc[k], v);
          }
          xr[c] = v;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < nb) Hi[c] = xr[c];
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) xr[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) stage[c * ld + i] = xr[c];
    }
    __syncthreads();
    syrk_dmma(H, lo, Nrows, stage, ld, 8, -1.0);
  }
  __syncthreads();
  return true;
}

// ---- backward substitution L' x = w for the right-hand sides in rows N and N+1 (warps 0 and 1), 8-row blocks,
// diagonal blocks given as inverses: x_blk = X' w_blk
__device__ __forceinline__ void backsolve_inv(double *H, int N) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < 2) {
    double *b = H + tri(N + w);
    const int nblk = (N + 7) >> 3;
    for (int blk = nblk - 1; blk >= 0; --blk) {
      const int i0 = blk << 3, nb = min(8, N - i0);
      double xc = 0.0;   // lane c < nb computes x_c = sum_{k>=c} X[k][c] w_k
      if (lane < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nb && k >= lane) xc = fma(H[tri(i0 + k) + i0 + lane], b[i0 + k], xc);
      }
      double xr[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xr[c] = __shfl_sync(0xffffffffu, xc, c);
      __syncwarp();
      if (lane < nb) b[i0 + lane] = xc;
      for (int k = lane; k < i0; k += 32) {
        double v = b[k];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < nb) v = fma(-H[tri(i0 + c) + k], xr[c], v);
        b[k] = v;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

struct CoopBufs {
  double *xw;       // G x nqpad : Hessian weights published by every CTA
  double *Wbuf;     // G x P     : W_b produced by the cooperative W pass
  double *xd;       // G x ldd   : direction (or start point at initialisation) published by every CTA
  int *idle_cnt;    // [2]
  int nqpad, ldd;
};

enum { CS_NEED_CHILD = 0, CS_INIT = 1, CS_ITER = 2, CS_IDLE = 3 };

__global__ void __launch_bounds__(IPM_NT, 1) ipm_coop_kernel(DevModel M, IpmArgs A, CoopBufs CB) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double smem[];
  const int n = M.n, m = M.m, N = M.N, R = M.R, npad = M.npad;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const IpmOpt &o = A.o;
  const int G = gridDim.x, cta = blockIdx.x;
  // ---- carve shared memory (same budget as ipm_kernel; the staging rows get a bank-friendly stride)
  const int lds = ((N + 2 + 7) & ~7) + 4;   // >= roundup8(N+2), stride = 4 (mod 8) doubles: conflict-free DMMA fragments
  double *sp = smem;
  const int Haug = tri(N + 2);
  double *H = sp; sp += (Haug + 1) & ~1;
  double *stage = sp; sp += (IPM_NW + 1) * lds;
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
  double *fsc = sp; sp += Rp;
  double *jd0 = sp; sp += Rp;
  double *jd1 = sp; sp += Rp;
  double *val = sp; sp += Kp;
  double *gd0 = sp; sp += Kp;
  double *gd1 = sp; sp += Kp;
  double *du = sp; sp += Kp;
  double *wq = sp; sp += 2 * Kp;
  double *red = sp; sp += IPM_NW * 12;
  int *flg = reinterpret_cast<int *>(sp);
  int *fixlist = flg + ((N + 1) & ~1);
  int *qblk = fixlist + ((N + 1) & ~1);
  int *qfn = qblk + Kp;
  __shared__ int s_child, s_nfix, s_fail, s_nqf;
  double *ws = A.ws + (size_t)cta * A.ws_stride;
  double *grad = ws;
  double *uq = grad + (size_t)(m + 1) * n;
  double *d0 = H + tri(N), *d1 = H + tri(N + 1);
  // the three per-row scalars used by the J' products reuse ds/dv/jd0 (free at that point of the iteration)
  double *pv = ds, *p0 = dv, *p1 = jd0;

  if (tid == 0) {
    int c = 0;
    for (int k = 0; k <= m; ++k)
      if (M.qidx[k] >= 0) qfn[c++] = k;
    s_nqf = c;
  }
  for (int k = tid; k <= m; k += IPM_NT) du[k] = 0.0;
  __syncthreads();
  const int nqf = s_nqf;
  const int nqpad = CB.nqpad, ldd = CB.ldd;
  // cooperative decomposition: CTA = (child group g, slice s)
  const int NG = (G + COOP_RB - 1) / COOP_RB;
  const int NS = max(1, G / NG);
  const int my_g = cta / NS, my_s = cta % NS;
  const bool coop_on = my_g < NG;
  const int P = M.P;

  // ---- per-child state
  int state = CS_NEED_CHILD;
  int child = -1;
  const double *xl = nullptr, *xu = nullptr, *x0 = nullptr;
  int status = 0, iters = 0;
  double objv = 0.0, sf = 1.0, badw = 0.0;
  bool with_duals = false, inconsistent = false, all_fixed = false, phase1 = false, retry = false;
  int phase1_entries = 0, iters_main = 0, acceptable_cnt = 0, fact_attempts = 0;
  double theta_init = -1.0, mu_max = -1.0, delta_w_last = 0.0, delta_w = 0.0;
  double th_hist[4] = {0, 0, 0, 0}, a_last0 = 1.0, a_last1 = 1.0;
  long long nfact = 0;
  // values carried from the top of an iteration to its later phases
  double it_nz = 0, it_avg_compl = 0, it_cviol = 0, it_dual_inf = 0, it_s_d = 1, it_viol_max = 0, it_of = 0, it_tcur = 0;
  double it_mu = 0, it_ap_max = 0, it_ad = 0;

  auto init_slacks_mults = [&]() {
    for (int r = tid; r < R; r += IPM_NT) {
      const double h = fsc[r] * val[M.row_fn[r]] - (phase1 ? x[n] : 0.0);
      s[r] = fmax(bb[r] - h, o.bound_push * fmax(1.0, fabs(bb[r])));
      v[r] = 1.0;
    }
    for (int j = tid; j < N; j += IPM_NT) {
      const int f = flg[j];
      zL[j] = (!(f & F_FIXED) && (f & F_HASL)) ? 1.0 : 0.0;
      zU[j] = (!(f & F_FIXED) && (f & F_HASU)) ? 1.0 : 0.0;
    }
    __syncthreads();
  };
  auto enter_phase1 = [&]() {
    __syncthreads();
    phase1 = true; phase1_entries++;
    const double vm = fmax(0.0, it_viol_max), T = fmax(1.0, vm);
    if (tid == 0) { flg[n] &= ~F_FIXED; Lr[n] = -T; x[n] = 1.1 * vm + 0.1 * T; }
    __syncthreads();
    init_slacks_mults();
    theta_init = -1; mu_max = -1; acceptable_cnt = 0;
  };

  // ---- fetch + set up the next child; returns false when the queue is empty
  auto setup_child = [&]() -> bool {
    __syncthreads();
    if (tid == 0) s_child = atomicAdd(A.queue, 1);
    __syncthreads();
    child = s_child;
    if (child >= A.B) return false;
    xl = A.XL + (size_t)child * n; xu = A.XU + (size_t)child * n;
    x0 = A.X0 + (size_t)child * A.x0_stride;
    status = 0; iters = 0; objv = 0.0; sf = 1.0; with_duals = false; phase1 = false; retry = false;
    phase1_entries = 0; iters_main = 0; acceptable_cnt = 0; fact_attempts = 0;
    theta_init = -1.0; mu_max = -1.0; delta_w_last = 0.0; delta_w = 0.0;
    th_hist[0] = th_hist[1] = th_hist[2] = th_hist[3] = 0.0; a_last0 = a_last1 = 1.0; nfact = 0;
    double acc0 = 0.0, acc1 = 0.0, bw = 0.0;
    for (int j = tid; j < N; j += IPM_NT) {
      int f = 0;
      double xx;
      if (j < n) {
        const double l = xl[j], u = xu[j];
        if (l > -kInfD) f |= F_HASL;
        if (u < kInfD) f |= F_HASU;
        if (u < l) { acc0 += 1.0; bw = fmax(bw, l - u); }
        if (u <= l) f |= F_FIXED;
        acc1 = fmax(acc1, u - l);
        xx = x0[j];
        double lr = l, ur = u;
        if (f & F_FIXED) { xx = l; }
        else {
          if (f & F_HASL) lr = l - o.bound_relax * fmax(1.0, fabs(l));
          if (f & F_HASU) ur = u + o.bound_relax * fmax(1.0, fabs(u));
          if ((f & F_HASL) && (f & F_HASU)) {
            const double pl = fmin(o.bound_push * fmax(1.0, fabs(lr)), o.bound_frac * (ur - lr));
            const double pu = fmin(o.bound_push * fmax(1.0, fabs(ur)), o.bound_frac * (ur - lr));
            xx = fmin(fmax(xx, lr + pl), ur - pu);
          } else if (f & F_HASL) xx = fmax(xx, lr + o.bound_push * fmax(1.0, fabs(lr)));
          else if (f & F_HASU) xx = fmin(xx, ur - o.bound_push * fmax(1.0, fabs(ur)));
        }
        Lr[j] = lr; Ur[j] = ur;
      } else {
        f = F_HASL | F_FIXED; xx = 0.0; Lr[j] = -o.bound_relax; Ur[j] = kInfD;
      }
      x[j] = xx; flg[j] = f; zL[j] = 0.0; zU[j] = 0.0;
    }
    double r3[3] = {acc0, acc1, bw};
    block_reduce<1, 2>(r3, red);
    inconsistent = r3[0] > 0.0;
    all_fixed = !inconsistent && r3[1] <= 1e-5;
    badw = r3[2];
    if (inconsistent) {
      for (int j = tid; j < n; j += IPM_NT) {
        const double lo = fmin(xl[j], xu[j]), hi = fmax(xl[j], xu[j]);
        x[j] = fmin(fmax(x0[j], lo), hi);
      }
    } else if (all_fixed) {
      for (int j = tid; j < n; j += IPM_NT) x[j] = xl[j];
    }
    __syncthreads();
    if (tid == 0) {
      int f = 0;
      for (int j = 0; j < N; ++j)
        if (flg[j] & F_FIXED) fixlist[f++] = j;
      s_nfix = f;
    }
    __syncthreads();
    return true;
  };

  // ---- after the U pass that evaluated Q_k x: raw values/gradients, scaling, rows, slacks.  Returns true if the
  // child is already finished (all-fixed short cut / crossed bounds).
  auto finish_init = [&]() -> bool {
    for (int k = wid; k <= m; k += IPM_NW) {
      const double *a = M.A + (size_t)k * n;
      double *g = grad + (size_t)k * n;
      const bool hq = M.qidx[k] >= 0;
      const double *u = uq + (size_t)k * n;
      double acc = 0.0;
      for (int j = lane; j < n; j += 32) {
        const double aj = a[j], uj = hq ? __ldcg(u + j) : 0.0;
        g[j] = aj + uj;
        acc = fma(x[j], aj + 0.5 * uj, acc);
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
      if (lane == 0) val[k] = M.c[k] + acc;
    }
    __syncthreads();
    if (inconsistent || all_fixed) {
      double t = 0.0, infe = 0.0;
      for (int i = tid; i < m; i += IPM_NT) {
        const double g = val[i + 1];
        if (M.gu[i] < kInfD) t = fmax(t, g - M.gu[i]);
        if (M.gl[i] > -kInfD) t = fmax(t, M.gl[i] - g);
        if (g - M.gl[i] < -1e-7 || g - M.gu[i] > 1e-7) infe = 1.0;
      }
      double r2[2] = {t, infe};
      block_reduce<0, 2>(r2, red);
      if (inconsistent) { status = ST_provenInfeasible; objv = fmax(badw, r2[0]); }
      else { status = (r2[1] > 0.0) ? ST_provenInfeasible : ST_solvedOptimal; objv = val[0]; }
      return true;
    }
    for (int k = wid; k <= m; k += IPM_NW) {
      const double *g = grad + (size_t)k * n;
      double gm = 0.0;
      for (int j = lane; j < n; j += 32)
        if (!(flg[j] & F_FIXED)) gm = fmax(gm, fabs(g[j]));
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, q));
      if (lane == 0) wq[k] = (gm > o.scal_max_grad) ? o.scal_max_grad / gm : 1.0;
    }
    __syncthreads();
    sf = wq[0];
    for (int r = tid; r < R; r += IPM_NT) {
      const int fn = M.row_fn[r];
      const double sc = wq[fn], sg = M.row_sgn[r];
      const double b = sc * M.row_bnd[r];
      fsc[r] = sc * sg;
      bb[r] = b + o.bound_relax * fmax(1.0, fabs(b));
    }
    __syncthreads();
    init_slacks_mults();
    return false;
  };

  auto write_results = [&]() {
    __syncthreads();
    if (status == ST_provenInfeasible && !(inconsistent || all_fixed)) {
      double t[1] = {0.0};
      for (int i = tid; i < m; i += IPM_NT) {
        const double g = val[i + 1];
        if (M.gu[i] < kInfD) t[0] = fmax(t[0], g - M.gu[i]);
        if (M.gl[i] > -kInfD) t[0] = fmax(t[0], M.gl[i] - g);
      }
      block_reduce<0, 1>(t, red);
      objv = t[0];
    }
    if (tid == 0) {
      A.OBJ[child] = objv; A.STATUS[child] = status; A.ITERS[child] = iters;
      if (A.stats) { atomicAdd((unsigned long long *)&A.stats[0], (unsigned long long)iters);
                     atomicAdd((unsigned long long *)&A.stats[1], (unsigned long long)nfact); }
    }
    if (A.X) for (int j = tid; j < n; j += IPM_NT)
      A.X[(size_t)child * n + j] = all_fixed ? xl[j] : fmin(fmax(x[j], fmin(xl[j], xu[j])), fmax(xl[j], xu[j]));
    if (A.G) for (int i = tid; i < m; i += IPM_NT) A.G[(size_t)child * m + i] = val[i + 1];
    if (A.DUALS) {
      double *D = A.DUALS + (size_t)child * (2 * n + m);
      for (int j = tid; j < n; j += IPM_NT) {
        D[j] = with_duals ? zL[j] / sf : 0.0;
        D[n + j] = with_duals ? zU[j] / sf : 0.0;
      }
      for (int i = tid; i < m; i += IPM_NT) {
        double lam = 0.0;
        if (with_duals) {
          const int ru = M.fn_row_up[i + 1], rl = M.fn_row_lo[i + 1];
          if (ru >= 0) lam += v[ru] * fsc[ru] / sf;
          if (rl >= 0) lam += v[rl] * fsc[rl] / sf;
        }
        D[2 * n + i] = lam;
      }
    }
    __syncthreads();
  };

  // ---- top of an iteration: residuals, J' Sigma J (speculatively, before the termination test), error measures,
  // termination tests, phase switches.  Returns 0 = proceed to factorisation, 1 = re-evaluate (phase switch), 2 = done.
  auto top_of_iteration = [&]() -> int {
    const double of = phase1 ? 0.0 : sf;
    const double tcur = phase1 ? x[n] : 0.0;
    it_of = of; it_tcur = tcur;
    double th_p = 0, sumv_p = 0, cs_p = 0, vl1_p = 0, ci_p = 0, cv_p = 0, cvu_p = 0, vmax_p = 0, viol_p = -1e300;
    for (int r = tid; r < R; r += IPM_NT) {
      const double hraw = fsc[r] * val[M.row_fn[r]];
      const double res = hraw - tcur + s[r] - bb[r];
      rc[r] = res;
      const double ss = v[r] / s[r], sq = sqrt(ss);
      Ss[r] = ss;
      pv[r] = v[r] / sq; p0[r] = sq * res; p1[r] = 1.0 / (sq * s[r]);
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
    // H := 0 (rows < N), then J' Sigma_s J by staged rank-16 DMMA updates; the same staged rows give J'v, J'(Ss rc), J'(1/s)
    for (int e = tid; e < tri(N); e += IPM_NT) H[e] = 0.0;
    double jv = 0.0, a0 = 0.0, a1 = 0.0;
    __syncthreads();
    for (int r0 = 0; r0 < R; r0 += IPM_KC) {
      const int kc = min(IPM_KC, R - r0), kc4 = (kc + 3) & ~3;
      for (int c = wid; c < kc4; c += IPM_NW) {
        const int r = r0 + c;
        if (c < kc) {
          const double sq = sqrt(Ss[r]), f = sq * fsc[r];
          const double *g = grad + (size_t)M.row_fn[r] * n;
          for (int j = lane; j < lds; j += 32) {
            double vv = 0.0;
            if (j < N && !(flg[j] & F_FIXED)) vv = (j < n) ? f * g[j] : -sq;
            stage[c * lds + j] = vv;
          }
        } else {
          for (int j = lane; j < lds; j += 32) stage[c * lds + j] = 0.0;
        }
      }
      __syncthreads();
      if (tid < N) {
        for (int c = 0; c < kc; ++c) {
          const double sv = stage[c * lds + tid];
          jv = fma(sv, pv[r0 + c], jv);
          a0 = fma(sv, p0[r0 + c], a0);
          a1 = fma(sv, p1[r0 + c], a1);
        }
      }
      syrk_dmma(H, 0, N, stage, lds, kc4, 1.0);
      __syncthreads();
    }
    // sweep over variables: Lagrangian gradient, right-hand sides (into the augmented rows), Sigma_x
    double di_p = 0, sumz_p = 0, nz_p = 0;
    for (int j = tid; j < N; j += IPM_NT) {   // N <= IPM_NT: one variable per thread, jv/a0/a1 belong to j = tid
      const int f = flg[j];
      double r0v = 0.0, r1v = 0.0, sx = 0.0;
      if (!(f & F_FIXED)) {
        const double g0 = (j < n) ? of * grad[j] : (phase1 ? 1.0 : 0.0);
        const double gl = g0 + jv;
        di_p = fmax(di_p, fabs(gl - zL[j] + zU[j]));
        r0v = -g0 - a0;
        r1v = -a1;
        if (f & F_HASL) {
          const double sl = x[j] - Lr[j];
          ci_p = fmax(ci_p, sl * zL[j]); cs_p += sl * zL[j]; sumz_p += zL[j]; nz_p += 1.0;
          sx += zL[j] / sl; r1v += 1.0 / sl;
        }
        if (f & F_HASU) {
          const double su = Ur[j] - x[j];
          ci_p = fmax(ci_p, su * zU[j]); cs_p += su * zU[j]; sumz_p += zU[j]; nz_p += 1.0;
          sx += zU[j] / su; r1v -= 1.0 / su;
        }
      }
      Sx[j] = sx;
      d0[j] = r0v;
      d1[j] = r1v;
    }
    double all[12] = {th_p, sumv_p, cs_p, vl1_p, sumz_p, nz_p, ci_p, cv_p, cvu_p, vmax_p, viol_p, di_p};
    block_reduce<6, 6>(all, red);
    const double th = all[0], sum_v = all[1], vl1 = all[3];
    const double nz = all[5] + (double)R, sum_z = all[4] + sum_v;
    const double compl_inf = all[6], cviol = all[7], cviol_unsc = all[8], ymax = all[9], viol_max = all[10], dual_inf = all[11];
    const double avg_compl = nz > 0 ? all[2] / nz : 0.0;
    if (theta_init < 0) theta_init = th;
    const double s_d = fmax(o.s_max, (sum_v + sum_z) / fmax(1.0, (double)R + nz)) / o.s_max;
    const double s_c = fmax(o.s_max, sum_z / fmax(1.0, nz)) / o.s_max;
    const double E0 = fmax(fmax(dual_inf / s_d, cviol), compl_inf / s_c);
    if (mu_max < 0) mu_max = fmin(o.mu_max, o.mu_max_fact * fmax(avg_compl, 1e-300));
    it_nz = nz; it_avg_compl = avg_compl; it_cviol = cviol; it_dual_inf = dual_inf; it_s_d = s_d; it_viol_max = viol_max;

    const double osc = phase1 ? 1.0 : sf;
    const bool conv = E0 <= o.tol && dual_inf / osc <= o.dual_inf_tol && cviol_unsc <= o.constr_viol_tol &&
                      compl_inf / osc <= o.compl_inf_tol;
    const bool acc = E0 <= o.acceptable_tol && cviol_unsc <= 1e-2 && compl_inf / osc <= 1e-2;
    if (!retry) acceptable_cnt = acc ? acceptable_cnt + 1 : 0;
    bool to_main = false, to_p1 = false;
    if (phase1) {
      if (vl1 <= 0.0) to_main = true;
      else if (conv || acceptable_cnt >= o.acceptable_iter) {
        if (viol_max > o.infeas_tol) { status = ST_provenInfeasible; return 2; }
        to_main = true;
      }
    } else {
      if (conv) { status = ST_solvedOptimal; objv = val[0]; with_duals = true; return 2; }
      if (acceptable_cnt >= o.acceptable_iter) { status = ST_solvedOptimalTol; objv = val[0]; with_duals = true; return 2; }
    }
    if (!to_main && iters >= o.max_iter) { status = ST_iterationLimit; objv = val[0]; with_duals = true; return 2; }
    if (!to_main && !phase1 && !retry) {
      th_hist[iters_main & 3] = vl1;
      const double vold = th_hist[(iters_main + 1) & 3];
      bool stall = iters_main >= 3 && vold > 0.0 && vl1 > 0.9 * vold && a_last0 < 0.3 && a_last1 < 0.3;
      iters_main++;
      if (phase1_entries > 0) stall = false;
      if (th > o.infeas_ctol && (ymax > o.infeas_ytol || stall) && vl1 > 0.0) {
        if (phase1_entries >= o.max_phase1) { status = ST_iterationLimit; objv = val[0]; with_duals = true; return 2; }
        to_p1 = true;
      }
    }
    if (to_main) {
      __syncthreads();
      phase1 = false; iters_main = 0;
      if (tid == 0) { flg[n] |= F_FIXED; x[n] = 0.0; }
      __syncthreads();
      if (tid == 0) {
        int f = 0;
        for (int j = 0; j < N; ++j)
          if (flg[j] & F_FIXED) fixlist[f++] = j;
        s_nfix = f;
      }
      init_slacks_mults();
      theta_init = -1; mu_max = -1; acceptable_cnt = 0;
      ++iters;
      return 1;
    }
    if (to_p1) {
      enter_phase1();
      if (tid == 0) {
        int f = 0;
        for (int j = 0; j < N; ++j)
          if (flg[j] & F_FIXED) fixlist[f++] = j;
        s_nfix = f;
      }
      __syncthreads();
      ++iters;
      return 1;
    }
    return 0;
  };

  for (int step = 0;; ++step) {
    // =================================================================== A: local pre-phase
    for (;;) {
      if (state == CS_NEED_CHILD) state = setup_child() ? CS_INIT : CS_IDLE;
      if (state != CS_ITER) break;
      const int act = top_of_iteration();
      if (act == 0) break;
      if (act == 2) { write_results(); state = CS_NEED_CHILD; }
    }
    // publish the Hessian weights (zero unless iterating) and, when initialising, the point to evaluate
    for (int qi = tid; qi < nqpad; qi += IPM_NT) {
      double w = 0.0;
      if (state == CS_ITER && qi < nqf) {
        const int k = qfn[qi];
        w = (k == 0) ? it_of : 0.0;
        const int ru = M.fn_row_up[k], rl = M.fn_row_lo[k];
        if (ru >= 0) w += v[ru] * fsc[ru];
        if (rl >= 0) w += v[rl] * fsc[rl];
      }
      __stcg(CB.xw + (size_t)cta * nqpad + qi, w);
    }
    if (state != CS_ITER)
      for (int j = tid; j < ldd; j += IPM_NT) __stcg(CB.xd + (size_t)cta * ldd + j, (state == CS_INIT && j < n) ? x[j] : 0.0);
    if (tid == 0) atomicAdd(&CB.idle_cnt[step & 1], state == CS_IDLE ? 1 : 0);
    grid.sync();   // ------------------------------------------------------------------------------ barrier 1
    if (*((volatile int *)&CB.idle_cnt[step & 1]) == G) break;
    if (cta == 0 && tid == 0) CB.idle_cnt[(step + 1) & 1] = 0;

    // =================================================================== B: cooperative W pass (DMMA)
    if (coop_on) {
      // weights of my group -> shared memory: wsm[16][nqpad]
      double *wsm = stage;
      for (int e = tid; e < COOP_RB * nqpad; e += IPM_NT) {
        const int b = my_g * COOP_RB + e / nqpad;
        wsm[e] = (b < G) ? __ldcg(CB.xw + (size_t)b * nqpad + e % nqpad) : 0.0;
      }
      __syncthreads();
      const int NT8 = (P + 7) >> 3;
      const int t0 = (int)(((long long)NT8 * my_s) / NS), t1 = (int)(((long long)NT8 * (my_s + 1)) / NS);
      const int lr = lane >> 2, lc = lane & 3;
      for (int t = t0 + wid; t < t1; t += IPM_NW) {
        const int e = min(8 * t + lr, P - 1);
        const bool ev = 8 * t + lr < P;
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        for (int ks = 0; ks < nqpad; ks += 16) {   // 4 k-steps in flight
          double bq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int qi = ks + 4 * u + lc;
            const int qc = min(qi, nqf - 1);
            const double qv = ld_stream(M.Qp + (size_t)M.qidx[qfn[qc]] * P + e);
            bq[u] = (qi < nqf && ev && ks + 4 * u < nqpad) ? qv : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (ks + 4 * u < nqpad) {
              const int kk = ks + 4 * u + lc;
              dmma(c00, c01, wsm[lr * nqpad + kk], bq[u]);
              dmma(c10, c11, wsm[(8 + lr) * nqpad + kk], bq[u]);
            }
          }
        }
        // C[row = child lr (+8)][col = element 2*lc, 2*lc+1]
        const int ecol = 8 * t + 2 * lc;
        const int b0i = my_g * COOP_RB + lr, b1i = b0i + 8;
        if (b0i < G) {
          if (ecol < P) __stcg(CB.Wbuf + (size_t)b0i * P + ecol, c00);
          if (ecol + 1 < P) __stcg(CB.Wbuf + (size_t)b0i * P + ecol + 1, c01);
        }
        if (b1i < G) {
          if (ecol < P) __stcg(CB.Wbuf + (size_t)b1i * P + ecol, c10);
          if (ecol + 1 < P) __stcg(CB.Wbuf + (size_t)b1i * P + ecol + 1, c11);
        }
      }
    }
    grid.sync();   // ------------------------------------------------------------------------------ barrier 2

    // =================================================================== C: factorisation, direction (per child)
    if (state == CS_ITER) {
      const double of = it_of;
      // H += W  (rows < n), fixed rows/cols -> identity, diagonal += Sigma_x + delta_w
      const double *Wb = CB.Wbuf + (size_t)cta * P;
      for (int e = tid; e < P; e += IPM_NT) H[e] += __ldcg(Wb + e);
      __syncthreads();
      const int nfix = s_nfix;
      for (int e = tid; e < nfix * N; e += IPM_NT) {
        const int j = fixlist[e / N], t = e % N;
        if (t < j) H[tri(j) + t] = 0.0;
        else if (t > j) H[tri(t) + j] = 0.0;
      }
      __syncthreads();
      for (int j = tid; j < N; j += IPM_NT) {
        if (flg[j] & F_FIXED) H[tri(j) + j] = 1.0;
        else H[tri(j) + j] += Sx[j] + delta_w;
      }
      ++nfact;
      retry = false;
      if (!chol_dmma(H, N, N + 2, stage, lds, &s_fail)) {
        // inertia correction: retry the whole step with a larger delta_w (first 1e-4, x100 then x8, /3 between iterations)
        if (delta_w == 0.0) delta_w = (delta_w_last == 0.0) ? 1e-4 : fmax(1e-20, delta_w_last / 3.0);
        else delta_w *= (delta_w_last == 0.0 ? 100.0 : 8.0);
        retry = true;
        if (delta_w > 1e20 || ++fact_attempts > 45) { status = ST_computationError; objv = val[0]; with_duals = true; }
        for (int j = tid; j < N; j += IPM_NT) dx[j] = 0.0;
        __syncthreads();
      } else {
        if (delta_w > 0) delta_w_last = delta_w;
        delta_w = 0.0; fact_attempts = 0;
        backsolve_inv(H, N);
        dots_grad(M, grad, d0, d1, gd0, gd1);
        for (int r = tid; r < R; r += IPM_NT) {
          const int fn = M.row_fn[r];
          jd0[r] = fsc[r] * gd0[fn] - d0[n];
          jd1[r] = fsc[r] * gd1[fn] - d1[n];
        }
        __syncthreads();
        auto ftb = [&](double mu, double tau, double &ap, double &ad) {
          double mn[2] = {-1.0, -1.0};
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
          for (int r = tid; r < R; r += IPM_NT) {
            const double dsr = -rc[r] - (jd0[r] + mu * jd1[r]);
            const double dvr = mu / s[r] - v[r] - Ss[r] * dsr;
            if (dsr < 0) mn[0] = fmax(mn[0], -(-tau * s[r] / dsr));
            if (dvr < 0) mn[1] = fmax(mn[1], -(-tau * v[r] / dvr));
          }
          block_reduce<0, 2>(mn, red);
          ap = fmin(1.0, -mn[0]); ad = fmin(1.0, -mn[1]);
        };
        double mu;
        {
          double ap, ad;
          ftb(0.0, 1.0, ap, ad);
          double t[1] = {0.0};
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
          for (int r = tid; r < R; r += IPM_NT) {
            const double dsr = -rc[r] - jd0[r];
            const double dvr = -v[r] - Ss[r] * dsr;
            t[0] += (s[r] + ap * dsr) * (v[r] + ad * dvr);
          }
          block_reduce<1, 0>(t, red);
          const double mu_aff = it_nz > 0 ? t[0] / it_nz : 0.0;
          double sigma = it_avg_compl > 0 ? pow(mu_aff / it_avg_compl, 3.0) : o.sigma_min;
          sigma = fmin(fmax(sigma, o.sigma_min), o.sigma_max);
          mu = fmin(fmax(sigma * it_avg_compl, o.mu_min), mu_max);
          mu = fmax(mu, fmin(it_avg_compl, o.mu_safeguard * fmax(it_cviol, it_dual_inf / it_s_d)));
        }
        const double tau = fmax(o.tau_min, 1.0 - mu);
        double ap_max, ad;
        ftb(mu, tau, ap_max, ad);
        it_mu = mu; it_ap_max = ap_max; it_ad = ad;
        for (int j = tid; j < N; j += IPM_NT) dx[j] = (flg[j] & F_FIXED) ? 0.0 : d0[j] + mu * d1[j];
        for (int r = tid; r < R; r += IPM_NT) {
          ds[r] = -rc[r] - (jd0[r] + mu * jd1[r]);
          dv[r] = mu / s[r] - v[r] - Ss[r] * ds[r];
        }
        for (int k = tid; k <= m; k += IPM_NT) gd0[k] = gd0[k] + mu * gd1[k];
        __syncthreads();
      }
      (void)of;
      for (int j = tid; j < ldd; j += IPM_NT) __stcg(CB.xd + (size_t)cta * ldd + j, j < n ? dx[j] : 0.0);
    }
    grid.sync();   // ------------------------------------------------------------------------------ barrier 3

    // =================================================================== D: cooperative U pass (DMMA)
    if (coop_on) {
      double *Dsm = stage;   // [16][ldd]
      for (int e = tid; e < COOP_RB * ldd; e += IPM_NT) {
        const int b = my_g * COOP_RB + e / ldd;
        Dsm[e] = (b < G) ? __ldcg(CB.xd + (size_t)b * ldd + e % ldd) : 0.0;
      }
      __syncthreads();
      const int tpb = (n + 7) >> 3, ntile = nqf * tpb;
      const int t0 = (int)(((long long)ntile * my_s) / NS), t1 = (int)(((long long)ntile * (my_s + 1)) / NS);
      const int lr = lane >> 2, lc = lane & 3;
      const int n4 = (n + 3) & ~3;
      for (int t = t0 + wid; t < t1; t += IPM_NW) {
        const int qi = t / tpb, r0 = (t - qi * tpb) << 3;
        const int k = qfn[qi];
        const int r = r0 + lr;
        const double *Qrow = M.Qf + ((size_t)M.qidx[k] * n + min(r, n - 1)) * n;
        const bool rv = r < n;
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        for (int ks = 0; ks < n4; ks += 32) {   // 8 k-steps in flight
          double aq[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int col = ks + 4 * u + lc;
            const double qv = ld_stream(Qrow + min(col, n - 1));
            aq[u] = (rv && col < n) ? qv : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int col = ks + 4 * u + lc;
            if (ks + 4 * u < n4) {
              dmma(c00, c01, aq[u], Dsm[lr * ldd + col]);
              dmma(c10, c11, aq[u], Dsm[(8 + lr) * ldd + col]);
            }
          }
        }
        // C[row = matrix row lr][col = child 2*lc, 2*lc+1 (+8)]
        if (rv) {
          const size_t off = (size_t)(m + 1) * n + (size_t)k * n + r;   // uq of the child's workspace
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int b = my_g * COOP_RB + (h >> 1) * 8 + 2 * lc + (h & 1);
            const double cv = (h == 0) ? c00 : (h == 1) ? c01 : (h == 2) ? c10 : c11;
            if (b < G) __stcg(A.ws + (size_t)b * A.ws_stride + off, cv);
          }
        }
      }
    }
    grid.sync();   // ------------------------------------------------------------------------------ barrier 4

    // =================================================================== E: initialisation tail / line search + update
    if (state == CS_INIT) {
      if (finish_init()) { write_results(); state = CS_NEED_CHILD; }
      else state = CS_ITER;
    } else if (state == CS_ITER) {
      if (status == ST_computationError) { write_results(); state = CS_NEED_CHILD; continue; }
      if (retry) continue;   // the next step repeats this iteration with the larger delta_w
      const double mu = it_mu, ap_max = it_ap_max, ad = it_ad, tcur = it_tcur;
      // du_k = dx.u_k
      for (int qi = wid; qi < nqf; qi += IPM_NW) {
        const int k = qfn[qi];
        const double *uk = uq + (size_t)k * n;
        double part = 0.0;
        for (int j = lane; j < n; j += 32) part = fma(dx[j], __ldcg(uk + j), part);
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) part += __shfl_xor_sync(0xffffffffu, part, q);
        if (lane == 0) du[k] = part;
      }
      __syncthreads();
      const double dxt = phase1 ? dx[n] : 0.0;
      auto merit = [&](double a, double &theta, double &phi) {
        double t2[2] = {0.0, 0.0};
        const double t = tcur + a * dxt;
        for (int r = tid; r < R; r += IPM_NT) {
          const int fn = M.row_fn[r];
          const double hv = fsc[r] * (val[fn] + a * gd0[fn] + 0.5 * a * a * du[fn]) - t;
          const double sv = s[r] + a * ds[r];
          t2[0] += fabs(hv + sv - bb[r]);
          t2[1] -= mu * log(sv);
        }
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          const double xv = x[j] + a * dx[j];
          if (f & F_HASL) t2[1] -= mu * log(xv - Lr[j]);
          if (f & F_HASU) t2[1] -= mu * log(Ur[j] - xv);
        }
        block_reduce<2, 0>(t2, red);
        theta = t2[0];
        phi = t2[1] + (phase1 ? t : sf * (val[0] + a * gd0[0] + 0.5 * a * a * du[0]));
      };
      double theta0, phi0;
      merit(0.0, theta0, phi0);
      double dphi;
      {
        double t[1] = {0.0};
        for (int r = tid; r < R; r += IPM_NT) t[0] -= mu * ds[r] / s[r];
        for (int j = tid; j < N; j += IPM_NT) {
          const int f = flg[j];
          if (f & F_FIXED) continue;
          if (f & F_HASL) t[0] -= mu * dx[j] / (x[j] - Lr[j]);
          if (f & F_HASU) t[0] += mu * dx[j] / (Ur[j] - x[j]);
        }
        block_reduce<1, 0>(t, red);
        dphi = t[0] + (phase1 ? dxt : sf * gd0[0]);
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
        if (phase1 || phase1_entries >= o.max_phase1) {
          status = ST_iterationLimit; objv = val[0]; with_duals = true;
          write_results(); state = CS_NEED_CHILD;
          continue;
        }
        enter_phase1();
        if (tid == 0) {
          int f = 0;
          for (int j = 0; j < N; ++j)
            if (flg[j] & F_FIXED) fixlist[f++] = j;
          s_nfix = f;
        }
        __syncthreads();
        ++iters;
        continue;
      }
      a_last1 = a_last0; a_last0 = alpha;
      __syncthreads();
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
      for (int r = tid; r < R; r += IPM_NT) {
        const double sn = s[r] + alpha * ds[r];
        s[r] = sn;
        v[r] = fmin(fmax(v[r] + ad * dv[r], mu / (o.kappa_sigma * sn)), o.kappa_sigma * mu / sn);
      }
      for (int k = tid; k <= m; k += IPM_NT) val[k] += alpha * gd0[k] + 0.5 * alpha * alpha * du[k];
      for (int e = tid; e < nqf * n; e += IPM_NT) {
        const int qi = e / n, j = e - qi * n, k = qfn[qi];
        grad[(size_t)k * n + j] += alpha * __ldcg(uq + (size_t)k * n + j);
      }
      __syncthreads();
      ++iters;
    }
  }
}

}  // namespace bnlp
