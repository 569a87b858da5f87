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

namespace BNLP_NS {
namespace cg = cooperative_groups;

#if defined(UEXP) && UEXP == 2
#define UEXP_LOAD(p) (1.0 + 0.0 * (double)((size_t)(p) & 1))
#else
#define UEXP_LOAD(p) ld_stream_v(p)
#endif
__device__ unsigned long long g_dbg_cycles[4];
constexpr int COOP_RB = 32;   // children per cooperative group (four m8 tiles): every block element crosses L2->SM once per 32 children
constexpr int COOP_NP = COOP_RB / 8;

}  // namespace BNLP_NS
#include "bnlp_dense.cuh"
namespace BNLP_NS {

// ---- H[i,k] += sign * sum_c M[c][i] M[c][k], lo <= k <= i < hi, 8x8 tiles on the FP64 tensor pipe.
// M: K rows (K multiple of 4) of length ld in shared memory, zero outside [lo,hi) within the touched tile range.
// A warp owns strips of up to four column tiles of one tile row (A fragment loaded once, four independent DMMA chains).
// Only column tiles bk in [bkmin, bkmax] are touched; the work is shared by warps w0 .. w0+nw-1.
__device__ __forceinline__ void syrk_dmma(double *H, int lo, int hi, const double *M, int ld, int K, double sign,
                                          int bkmin = 0, int bkmax = 1 << 28, int w0 = 0, int nw = IPM_NW) {
  const int w = (threadIdx.x >> 5) - w0, lane = threadIdx.x & 31;
  if (w < 0 || w >= nw) return;
  const int b0 = lo >> 3, b1 = (hi + 7) >> 3;
  const int kb0 = max(b0, bkmin), kb1 = min(b1 - 1, bkmax);
  if (kb1 < kb0) return;
  const int lr = lane >> 2, lc = lane & 3;
  // walk over the valid strips only: tile row bi >= kb0 owns ((min(bi,kb1) - kb0) >> 2) + 1 strips of <= 4 tiles
  int bi = kb0, qq = w;
  for (;;) {
    int nqr = ((min(bi, kb1) - kb0) >> 2) + 1;
    while (qq >= nqr) {
      qq -= nqr;
      if (++bi >= b1) return;
      nqr = ((min(bi, kb1) - kb0) >> 2) + 1;
    }
    const int kq = kb0 + 4 * qq;
    const int klast = min(bi, kb1);
    const int nt = min(4, klast - kq + 1);
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    const double *Ma = M + lc * ld + (bi << 3) + lr, *Mb = M + lc * ld + (kq << 3) + lr;
    if (nt == 4) {
      for (int c = 0; c < K; c += 4) {
        const double av = Ma[c * ld];
        const double b0v = Mb[c * ld], b1v = Mb[c * ld + 8], b2v = Mb[c * ld + 16], b3v = Mb[c * ld + 24];
        dmma(acc[0][0], acc[0][1], av, b0v);
        dmma(acc[1][0], acc[1][1], av, b1v);
        dmma(acc[2][0], acc[2][1], av, b2v);
        dmma(acc[3][0], acc[3][1], av, b3v);
      }
    } else {
      for (int c = 0; c < K; c += 4) {
        const double av = Ma[c * ld];
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (t < nt) dmma(acc[t][0], acc[t][1], av, Mb[c * ld + 8 * t]);
      }
    }
    const int i = (bi << 3) + lr;
    double *Hi = H + tri(i) + (kq << 3) + 2 * lc;
    // interior strip: every touched element is inside [lo,hi) and strictly below the diagonal -> no per-element tests
    const bool interior = (bi << 3) >= lo && (bi << 3) + 7 < hi && (kq << 3) >= lo && kq + nt - 1 < bi;
    if (interior) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < nt) {
          Hi[8 * t] = fma(sign, acc[t][0], Hi[8 * t]);
          Hi[8 * t + 1] = fma(sign, acc[t][1], Hi[8 * t + 1]);
        }
      }
    } else if (i >= lo && i < hi) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < nt) {
          const int k = ((kq + t) << 3) + 2 * lc;
          if (k >= lo && k <= i) Hi[8 * t] = fma(sign, acc[t][0], Hi[8 * t]);
          if (k + 1 >= lo && k + 1 <= i) Hi[8 * t + 1] = fma(sign, acc[t][1], Hi[8 * t + 1]);
        }
      }
    }
    qq += nw;
  }
}

// factor the 8x8 diagonal block at j0 and overwrite it with the INVERSE of its Cholesky factor; *fail = 1 on a
// non-positive pivot.  Done by ONE thread entirely in registers (fully unrolled): the 8 dependent rsqrt steps are
// the critical path of the whole factorisation, and the in-register version is ~3x shorter than a warp-shuffle one
// (f64 shuffles cost 36 cycles each on this part; measured in scripts/micro/dmma_bench.cu).
static __device__ __noinline__ void chol_diag_block(double *H, int j0, int nb, int *fail) {
  if ((threadIdx.x & 31) != 0) return;
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) a[i][k] = (i < nb) ? H[tri(j0 + i) + j0 + k] : ((i == k) ? 1.0 : 0.0);
  bool ok = true;
  double rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // right-looking: the updates of one step are independent FMAs
    const double d = a[j][j];
    if (!(d > 0.0) || !isfinite(d)) ok = false;
    rs[j] = rsqrt(d);
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= rs[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
  }
  // in-place inverse of the lower-triangular factor, column by column: X_jj = 1/L_jj = rs_j,
  // X_ij = -rs_i * sum_{k=j}^{i-1} L_ik X_kj
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double x[8];
    x[j] = rs[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) acc = fma(a[i][k], x[k], acc);
      x[i] = -rs[i] * acc;
    }
#pragma unroll
    for (int i = j; i < 8; ++i) a[i][j] = x[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k)
      if (i < nb) H[tri(j0 + i) + j0 + k] = a[i][k];
  *fail = ok ? 0 : 1;
}

// ---- blocked Cholesky, 8-column panels; rows [N, Nrows) carry right-hand sides (forward substitution for free).
// After the call the 8x8 DIAGONAL BLOCKS of H hold the INVERSES of the diagonal blocks of L (lower triangular), the
// off-diagonal part holds L.  Per panel: TRSM as an 8x8 DMMA with the inverse block (one warp per 8 rows), the next
// panel's block column is updated first, then warp 0 factors the next diagonal block WHILE warps 1..15 finish the
// trailing update (look-ahead).  Returns false (block-uniform) on a non-positive pivot.
__device__ __forceinline__ bool chol_dmma(double *H, int N, int Nrows, double *stage, int ld, int *fail, long long *dbg) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  long long t0 = dbg ? clock64() : 0;
#define CH_MARK(i) do { if (dbg) { __syncthreads(); long long t_ = clock64(); dbg[i] += t_ - t0; t0 = t_; } } while (0)
  __syncthreads();
  if (w == 0) chol_diag_block(H, 0, min(8, N), fail);
  __syncthreads();
  if (*fail) return false;
  CH_MARK(0);
  for (int j0 = 0; j0 < N; j0 += 8) {
    const int nb = min(8, N - j0);
    const int lo = j0 + nb;
    if (lo >= Nrows) break;
    const int lo8 = lo & ~7, hi8 = (Nrows + 7) & ~7;
    // TRSM: X_panel = A_panel * Xinv'   (C[8 rows x 8] = A[8 x 8] * B[8 x 8], B[k][c] = Xinv[c][k])
    for (int i0 = lo8 + 8 * w; i0 < hi8; i0 += 8 * IPM_NW) {
      const int i = i0 + lr;
      const bool iv = i >= lo && i < Nrows;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int kk = 4 * ks + lc;
        const double av = (iv && kk < nb) ? H[tri(i) + j0 + kk] : 0.0;
        const double bv = (lr < nb && kk <= lr) ? H[tri(j0 + lr) + j0 + kk] : 0.0;
        dmma(c0, c1, av, bv);
      }
      const int cc = 2 * lc;
      if (iv) {
        if (cc < nb) H[tri(i) + j0 + cc] = c0;
        if (cc + 1 < nb) H[tri(i) + j0 + cc + 1] = c1;
      }
      stage[cc * ld + i] = (iv && cc < nb) ? c0 : 0.0;
      stage[(cc + 1) * ld + i] = (iv && cc + 1 < nb) ? c1 : 0.0;
    }
    __syncthreads();
    CH_MARK(1);
    if (lo >= N) break;                     // only right-hand-side rows were left
    const int bnext = lo >> 3;              // lo is a multiple of 8 here
    syrk_dmma(H, lo, Nrows, stage, ld, 8, -1.0, bnext, bnext);          // the next panel's block column first
    __syncthreads();
    {
      const long long tq0 = dbg ? clock64() : 0;
      if (w == 0) chol_diag_block(H, lo, min(8, N - lo), fail);           // look-ahead: next diagonal block ...
      else syrk_dmma(H, lo, Nrows, stage, ld, 8, -1.0, bnext + 1, 1 << 28, 1, IPM_NW - 1);   // ... under the rest of the update
      if (dbg && (threadIdx.x == 0 || threadIdx.x == 32)) atomicAdd((unsigned long long *)&g_dbg_cycles[threadIdx.x >> 5], (unsigned long long)(clock64() - tq0));
    }
    __syncthreads();
    if (*fail) return false;
    CH_MARK(2);
  }
  __syncthreads();
  return true;
}

// ---- backward substitution L' x = w for the two right-hand sides in rows N and N+1, 8-row blocks, diagonal blocks
// given as inverses.  Per block: 16 threads form x_blk = X' w_blk for both right-hand sides, then every remaining
// entry w[k], k < i0, is updated by its own thread (8 FMAs).
__device__ __forceinline__ void backsolve_inv(double *H, int N, double *xs /* 16 doubles of scratch */) {
  const int tid = threadIdx.x;
  const int nblk = (N + 7) >> 3;
  double *b0 = H + tri(N), *b1 = H + tri(N + 1);
  const int half = tid / (IPM_NT / 2), kk = tid % (IPM_NT / 2);   // one half of the CTA per right-hand side (N <= IPM_NT / 2)
  double *bh = half ? b1 : b0;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int i0 = blk << 3, nb = min(8, N - i0);
    if (tid < 16) {
      const int rhs = tid >> 3, c = tid & 7;
      const double *bb_ = rhs ? b1 : b0;
      double xc = 0.0;
      if (c < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nb && k >= c) xc = fma(H[tri(i0 + k) + i0 + c], bb_[i0 + k], xc);
      }
      xs[tid] = xc;
    }
    __syncthreads();
    if (kk < i0) {
      double v = bh[kk];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < nb) v = fma(-H[tri(i0 + c) + kk], xs[half * 8 + c], v);
      bh[kk] = v;
    } else if (kk < i0 + nb) {
      bh[kk] = xs[half * 8 + (kk - i0)];
    }
    __syncthreads();
  }
}

// =====================================================================================================================
// Back-substitution with 32 x 32 INVERSE diagonal blocks.
//
// backsolve_inv walks the factor in 8-row blocks: 26 blocks at n = 200, two CTA barriers and two dependent 8-FMA chains
// per block, 23-25 k cycles for 0.08 MFLOP.  Here the diagonal 32 x 32 blocks of L that are made of four FULL tiles are
// first replaced by their inverses (one warp per block, all blocks at once: Inv_pq = -X_p sum_{r=q}^{p-1} L_pr Inv_rq
// over 8 x 8 tiles on DMMA, in place column by column — L_pq is last needed when Inv_pq is formed), then a 32-row block is
// ONE dot-product stage (x_B = Inv_B' w_B) and one update stage.  Ragged tail tiles keep the 8-row procedure.

// C (8 x 8, accumulator layout) += sgn * A(tile at rows ia.., cols ka..) * B(tile at rows ib.., cols kb..); lower-triangular
// operand tiles (diagonal tiles of X / Inv) are masked by a_lo / b_lo (entries above the diagonal do not exist in the packed
// storage).
__device__ __forceinline__ void tile_mm(const double *H, int ia, int ka, bool a_lo, int ib, int kb, bool b_lo, double sgn,
                                        double &c0, double &c1) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int kk = 4 * ks + lc;
    // A[m = lr][k = kk] = tile_a(lr, kk);  B[k = kk][n = lr] = tile_b(kk, lr)
    const double av = (!a_lo || kk <= lr) ? H[tri(ia + lr) + ka + kk] : 0.0;
    const double bv = (!b_lo || lr <= kk) ? H[tri(ib + kk) + kb + lr] : 0.0;
    dmma(c0, c1, sgn * av, bv);
  }
}
__device__ __forceinline__ void tile_store(double *H, int i0, int k0, double c0, double c1) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  double *p = H + tri(i0 + lr) + k0 + 2 * lc;
  p[0] = c0; p[1] = c1;
}

// in-place inverse of the 32 x 32 lower-triangular diagonal block at i0 (four full tiles; diagonal tiles already hold the
// inverses X_p of the 8 x 8 diagonal blocks).  One whole warp.
__device__ __forceinline__ void invert_block32(double *H, int i0) {
  for (int q = 0; q < 3; ++q) {
    for (int p = q + 1; p < 4; ++p) {
      // S = sum_{r=q}^{p-1} L_pr Inv_rq   (Inv_qq = X_q is lower triangular; Inv_rq, r > q, full)
      double s0 = 0.0, s1 = 0.0;
      for (int r = q; r < p; ++r) tile_mm(H, i0 + 8 * p, i0 + 8 * r, false, i0 + 8 * r, i0 + 8 * q, r == q, 1.0, s0, s1);
      __syncwarp();
      tile_store(H, i0 + 8 * p, i0 + 8 * q, s0, s1);        // L_pq is dead: its place takes S, then Inv_pq
      __syncwarp();
      double c0 = 0.0, c1 = 0.0;
      tile_mm(H, i0 + 8 * p, i0 + 8 * p, true, i0 + 8 * p, i0 + 8 * q, false, -1.0, c0, c1);   // -X_p S
      __syncwarp();
      tile_store(H, i0 + 8 * p, i0 + 8 * q, c0, c1);
      __syncwarp();
    }
  }
}

// L' x = w for the two right-hand sides in rows N and N + 1 (same contract as backsolve_inv; xs: 64 doubles of scratch).
__device__ __forceinline__ void backsolve32(double *H, int N, double *xs) {
  const int tid = threadIdx.x, w = tid >> 5;
  const int nfull = N >> 5;                      // 32-row blocks made of full tiles
  for (int b = w; b < nfull; b += IPM_NW) invert_block32(H, 32 * b);
  __syncthreads();
  double *b0 = H + tri(N), *b1 = H + tri(N + 1);
  const int half = tid / (IPM_NT / 2), kk = tid % (IPM_NT / 2);
  double *bh = half ? b1 : b0;
  // ragged tail: 8-row blocks above the last full 32-block
  const int nblk = (N + 7) >> 3;
  for (int blk = nblk - 1; blk >= 4 * nfull; --blk) {
    const int i0 = blk << 3, nb = min(8, N - i0);
    if (tid < 16) {
      const int rhs = tid >> 3, c = tid & 7;
      const double *bb_ = rhs ? b1 : b0;
      double xc = 0.0;
      if (c < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nb && k >= c) xc = fma(H[tri(i0 + k) + i0 + c], bb_[i0 + k], xc);
      }
      xs[tid] = xc;
    }
    __syncthreads();
    if (kk < i0) {
      double v = bh[kk];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < nb) v = fma(-H[tri(i0 + c) + kk], xs[half * 8 + c], v);
      bh[kk] = v;
    } else if (kk < i0 + nb) {
      bh[kk] = xs[half * 8 + (kk - i0)];
    }
    __syncthreads();
  }
  // full blocks: LPD lanes per dot product (64 dot products: 32 entries x 2 right-hand sides)
  // (four lanes per dot product in BOTH builds of the kernel: the 256- and the 512-thread variant must sum in the same order,
  // results are bit-identical whichever one a launch picks)
  constexpr int LPD = 4;
  const int dot = tid / LPD, part = tid % LPD;   // dot = rhs * 32 + c
  const int drhs = (dot >> 5) & 1, dc = dot & 31;
  for (int b = nfull - 1; b >= 0; --b) {
    const int i0 = 32 * b;
    if (tid < 64 * LPD) {
      const double *wv = drhs ? b1 : b0;
      double a = 0.0;
#pragma unroll
      for (int u = 0; u < 32 / LPD; ++u) {
        const int k = part + LPD * u;            // x[c] = sum_{k >= c} Inv[k][c] w[k]
        if (k >= dc) a = fma(H[tri(i0 + k) + i0 + dc], wv[i0 + k], a);
      }
#pragma unroll
      for (int o = LPD / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (part == 0) xs[dot] = a;
    }
    __syncthreads();
    if (kk < i0) {
      double v0 = bh[kk], v1 = 0.0, v2 = 0.0, v3 = 0.0;
      const double *xv = xs + half * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        v0 = fma(-H[tri(i0 + c) + kk], xv[c], v0);
        v1 = fma(-H[tri(i0 + c + 1) + kk], xv[c + 1], v1);
        v2 = fma(-H[tri(i0 + c + 2) + kk], xv[c + 2], v2);
        v3 = fma(-H[tri(i0 + c + 3) + kk], xv[c + 3], v3);
      }
      bh[kk] = (v0 + v1) + (v2 + v3);
    } else if (kk < i0 + 32) {
      bh[kk] = xs[half * 32 + (kk - i0)];
    }
    __syncthreads();
  }
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) global -> shared memory, completion on an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// orders earlier generic-proxy accesses (ordinary loads / stores) before later async-proxy ones (bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct CoopBufs {
  double *xw;       // G x nqpad : Hessian weights published by every CTA
  double *Wbuf;     // G x P     : W_b produced by the cooperative W pass
  double *xd;       // G x ldd   : direction (or start point at initialisation) published by every CTA
  int *idle_cnt;    // [2]
  int nqpad, ldd;
};

// CS_SKIP: an iterating child that sits one step out (after a phase switch everything is re-evaluated at the top of the
// NEXT step instead of a second time in this one); CS_FINISH: converged at the top of this step — results are written
// and the next child is fetched in phase C, where this CTA has nothing else to do.  Both keep the pre-barrier phase of
// every CTA the same length (a straggler there makes all 148 CTAs wait).
struct CoopScalars {
  double objv_, sf_, badw_, theta_init_, mu_max_, th_hist_[4];
  double it_nz_, it_avg_compl_, it_cviol_, it_dual_inf_, it_s_d_, it_viol_max_, it_of_, it_tcur_, it_mu_, it_ap_max_, it_ad_;
};

enum { CS_NEED_CHILD = 0, CS_INIT = 1, CS_ITER = 2, CS_IDLE = 3, CS_SKIP = 4, CS_FINISH = 5 };

template <bool TIMERS>
__global__ void __launch_bounds__(IPM_NT, 512 / IPM_NT) ipm_coop_kernel(DevModel M, IpmArgs A, CoopBufs CB) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double smem[];
  const int n = M.n, m = M.m, N = M.N, R = M.R, npad = M.npad;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const IpmOpt &o = A.o;
  const int G = gridDim.x, cta = blockIdx.x;
  // ---- carve shared memory (same budget as ipm_kernel; the staging rows get a bank-friendly stride)
  const int lds = ((N + 2 + 7) & ~7) + 4;   // >= roundup8(N+2), stride = 4 (mod 8) doubles: conflict-free DMMA fragments
  // Only a few base pointers are kept; the vectors behind them are macros (base + k * stride), so their addresses are
  // recomputed where used instead of holding ~30 registers for the whole kernel (same register-cap reason as CoopScalars).
  const int Haug = tri(N + 2);
  double *const H = smem;
  double *const stage = H + ((Haug + 1) & ~1);
  double *const x = stage + (IPM_KC + 1) * lds;      // 7 vectors of npad: x zL zU Lr Ur dx Sx
  const int Rp = (R + 1) & ~1, Kp = (m + 2) & ~1;
  double *const s = x + 7 * npad;                    // 10 vectors of Rp: s v rc ds dv Ss bb fsc jd0 jd1
  double *const val = s + 10 * Rp;                   // val gd0 gd1 du (Kp each), wq (2 Kp)
  double *const red = val + 6 * Kp;
#define zL (x + npad)
#define zU (x + 2 * npad)
#define Lr (x + 3 * npad)
#define Ur (x + 4 * npad)
#define dx (x + 5 * npad)
#define Sx (x + 6 * npad)
#define v (s + Rp)
#define rc (s + 2 * Rp)
#define ds (s + 3 * Rp)
#define dv (s + 4 * Rp)
#define Ss (s + 5 * Rp)
#define bb (s + 6 * Rp)
#define fsc (s + 7 * Rp)
#define jd0 (s + 8 * Rp)
#define jd1 (s + 9 * Rp)
#define gd0 (val + Kp)
#define gd1 (val + 2 * Kp)
#define du (val + 3 * Kp)
#define wq (val + 4 * Kp)
  int *const flg = reinterpret_cast<int *>(red + IPM_NW * 12);
#define fixlist (flg + ((N + 1) & ~1))
#define qblk (flg + 2 * ((N + 1) & ~1))
#define qfn (flg + 2 * ((N + 1) & ~1) + Kp)
#define drow (flg + 2 * ((N + 1) & ~1) + 2 * Kp)
  __shared__ int s_child, s_nfix, s_fail, s_nqf;
  // The scaled Jacobian rows of a J' Sigma J chunk are staged by BULK COPIES (TMA): thread 0 issues one cp.async.bulk per
  // raw gradient row (n doubles from this CTA's workspace in L2 into the staging area), every warp then scales its row in
  // place.  The copies of the first chunk are issued at the very top of the iteration and land under the residual loops; the
  // register path (eight loads in flight per lane) cost 4.3 k cycles per chunk, all of it exposed.  Needs 16-byte aligned
  // rows: even n (otherwise the register path below is used).
  __shared__ __align__(8) unsigned long long s_mbar;
  // MEASURED (round 2, C3, 4096 children): 41.8 k relaxations/s with the bulk copies against 43.0 k with the register path —
  // staging itself drops from 7.8 M to 5.8 M cycles per launch and CTA, but issuing sixteen copies per chunk from one
  // thread after the SYRK barrier, the extra in-place pass and the proxy fences behind the gradient updates cost more than
  // that.  The path is therefore OFF unless BNLP_COOP_TMA=1 (IpmArgs::coop_tma); profiles/r2_coop_tma_staging.txt.
  const bool use_tma = A.coop_tma && (n & 1) == 0 && ((A.ws_stride & 1) == 0);
  unsigned tma_parity = 0;
  if (tid == 0) mbar_init(&s_mbar, 1);
#define grad (A.ws + (size_t)cta * A.ws_stride)
#define uq (A.ws + (size_t)cta * A.ws_stride + (size_t)(m + 1) * n)
#define d0 (H + tri(N))
#define d1 (H + tri(N + 1))
  // the three per-row scalars used by the J' products reuse ds/dv/jd0 (free at that point of the iteration)
#define pv ds
#define p0 dv
#define p1 jd0
#define psq jd1

  if (tid == 0) {
    int c = 0;
    for (int k = 0; k <= m; ++k)
      if (M.qidx[k] >= 0) { qfn[c] = k; qblk[c] = M.qidx[k]; ++c; }
    s_nqf = c;
  }
#pragma unroll 1
  for (int k = tid; k <= m; k += IPM_NT) du[k] = 0.0;
  // rows that go through the staged DMMA updates of J' Sigma J (the sparse linear rows are scattered entry by entry)
  const int Rd = M.Rd;
#pragma unroll 1
  for (int r = tid; r < Rd; r += IPM_NT) drow[r] = M.dense_rows[r];
#pragma unroll 1
  for (int r = tid; r < M.Rs; r += IPM_NT) drow[Rd + r] = M.sp_rows[r];
  __syncthreads();
  const int nqf = s_nqf;
  const int nqpad = CB.nqpad, ldd = CB.ldd;
  // cooperative decomposition: CTA = (child group g, slice s)
  // Only the first Gc CTAs take children (Gc = the batch rounded up to whole groups, at most G); the cooperative passes
  // of those NG groups are sliced over ALL G CTAs, so a small batch (one B-BB node: 64 children) still has every SM
  // streaming the shared blocks.
  const int Gc = min(G, (A.B + COOP_RB - 1) / COOP_RB * COOP_RB);
  const int NG = (Gc + COOP_RB - 1) / COOP_RB;
  const int NS = max(1, G / NG);
  const int my_g = cta / NS, my_s = cta % NS;
  const bool coop_on = my_g < NG;
  const int P = M.P;

  // ---- per-child state
  int state = CS_NEED_CHILD;
  long long seg_t0 = 0;   // (timers build) start and start state of the per-child segment between barriers 4 and 1
  int seg_state = 0;
  (void)seg_t0; (void)seg_state;
  int child = -1;
#define xl (A.XL + (size_t)child * n)
#define xu (A.XU + (size_t)child * n)
#define x0 (A.X0 + (size_t)child * A.x0_stride)
  int status = 0, iters = 0;
  // CTA-uniform doubles that are only ever written with a value every thread has computed identically (after a block
  // reduction) live in shared memory instead of 40 registers per thread: the kernel runs at the 128-register cap of a
  // 512-thread CTA and what does not fit is spilled to local memory, which misses the 28 KB L1 left beside 220 KB of
  // shared memory.  (Counters and anything updated read-modify-write stay in registers: no ordering between threads.)
  __shared__ CoopScalars cs_;
#define objv cs_.objv_
#define sf cs_.sf_
#define badw cs_.badw_
#define theta_init cs_.theta_init_
#define mu_cap cs_.mu_max_
#define th_hist cs_.th_hist_
#define it_nz cs_.it_nz_
#define it_avg_compl cs_.it_avg_compl_
#define it_cviol cs_.it_cviol_
#define it_dual_inf cs_.it_dual_inf_
#define it_s_d cs_.it_s_d_
#define it_viol_max cs_.it_viol_max_
#define it_of cs_.it_of_
#define it_tcur cs_.it_tcur_
#define it_mu cs_.it_mu_
#define it_ap_max cs_.it_ap_max_
#define it_ad cs_.it_ad_
  objv = 0.0; sf = 1.0; badw = 0.0;
  bool with_duals = false, inconsistent = false, all_fixed = false, phase1 = false, retry = false;
  int phase1_entries = 0, iters_main = 0, acceptable_cnt = 0, fact_attempts = 0;
  double delta_w_last = 0.0, delta_w = 0.0;
  double a_last0 = 1.0, a_last1 = 1.0;
  theta_init = -1.0; mu_cap = -1.0;
  th_hist[0] = th_hist[1] = th_hist[2] = th_hist[3] = 0.0;
  long long nfact = 0;
  // timer build: counters live in shared memory and only thread 0 touches them, so the instrumented kernel keeps the
  // register allocation (and spills) of the production one.  [0..9] phases, [10..13] barrier waits, [16..27] sub-phases
  __shared__ long long s_tm[TIMERS ? 44 : 1];
  if constexpr (TIMERS) { if (threadIdx.x < 44) s_tm[threadIdx.x] = 0; }
  long long dbg_t0 = 0;
#define D2_START() do { if constexpr (TIMERS) { if (tid == 0) dbg_t0 = clock64(); } } while (0)
#define D2_MARK(i) do { if constexpr (TIMERS) { __syncthreads(); if (tid == 0) { long long t_ = clock64(); s_tm[16 + (i)] += t_ - dbg_t0; dbg_t0 = t_; } } } while (0)
#undef PH_MARK
#define PH_MARK(idx) do { if constexpr (TIMERS) { if (tid == 0) { const long long t_ = clock64(); s_tm[idx] += t_ - ph_t0; ph_t0 = t_; } } } while (0)
  // values carried from the top of an iteration to its later phases
  it_nz = 0; it_avg_compl = 0; it_cviol = 0; it_dual_inf = 0; it_s_d = 1; it_viol_max = 0; it_of = 0; it_tcur = 0;
  it_mu = 0; it_ap_max = 0; it_ad = 0;

  // warm (first initialisation only): the caller's multipliers [z_L | z_U | lambda] (unscaled) start the solve, each pushed
  // to at least warm_push, slacks pushed by warm_push; a phase switch restarts from unit multipliers
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
    __syncthreads();
  };
  auto enter_phase1 = [&]() {
    __syncthreads();
    phase1 = true; phase1_entries++;
    const double vm = fmax(0.0, it_viol_max), T = fmax(1.0, vm);
    if (tid == 0) { flg[n] &= ~F_FIXED; Lr[n] = -T; x[n] = 1.1 * vm + 0.1 * T; }
    __syncthreads();
    init_slacks_mults(false);
    theta_init = -1; mu_cap = -1; acceptable_cnt = 0;
  };

  // ---- fetch + set up the next child; returns false when the queue is empty
  auto setup_child = [&]() -> bool {
    __syncthreads();
    if (tid == 0) s_child = atomicAdd(A.queue, 1);
    __syncthreads();
    child = s_child;
    if (child >= A.B) return false;
    status = 0; iters = 0; objv = 0.0; sf = 1.0; with_duals = false; phase1 = false; retry = false;
    phase1_entries = 0; iters_main = 0; acceptable_cnt = 0; fact_attempts = 0;
    theta_init = -1.0; mu_cap = -1.0; delta_w_last = 0.0; delta_w = 0.0;
    th_hist[0] = th_hist[1] = th_hist[2] = th_hist[3] = 0.0; a_last0 = a_last1 = 1.0; nfact = 0;
    double acc0 = 0.0, acc1 = 0.0, bw = 0.0;
#pragma unroll 1
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
          const double bp = A.DUAL0 ? o.warm_push : o.bound_push, bf = A.DUAL0 ? o.warm_push : o.bound_frac;   // warm start: 1e-3
          if ((f & F_HASL) && (f & F_HASU)) {
            const double pl = fmin(bp * fmax(1.0, fabs(lr)), bf * (ur - lr));
            const double pu = fmin(bp * fmax(1.0, fabs(ur)), bf * (ur - lr));
            xx = fmin(fmax(xx, lr + pl), ur - pu);
          } else if (f & F_HASL) xx = fmax(xx, lr + bp * fmax(1.0, fabs(lr)));
          else if (f & F_HASU) xx = fmin(xx, ur - bp * fmax(1.0, fabs(ur)));
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
      if (use_tma) fence_proxy_async();     // these rows are read by bulk copies (async proxy) at the top of the next iteration
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
      if (lane == 0) val[k] = M.c[k] + acc;
    }
    __syncthreads();
    if (inconsistent || all_fixed) {
      double t = 0.0, infe = 0.0;
#pragma unroll 1
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
#pragma unroll 1
    for (int r = tid; r < R; r += IPM_NT) {
      const int fn = M.row_fn[r];
      const double sc = wq[fn], sg = M.row_sgn[r];
      const double b = sc * M.row_bnd[r];
      fsc[r] = sc * sg;
      bb[r] = b + o.bound_relax * fmax(1.0, fabs(b));
    }
    __syncthreads();
    init_slacks_mults(true);
    return false;
  };

  auto write_results = [&]() {
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
      objv = t[0];
    }
    if (tid == 0) {
      A.OBJ[child] = objv; A.STATUS[child] = status; A.ITERS[child] = iters;
      if (A.stats) { atomicAdd((unsigned long long *)&A.stats[0], (unsigned long long)iters);
                     atomicAdd((unsigned long long *)&A.stats[1], (unsigned long long)nfact); }
    }
    if (A.X) for (int j = tid; j < n; j += IPM_NT)
      A.X[(size_t)child * n + j] = all_fixed ? xl[j] : fmin(fmax(x[j], fmin(xl[j], xu[j])), fmax(xl[j], xu[j]));
    if (A.G) {
#pragma unroll 1
      for (int i = tid; i < m; i += IPM_NT) A.G[(size_t)child * m + i] = val[i + 1];
    }
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
    auto tma_issue = [&](int r0) {     // raw gradient rows r0 .. r0 + kc of the chunk -> staging rows (thread 0)
      if (tid == 0) {
        const int kc = min(IPM_KC, Rd - r0);
        // (write-after-read on the staging area needs no proxy fence: the generic-proxy reads of the previous chunk were
        // complete at the CTA barrier this thread has just left)
        mbar_expect_tx(&s_mbar, (unsigned)(kc * n * 8));
        for (int c = 0; c < kc; ++c) tma_load_1d(stage + c * lds, grad + (size_t)M.row_fn[drow[r0 + c]] * n, (unsigned)(n * 8), &s_mbar);
      }
    };
    if (use_tma && Rd > 0) tma_issue(0);
    D2_START();
    double th_p = 0, sumv_p = 0, cs_p = 0, vl1_p = 0, ci_p = 0, cv_p = 0, cvu_p = 0, vmax_p = 0, viol_p = -1e300;
#pragma unroll 1
    for (int r = tid; r < R; r += IPM_NT) {
      const double hraw = fsc[r] * val[M.row_fn[r]];
      const double res = hraw - tcur + s[r] - bb[r];
      rc[r] = res;
      const double ss = v[r] / s[r], sq = sqrt(ss);
      Ss[r] = ss;
      pv[r] = v[r] / sq; p0[r] = sq * res; p1[r] = 1.0 / (sq * s[r]); psq[r] = sq;
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
    D2_MARK(0);
    // H := 0 (rows < N), then J' Sigma_s J by staged rank-16 DMMA updates; the same staged rows give J'v, J'(Ss rc), J'(1/s)
    for (int e = tid; e < tri(N); e += IPM_NT) H[e] = 0.0;
    double jv = 0.0, a0 = 0.0, a1 = 0.0;
    __syncthreads();
    D2_MARK(1);
    // warp c stages row c of a chunk; its raw gradient row is prefetched into registers one chunk ahead
    double gpre[8];
    auto prefetch_row = [&](int r) {
      const double *g = grad + (size_t)M.row_fn[r] * n;
#pragma unroll
      for (int u = 0; u < 8; ++u) gpre[u] = g[min(lane + 32 * u, n - 1)];
    };
    for (int r0 = 0; r0 < Rd; r0 += IPM_KC) {
      const int kc = min(IPM_KC, Rd - r0), kc4 = (kc + 3) & ~3;
      if (use_tma) {
        mbar_wait(&s_mbar, tma_parity);
        tma_parity ^= 1u;
        for (int c = wid; c < kc4; c += IPM_NW) {   // scale / mask the landed row in place
          if (c < kc) {
            const int r = drow[r0 + c];
            const double sq = psq[r], f = sq * fsc[r];
            for (int j = lane; j < lds; j += 32) {
              double vv = 0.0;
              if (j < N && !(flg[j] & F_FIXED)) vv = (j < n) ? f * stage[c * lds + j] : -sq;
              stage[c * lds + j] = vv;
            }
          } else {
            for (int j = lane; j < lds; j += 32) stage[c * lds + j] = 0.0;
          }
        }
      } else
      for (int c = wid; c < kc4; c += IPM_NW) {   // one staged row per warp (two with the 8-warp build)
        const int r = drow[min(r0 + c, Rd - 1)];
        if (c < kc) prefetch_row(r);       // eight loads in flight per lane; a prefetch one chunk ahead only spilled
        if (c < kc) {
          const double sq = psq[r], f = sq * fsc[r];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = lane + 32 * u;
            if (j < lds) {
              double vv = 0.0;
              if (j < N && !(flg[j] & F_FIXED)) vv = (j < n) ? f * gpre[u] : -sq;
              stage[c * lds + j] = vv;
            }
          }
        } else {
          for (int j = lane; j < lds; j += 32) stage[c * lds + j] = 0.0;
        }
      }
      __syncthreads();
      D2_MARK(8);
      if (tid < N) {
        for (int c = 0; c < kc; ++c) {
          const double sv = stage[c * lds + tid];
          const int r = drow[r0 + c];
          jv = fma(sv, pv[r], jv);
          a0 = fma(sv, p0[r], a0);
          a1 = fma(sv, p1[r], a1);
        }
      }
      D2_MARK(9);
      if (kc4 == IPM_KC) syrk_dmma(H, 0, N, stage, lds, IPM_KC, 1.0);   // compile-time rank: the k loop unrolls
      else syrk_dmma(H, 0, N, stage, lds, kc4, 1.0);
      __syncthreads();
      if (use_tma && r0 + IPM_KC < Rd) tma_issue(r0 + IPM_KC);   // (single staging buffer: the next chunk can only start now)
      D2_MARK(10);
    }
    // ---- sparse linear rows (DevModel::spc_* / sph_*, layouts explained in build_ipm_view): no staged rank-1 updates.
    // Same values as a staged row would give (sv = (sqrt(Ss_r) fsc_r) a_rk, -sqrt(Ss_r) in the phase-1 elastic's column),
    // summed in a fixed order; every load of the lists is independent of the arithmetic, several are kept in flight.
    if (M.Rs > 0) {
      const bool el_free = !(flg[n] & F_FIXED);       // phase 1: the elastic's row / column of H is live
      double a2 = 0.0;
      if (tid < n && !(flg[tid] & F_FIXED)) {
        const int S2 = M.spc_slots, ldc = M.spc_ld;
        for (int s0 = 0; s0 < S2; s0 += 4) {
          int rr[4];
          double av[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sl = min(s0 + u, S2 - 1);
            rr[u] = __ldg(M.spc_row + sl * ldc + tid);
            av[u] = __ldg(M.spc_val + sl * ldc + tid);
            if (s0 + u >= S2) rr[u] = -1;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (rr[u] >= 0) {
              const int r = rr[u];
              const double sv = (psq[r] * fsc[r]) * av[u];
              jv = fma(sv, pv[r], jv);
              a0 = fma(sv, p0[r], a0);
              a1 = fma(sv, p1[r], a1);
              a2 = fma(Ss[r] * fsc[r], av[u], a2);
            }
          }
        }
        if (el_free) H[tri(n) + tid] -= a2;
      } else if (tid == n && el_free) {
        double hnn = 0.0;
        for (int q = 0; q < M.Rs; ++q) {
          const int r = drow[Rd + q];
          const double sv = -psq[r];
          jv = fma(sv, pv[r], jv);
          a0 = fma(sv, p0[r], a0);
          a1 = fma(sv, p1[r], a1);
          hnn += Ss[r];
        }
        H[tri(n) + n] += hnn;
      }
      const int SH = M.sph_slots;
      const double2 *rec = reinterpret_cast<const double2 *>(M.sph_rec);
      for (int vt = tid; vt < 512; vt += IPM_NT) {
        double acc = 0.0;
        for (int s0 = 0; s0 < SH; s0 += 8) {
          double2 q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) q[u] = __ldg(rec + (size_t)min(s0 + u, SH - 1) * 512 + vt);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (s0 + u < SH) {
              const long long key = __double_as_longlong(q[u].x);
              const int r = (int)(key >> 32) & 0xfffffff;
              acc = fma(Ss[r] * fsc[r] * fsc[r], q[u].y, acc);
              if ((key >> 62) & 1) { H[(int)(key & 0x7fffffff)] += acc; acc = 0.0; }
            }
          }
        }
      }
    }
    D2_MARK(2);
    // sweep over variables: Lagrangian gradient, right-hand sides (into the augmented rows), Sigma_x
    double di_p = 0, sumz_p = 0, nz_p = 0;
#pragma unroll 1
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
    D2_MARK(3);
    const double th = all[0], sum_v = all[1], vl1 = all[3];
    const double nz = all[5] + (double)R, sum_z = all[4] + sum_v;
    const double compl_inf = all[6], cviol = all[7], cviol_unsc = all[8], ymax = all[9], viol_max = all[10], dual_inf = all[11];
    const double avg_compl = nz > 0 ? all[2] / nz : 0.0;
    if (theta_init < 0) theta_init = th;
    const double s_d = fmax(o.s_max, (sum_v + sum_z) / fmax(1.0, (double)R + nz)) / o.s_max;
    const double s_c = fmax(o.s_max, sum_z / fmax(1.0, nz)) / o.s_max;
    const double E0 = fmax(fmax(dual_inf / s_d, cviol), compl_inf / s_c);
    if (mu_cap < 0) mu_cap = fmin(o.mu_max, o.mu_max_fact * fmax(avg_compl, 1e-300));
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
        // (an only "acceptable" phase-1 optimum proves infeasibility only beyond what its own error E0 could explain)
        if (viol_max > (conv ? o.infeas_tol : fmax(o.infeas_tol, 10.0 * E0))) { status = ST_provenInfeasible; return 2; }
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
      init_slacks_mults(false);
      theta_init = -1; mu_cap = -1; acceptable_cnt = 0;
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

#define PH_MARKB(b) do { if constexpr (TIMERS) { if (tid == 0) { const long long t_ = clock64(); s_tm[1] += t_ - ph_t0; if (state == CS_ITER) s_tm[10 + (b)] += t_ - ph_t0; ph_t0 = t_; } } } while (0)
  long long ph_t0 = TIMERS ? clock64() : 0;
  for (int step = 0;; ++step) {
    // =================================================================== A: local pre-phase
    if (state == CS_SKIP) state = CS_ITER;
    if (state == CS_ITER) {
      const int act = top_of_iteration();
      if (act == 1) state = CS_SKIP;
      else if (act == 2) state = CS_FINISH;
    }
    if (state == CS_NEED_CHILD) state = (cta < Gc && setup_child()) ? CS_INIT : CS_IDLE;
    // publish the Hessian weights (zero unless iterating) and, when initialising, the point to evaluate
#pragma unroll 1
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
#pragma unroll 1
      for (int j = tid; j < ldd; j += IPM_NT) __stcg(CB.xd + (size_t)cta * ldd + j, (state == CS_INIT && j < n) ? x[j] : 0.0);
    if (tid == 0) atomicAdd(&CB.idle_cnt[step & 1], state == CS_IDLE ? 1 : 0);
    if constexpr (TIMERS) {      // length of the per-child segment between barrier 4 and barrier 1, by the state it started in
      if (tid == 0 && step > 0) {
        const long long dt = clock64() - seg_t0;
        const int bkt = seg_state == CS_INIT ? 0 : seg_state == CS_ITER ? 1 : 2;
        s_tm[31 + 3 * bkt] += dt; s_tm[32 + 3 * bkt] += 1; if (dt > s_tm[33 + 3 * bkt]) s_tm[33 + 3 * bkt] = dt;
        if (bkt == 1 && dt > 120000) s_tm[40] += 1;
        if (bkt == 1 && dt > 160000) s_tm[41] += 1;
      }
    }
    PH_MARK(0);
    grid.sync();   // ------------------------------------------------------------------------------ barrier 1
    PH_MARKB(0);
    if (*((volatile int *)&CB.idle_cnt[step & 1]) == G) break;
    if (cta == 0 && tid == 0) CB.idle_cnt[(step + 1) & 1] = 0;

    // =================================================================== B: cooperative W pass (DMMA)
    // Every global request is a whole 128-byte line: lane (lc, lr) loads the double2 {e0+2lr, e0+2lr+1} of packed block
    // 4u+lc; the even elements feed one m8n8k4 tile, the odd ones a second tile, so a lane ends up owning four
    // consecutive elements of W for one child (two 16-byte stores).
    if (coop_on) {
      double *wsm = stage;   // wsm[16][ldw], ldw = 4 (mod 8): conflict-free A fragments
      const int ldw = nqpad + ((nqpad & 7) == 4 ? 0 : 4);
      for (int e = tid; e < COOP_RB * nqpad; e += IPM_NT) {
        const int row = e / nqpad, col = e % nqpad;
        const int b = my_g * COOP_RB + row;
        wsm[row * ldw + col] = (b < Gc) ? __ldcg(CB.xw + (size_t)b * nqpad + col) : 0.0;
      }
      __syncthreads();
      const int NT16 = (P + 15) >> 4;
      const int t0 = (int)(((long long)NT16 * my_s) / NS), t1 = (int)(((long long)NT16 * (my_s + 1)) / NS);
      const int lr = lane >> 2, lc = lane & 3;
      for (int t = t0 + wid; t < t1; t += IPM_NW) {
        const int e0 = 16 * t;                       // ldp is a multiple of 16: the line always exists (zero padded)
        double cE[COOP_NP][2], cO[COOP_NP][2];
#pragma unroll
        for (int p = 0; p < COOP_NP; ++p) cE[p][0] = cE[p][1] = cO[p][0] = cO[p][1] = 0.0;
        for (int ks = 0; ks < nqpad; ks += 32) {     // 8 k-steps (32 blocks) per round
          double2 qv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int qi = ks + 4 * u + lc;
            const double2 *Qb = reinterpret_cast<const double2 *>(M.Qp + (size_t)qblk[min(qi, nqf - 1)] * M.ldp + e0) + lr;
            qv[u] = __ldg(Qb);
            if (qi >= nqf) { qv[u].x = 0.0; qv[u].y = 0.0; }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (ks + 4 * u < nqpad) {
              const int kk = ks + 4 * u + lc;
#pragma unroll
              for (int p = 0; p < COOP_NP; ++p) {
                const double wv = wsm[(8 * p + lr) * ldw + kk];
                dmma(cE[p][0], cE[p][1], wv, qv[u].x);
                dmma(cO[p][0], cO[p][1], wv, qv[u].y);
              }
            }
          }
        }
        // even tile: C[child 8p + lr][n = 2lc, 2lc+1] <-> elements e0 + 4lc, e0 + 4lc + 2; odd tile: e0 + 4lc + 1, + 3
        const int ec = e0 + 4 * lc;
#pragma unroll
        for (int p = 0; p < COOP_NP; ++p) {
          const int bi = my_g * COOP_RB + 8 * p + lr;
          if (bi < Gc && ec < P) {   // Wbuf rows have stride ldp too
            double2 *dst = reinterpret_cast<double2 *>(CB.Wbuf + (size_t)bi * M.ldp + ec);
            __stcg(dst, make_double2(cE[p][0], cO[p][0]));
            __stcg(dst + 1, make_double2(cE[p][1], cO[p][1]));
          }
        }
      }
    }
    PH_MARK(2);
    grid.sync();   // ------------------------------------------------------------------------------ barrier 2
    PH_MARKB(1);

    // =================================================================== C: factorisation, direction (per child)
    if (state == CS_FINISH) {
      write_results();
      state = setup_child() ? CS_INIT : CS_IDLE;   // (only CTAs below Gc ever get here)
#pragma unroll 1
      for (int j = tid; j < ldd; j += IPM_NT) __stcg(CB.xd + (size_t)cta * ldd + j, (state == CS_INIT && j < n) ? x[j] : 0.0);
    } else if (state == CS_ITER) {
      const double of = it_of;
      // H += W  (rows < n), fixed rows/cols -> identity, diagonal += Sigma_x + delta_w
      const double *Wb = CB.Wbuf + (size_t)cta * M.ldp;
      {   // 16-byte requests (Wbuf rows are 128-byte aligned, H is 16-byte aligned in shared memory), four in flight per thread
        const double2 *W2 = reinterpret_cast<const double2 *>(Wb);
        double2 *H2 = reinterpret_cast<double2 *>(H);
        const int P2 = P >> 1;
        int e = tid;
        for (; e + 3 * IPM_NT < P2; e += 4 * IPM_NT) {
          const double2 w0 = __ldcg(W2 + e), w1 = __ldcg(W2 + e + IPM_NT), w2 = __ldcg(W2 + e + 2 * IPM_NT), w3 = __ldcg(W2 + e + 3 * IPM_NT);
          double2 h0 = H2[e], h1 = H2[e + IPM_NT], h2 = H2[e + 2 * IPM_NT], h3 = H2[e + 3 * IPM_NT];
          h0.x += w0.x; h0.y += w0.y; h1.x += w1.x; h1.y += w1.y; h2.x += w2.x; h2.y += w2.y; h3.x += w3.x; h3.y += w3.y;
          H2[e] = h0; H2[e + IPM_NT] = h1; H2[e + 2 * IPM_NT] = h2; H2[e + 3 * IPM_NT] = h3;
        }
        for (; e < P2; e += IPM_NT) { const double2 w = __ldcg(W2 + e); double2 h = H2[e]; h.x += w.x; h.y += w.y; H2[e] = h; }
        if ((P & 1) && tid == 0) H[P - 1] += __ldcg(Wb + P - 1);
      }
      __syncthreads();
      const int nfix = s_nfix;
      for (int e = tid; e < nfix * N; e += IPM_NT) {
        const int j = fixlist[e / N], t = e % N;
        if (t < j) H[tri(j) + t] = 0.0;
        else if (t > j) H[tri(t) + j] = 0.0;
      }
      __syncthreads();
#pragma unroll 1
      for (int j = tid; j < N; j += IPM_NT) {
        if (flg[j] & F_FIXED) H[tri(j) + j] = 1.0;
        else H[tri(j) + j] += Sx[j] + delta_w;
      }
      ++nfact;
      retry = false;
      PH_MARK(8);
#ifdef BNLP_COOP_CHOL3
      if (!chol3(H, N, N + 2, stage, lds, &s_fail, TIMERS ? &s_tm[27] : nullptr)) {
#else
      if (!chol4(H, N, N + 2, stage, lds, &s_fail, TIMERS ? &s_tm[27] : nullptr)) {
#endif
        // inertia correction: retry the whole step with a larger delta_w (first 1e-4, x100 then x8, /3 between iterations)
        if (delta_w == 0.0) delta_w = (delta_w_last == 0.0) ? 1e-4 : fmax(1e-20, delta_w_last / 3.0);
        else delta_w *= (delta_w_last == 0.0 ? 100.0 : 8.0);
        retry = true;
        if (delta_w > 1e20 || ++fact_attempts > 45) { status = ST_computationError; objv = val[0]; with_duals = true; }
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) dx[j] = 0.0;
        __syncthreads();
      } else {
        if (delta_w > 0) delta_w_last = delta_w;
        delta_w = 0.0; fact_attempts = 0;
        PH_MARK(3);
        D2_START();
#ifdef BNLP_OLD_BACKSOLVE
        backsolve_inv(H, N, red);
#else
        backsolve32(H, N, red);
#endif
        D2_MARK(4);
        // (requesting two whole rows per warp up front — sixteen loads in flight per lane — was measured: the phase itself gets
        // shorter, the kernel 4 % slower; the register arrays cost more elsewhere than the exposed L2 round trips do here)
        dots_grad(M, grad, d0, d1, gd0, gd1);
        D2_MARK(5);
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          const int fn = M.row_fn[r];
          jd0[r] = fsc[r] * gd0[fn] - d0[n];
          jd1[r] = fsc[r] * gd1[fn] - d1[n];
        }
        __syncthreads();
        auto ftb = [&](double mu, double tau, double &ap, double &ad) {
          double mn[2] = {-1.0, -1.0};
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
          // (rows go to the threads at the top of the CTA, variables to those at the bottom: the chain of FP64 divisions a
          // thread runs is the critical path of this pass; a max-reduction does not care who holds what)
#pragma unroll 1
          for (int r = IPM_NT - 1 - tid; r < R; r += IPM_NT) {
            const double dsr = -rc[r] - (jd0[r] + mu * jd1[r]);
            const double dvr = mu / s[r] - v[r] - Ss[r] * dsr;
            if (dsr < 0) mn[0] = fmax(mn[0], -(-tau * s[r] / dsr));
            if (dvr < 0) mn[1] = fmax(mn[1], -(-tau * v[r] / dvr));
          }
          block_reduce<0, 2>(mn, red);
          ap = fmin(1.0, -mn[0]); ad = fmin(1.0, -mn[1]);
        };
        // the fraction-to-the-boundary pass is instantiated ONCE and run twice (affine step: mu = 0, tau = 1; then the real
        // step): every inlined copy of these loops is instructions the SM has to fetch again each iteration
        double mu = 0.0, tau = 1.0, ap_max = 1.0, ad = 1.0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          double ap;
          ftb(mu, tau, ap, ad);
          if (pass == 1) { ap_max = ap; break; }
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
          const double mu_aff = it_nz > 0 ? t[0] / it_nz : 0.0;
          double sigma = it_avg_compl > 0 ? pow(mu_aff / it_avg_compl, 3.0) : o.sigma_min;
          sigma = fmin(fmax(sigma, o.sigma_min), o.sigma_max);
          mu = fmin(fmax(sigma * it_avg_compl, o.mu_min), mu_cap);
          mu = fmax(mu, fmin(it_avg_compl, o.mu_safeguard * fmax(it_cviol, it_dual_inf / it_s_d)));
          tau = fmax(o.tau_min, 1.0 - mu);
        }
        it_mu = mu; it_ap_max = ap_max; it_ad = ad;
        D2_MARK(6);
#pragma unroll 1
        for (int j = tid; j < N; j += IPM_NT) dx[j] = (flg[j] & F_FIXED) ? 0.0 : d0[j] + mu * d1[j];
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          ds[r] = -rc[r] - (jd0[r] + mu * jd1[r]);
          dv[r] = mu / s[r] - v[r] - Ss[r] * ds[r];
        }
#pragma unroll 1
        for (int k = tid; k <= m; k += IPM_NT) gd0[k] = gd0[k] + mu * gd1[k];
        __syncthreads();
        D2_MARK(7);
      }
      (void)of;
#pragma unroll 1
      for (int j = tid; j < ldd; j += IPM_NT) __stcg(CB.xd + (size_t)cta * ldd + j, j < n ? dx[j] : 0.0);
    }
    PH_MARK(4);
    grid.sync();   // ------------------------------------------------------------------------------ barrier 3
    PH_MARKB(2);

    // =================================================================== D: cooperative U pass (DMMA)
    // Whole-line requests again: lane (lr, lc) loads the 32-byte sector lc of the 128-byte line {row r0+lr, columns
    // 16g .. 16g+15}; k-step s of that group contracts over the columns {16g + 4lc + s}, which is only a permutation of
    // the summation index, so the B fragments are read from Dsm at the same permuted columns.
    if (coop_on) {
      double *Dsm = H;   // [COOP_RB][ldd]: the factor in H (and the staging rows behind it) is dead until the next iteration
      for (int e = tid; e < COOP_RB * ldd; e += IPM_NT) {
        const int b = my_g * COOP_RB + e / ldd;
        Dsm[e] = (b < Gc) ? __ldcg(CB.xd + (size_t)b * ldd + e % ldd) : 0.0;
      }
      __syncthreads();
      const int tpb = (n + 7) >> 3, ntile = nqf * tpb;
      const int t0 = (int)(((long long)ntile * my_s) / NS), t1 = (int)(((long long)ntile * (my_s + 1)) / NS);
      const int lr = lane >> 2, lc = lane & 3;
      const int ngrp = (n + 15) >> 4;       // 16-column groups (Qf rows are zero padded to ldq = roundup16(n))
      for (int t = t0 + wid; t < t1; t += IPM_NW) {
        const int qi = t / tpb, r0 = (t - qi * tpb) << 3;
        const int k = qfn[qi];
        const int r = r0 + lr;
        const double2 *Qrow = reinterpret_cast<const double2 *>(M.Qf + ((size_t)qblk[qi] * n + min(r, n - 1)) * M.ldq) + 2 * lc;
        const bool rv = r < n;
        double cu[COOP_NP][2];
#pragma unroll
        for (int p = 0; p < COOP_NP; ++p) cu[p][0] = cu[p][1] = 0.0;
        // software pipeline over pairs of groups: 4 x 16-byte loads in flight ahead of the DMMAs that consume them
        double2 aq[2][4];
        auto loadgrp = [&](int g, double2 *dst) {
          const int g0 = min(g, ngrp - 1), g1 = min(g + 1, ngrp - 1);
          dst[0] = __ldg(Qrow + 8 * g0); dst[1] = __ldg(Qrow + 8 * g0 + 1);
          dst[2] = __ldg(Qrow + 8 * g1); dst[3] = __ldg(Qrow + 8 * g1 + 1);
        };
        loadgrp(0, aq[0]);
        for (int g = 0; g < ngrp; g += 4) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int gc = g + 2 * half;
            if (gc < ngrp) {
              if (gc + 2 < ngrp) loadgrp(gc + 2, aq[half ^ 1]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (gc + h < ngrp) {
                  const int cb = 16 * (gc + h) + 4 * lc;
                  const double av[4] = {aq[half][2 * h].x, aq[half][2 * h].y, aq[half][2 * h + 1].x, aq[half][2 * h + 1].y};
#pragma unroll
                  for (int s4 = 0; s4 < 4; ++s4) {
                    const double a = rv ? av[s4] : 0.0;
#pragma unroll
                    for (int p = 0; p < COOP_NP; ++p) dmma(cu[p][0], cu[p][1], a, Dsm[(8 * p + lr) * ldd + cb + s4]);
                  }
                }
              }
            }
          }
        }
        // C[row = matrix row lr][col = child 8p + 2lc, + 1]
        if (rv) {
          const size_t off = (size_t)(m + 1) * n + (size_t)k * n + r;   // uq of the child's workspace
#pragma unroll
          for (int p = 0; p < COOP_NP; ++p)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int b = my_g * COOP_RB + 8 * p + 2 * lc + hh;
              if (b < Gc) __stcg(A.ws + (size_t)b * A.ws_stride + off, cu[p][hh]);
            }
        }
      }
    }
    PH_MARK(5);
    grid.sync();   // ------------------------------------------------------------------------------ barrier 4
    PH_MARKB(3);
    if constexpr (TIMERS) { if (tid == 0) { seg_t0 = clock64(); seg_state = state; } }

    // =================================================================== E: initialisation tail / line search + update
    if (state == CS_INIT) {
      PH_MARK(9);
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
#pragma unroll 1
        for (int r = tid; r < R; r += IPM_NT) {
          const int fn = M.row_fn[r];
          const double hv = fsc[r] * (val[fn] + a * gd0[fn] + 0.5 * a * a * du[fn]) - t;
          const double sv = s[r] + a * ds[r];
          t2[0] += fabs(hv + sv - bb[r]);
          t2[1] -= mu * log(sv);
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
        phi = t2[1] + (phase1 ? t : sf * (val[0] + a * gd0[0] + 0.5 * a * a * du[0]));
      };
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
        dphi = t[0] + (phase1 ? dxt : sf * gd0[0]);
      }
      const double theta_min = o.theta_min_fact * fmax(1.0, theta_init);
      const double theta_max = o.theta_max_fact * fmax(1.0, theta_init);
      double amin = o.gamma_theta, theta0 = 0.0, phi0 = 0.0;
      double alpha = ap_max;
      bool accepted = false;
#pragma unroll 1
      for (int bt = -1; bt < 60; ++bt) {       // bt = -1: the merit function at the current point (ONE inlined copy of merit)
        double th_t, ph_t;
        merit(bt < 0 ? 0.0 : alpha, th_t, ph_t);
        if (bt < 0) {
          theta0 = th_t; phi0 = ph_t;
          if (dphi < 0) {
            amin = fmin(amin, o.gamma_phi * theta0 / (-dphi));
            if (theta0 <= theta_min) amin = fmin(amin, o.delta * pow(theta0, o.s_theta) / pow(-dphi, o.s_phi));
          }
          amin *= o.alpha_min_frac;
          continue;
        }
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
      PH_MARK(6);
      __syncthreads();
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
      for (int k = tid; k <= m; k += IPM_NT) val[k] += alpha * gd0[k] + 0.5 * alpha * alpha * du[k];
      for (int qi = wid; qi < nqf; qi += IPM_NW) {        // one row per warp, sixteen loads in flight per lane
        const int k = qfn[qi];
        double *gk = grad + (size_t)k * n;
        const double *uk = uq + (size_t)k * n;
        double gv[8], uv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int j = min(lane + 32 * u, n - 1); gv[u] = gk[j]; uv[u] = __ldcg(uk + j); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = lane + 32 * u;
          if (j < n) gk[j] = gv[u] + alpha * uv[u];
        }
      }
      if (use_tma) fence_proxy_async();     // (see finish_init)
      __syncthreads();
      ++iters;
      PH_MARK(7);
    }
  }
  if constexpr (TIMERS) {
    if (tid == 0 && (cta % 37) == 0)
      printf("[dbg cta %d] barrier waits while iterating: b1 %lld b2 %lld b3 %lld b4 %lld | top: rows %lld zeroH %lld vars %lld, chunk stage %lld J'v %lld syrk %lld | C: backsolve %lld dots %lld ftb/mu %lld dir %lld | chol3: diag0 %lld trsm %lld col8+diag %lld trailing+diag %lld (cycles, whole launch) || segment b4->b1 by start state: INIT avg %lld n %lld max %lld | ITER avg %lld n %lld max %lld (>120k: %lld, >160k: %lld) | other avg %lld n %lld max %lld\n",
             cta, s_tm[10], s_tm[11], s_tm[12], s_tm[13], s_tm[16], s_tm[17], s_tm[19], s_tm[24], s_tm[25], s_tm[26], s_tm[20], s_tm[21], s_tm[22], s_tm[23], s_tm[27], s_tm[28], s_tm[29], s_tm[30], s_tm[31] / (s_tm[32] ? s_tm[32] : 1), s_tm[32], s_tm[33], s_tm[34] / (s_tm[35] ? s_tm[35] : 1), s_tm[35], s_tm[36], s_tm[40], s_tm[41], s_tm[37] / (s_tm[38] ? s_tm[38] : 1), s_tm[38], s_tm[39]);
    if (A.phase_cycles && tid == 0)
      for (int q = 0; q < 10; ++q) atomicAdd((unsigned long long *)&A.phase_cycles[q], (unsigned long long)s_tm[q]);
  }
}
#undef zL
#undef zU
#undef Lr
#undef Ur
#undef dx
#undef Sx
#undef v
#undef rc
#undef ds
#undef dv
#undef Ss
#undef bb
#undef fsc
#undef jd0
#undef jd1
#undef gd0
#undef gd1
#undef du
#undef wq
#undef fixlist
#undef qblk
#undef qfn
#undef d0
#undef d1
#undef pv
#undef p0
#undef p1
#undef psq
#undef drow
#undef xl
#undef xu
#undef x0
#undef grad
#undef uq
#undef objv
#undef sf
#undef badw
#undef theta_init
#undef mu_cap
#undef th_hist
#undef it_nz
#undef it_avg_compl
#undef it_cviol
#undef it_dual_inf
#undef it_s_d
#undef it_viol_max
#undef it_of
#undef it_tcur
#undef it_mu
#undef it_ap_max
#undef it_ad


}  // namespace BNLP_NS
