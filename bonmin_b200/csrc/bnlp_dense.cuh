// Per-child dense kernels of the cooperative IPM kernel (included by bnlp_ipm_coop.cuh after tri() and dmma()):
// the n x n work of one relaxation on one SM, H packed lower-triangular (row-major, tri(i) + k) in shared memory.
//
//   syrk_rt      H += sign * M' M  as 16 x 32 register tiles (2 x 4 DMMA tiles per warp: six operand fragments feed eight
//                m8n8k4 DMMAs per k-step), rank 16 per pass so the read-modify-write of H (the shared-memory bound of the
//                factorisation) happens once per 16 columns
//   chol2        two-level blocked Cholesky: 16-column panels = two 8-column sub-blocks; the 8 x 8 diagonal block is
//                factored by one thread in registers (8 dependent rsqrt: the critical path), the panel rows are solved
//                one row per thread by forward substitution (no block inverse on the critical path), the trailing matrix
//                gets ONE rank-16 update per panel, the next diagonal block is factored under the rest of that update
//   invert_diag_blocks   after the factorisation: every 8 x 8 diagonal block of L -> its inverse (all blocks in parallel),
//                which is what the blocked back-substitution multiplies with
#pragma once

namespace BNLP_NS {

// ---- H[i,k] += sign * sum_{c<K} M[c][i] M[c][k],  lo <= k <= i < hi,  lo % 8 == 0, K % 4 == 0.
// M: K rows of length ld (ld = 4 mod 8 doubles: conflict-free fragments), zero for i in [hi, roundup8(hi)).
// Work unit (r, c) = tile rows {b0+2r, b0+2r+1} x tile columns {b0+4c .. b0+4c+3}, c <= r/2, numbered row by row; the
// calling warp does units u0, u0 + ustride, ... (warp-uniform arguments).
__device__ __forceinline__ void syrk_rt(double *H, int lo, int hi, const double *M, int ld, int K, double sign,
                                        int u0, int ustride) {
  const int lane = threadIdx.x & 31;
  const int b0 = lo >> 3, b1 = (hi + 7) >> 3;
  const int nr = (b1 - b0 + 1) >> 1;
  const int lr = lane >> 2, lc = lane & 3;
  int r = 0, c = u0;
  for (;;) {
    while (r < nr && c > (r >> 1)) { c -= (r >> 1) + 1; ++r; }
    if (r >= nr) return;
    const int bi0 = b0 + 2 * r, bk0 = b0 + 4 * c;
    const bool has1 = bi0 + 1 < b1;
    const int nt = min(4, min(bi0 + (has1 ? 1 : 0), b1 - 1) - bk0 + 1);   // column tiles of this unit that hold anything
    double acc[2][4][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[u][t][0] = acc[u][t][1] = 0.0;
    const double *Ma = M + lc * ld + (bi0 << 3) + lr, *Mb = M + lc * ld + (bk0 << 3) + lr;
    if (nt == 4 && has1) {
#pragma unroll 2
      for (int k = 0; k < K; k += 4) {
        const double a0 = Ma[k * ld], a1 = Ma[k * ld + 8];
        const double v0 = Mb[k * ld], v1 = Mb[k * ld + 8], v2 = Mb[k * ld + 16], v3 = Mb[k * ld + 24];
        dmma(acc[0][0][0], acc[0][0][1], a0, v0);
        dmma(acc[1][0][0], acc[1][0][1], a1, v0);
        dmma(acc[0][1][0], acc[0][1][1], a0, v1);
        dmma(acc[1][1][0], acc[1][1][1], a1, v1);
        dmma(acc[0][2][0], acc[0][2][1], a0, v2);
        dmma(acc[1][2][0], acc[1][2][1], a1, v2);
        dmma(acc[0][3][0], acc[0][3][1], a0, v3);
        dmma(acc[1][3][0], acc[1][3][1], a1, v3);
      }
    } else {
      for (int k = 0; k < K; k += 4) {
        const double a0 = Ma[k * ld], a1 = has1 ? Ma[k * ld + 8] : 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < nt) {
            const double v = Mb[k * ld + 8 * t];
            dmma(acc[0][t][0], acc[0][t][1], a0, v);
            dmma(acc[1][t][0], acc[1][t][1], a1, v);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int bi = bi0 + u;
      if (bi >= b1) break;
      const int i = (bi << 3) + lr;
      double *Hi = H + tri(i) + (bk0 << 3) + 2 * lc;
      const bool rows_in = (bi << 3) + 7 < hi;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int bk = bk0 + t;
        if (t >= nt || bk > bi) break;
        if (rows_in && bk < bi) {            // interior tile: no per-element tests
          Hi[8 * t] = fma(sign, acc[u][t][0], Hi[8 * t]);
          Hi[8 * t + 1] = fma(sign, acc[u][t][1], Hi[8 * t + 1]);
        } else if (i < hi) {
          const int k = (bk << 3) + 2 * lc;
          if (k <= i) Hi[8 * t] = fma(sign, acc[u][t][0], Hi[8 * t]);
          if (k + 1 <= i) Hi[8 * t + 1] = fma(sign, acc[u][t][1], Hi[8 * t + 1]);
        }
      }
    }
    c += ustride;
  }
}

// the factorisations call the rank-16 trailing update from several places (warp 0's look-ahead unit, the other warps,
// chol3's tail behind chol4): ONE out-of-line copy, so that every warp runs the same instructions (the inlined copies were
// 3 000 of the cooperative kernel's instructions, and instruction fetch is what its warps wait for most after barriers)
static __device__ __noinline__ void syrk_rt16_sub(double *H, int lo, int hi, const double *M, int ld, int u0, int ustride) {
  syrk_rt(H, lo, hi, M, ld, 16, -1.0, u0, ustride);
}

// ---- rank-8 update of ONE 8-column block jc: H[i, jc .. jc+7] -= sum_{c<8} M[c][i] M[c][jc+k] for the tile rows from
// row0 (jc % 8 == row0 % 8 == 0, row0 >= jc) up to hi, dealt to warps: the calling warp is number wi of nw.
__device__ __forceinline__ void syrk_col8(double *H, int row0, int hi, const double *M, int ld, int wi, int nw, int jc = -1) {
  if (jc < 0) jc = row0;
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int b1 = (hi + 7) >> 3;
  const double v0 = M[lc * ld + jc + lr], v1 = M[(4 + lc) * ld + jc + lr];
  for (int bi = (row0 >> 3) + wi; bi < b1; bi += nw) {
    const int i = (bi << 3) + lr;
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, M[lc * ld + i], v0);
    dmma(c0, c1, M[(4 + lc) * ld + i], v1);
    if (i < hi) {
      double *Hi = H + tri(i) + jc + 2 * lc;
      const int k = jc + 2 * lc;
      if (k <= i) Hi[0] -= c0;
      if (k + 1 <= i) Hi[1] -= c1;
    }
  }
}

// ---- 8 x 8 diagonal block at j0 (nb valid rows): Cholesky in the registers of ONE thread.  On exit the strictly lower
// part holds L and the diagonal holds 1 / L_cc.  *fail = 1 on a non-positive pivot.
static __device__ __noinline__ void chol_diag8(double *H, int j0, int nb, int *fail) {
  if ((threadIdx.x & 31) != 0) return;
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) a[i][k] = (i < nb) ? H[tri(j0 + i) + j0 + k] : ((i == k) ? 1.0 : 0.0);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = a[j][j];
    if (!(d > 0.0) || !isfinite(d)) ok = false;
    const double rs = rsqrt(d);
    a[j][j] = rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k)
      if (i < nb) H[tri(j0 + i) + j0 + k] = a[i][k];
  *fail = ok ? 0 : 1;
}

// ---- panel rows below an 8-column diagonal block, one row per thread: l = a * L_d^{-T} by forward substitution.
// Rows [lo, Nrows); the solved rows go back to H and, transposed, to S[c][i] (c < 8; zero for c >= nb and for the
// padding rows i in [Nrows, roundup8(Nrows))).
__device__ __forceinline__ void trsm_rows8(double *H, int j0, int nb, int lo, int Nrows, double *S, int ld) {
  const int i = lo + (int)threadIdx.x;
  const int hi8 = (Nrows + 7) & ~7;
  if (i >= hi8) return;
  double l[8];
  if (i < Nrows) {
    const double *Hi = H + tri(i) + j0;
#pragma unroll
    for (int c = 0; c < 8; ++c) l[c] = (c < nb) ? Hi[c] : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < nb) {
        const double *Lk = H + tri(j0 + k) + j0;   // column k of L_d lives in rows k+1.. : L_d[c][k] = H[tri(j0+c) + j0 + k]
        l[k] *= Lk[k];
#pragma unroll
        for (int c = k + 1; c < 8; ++c)
          if (c < nb) l[c] = fma(-l[k], H[tri(j0 + c) + j0 + k], l[c]);
      }
    }
    double *Ho = H + tri(i) + j0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < nb) Ho[c] = l[c];
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) l[c] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) S[c * ld + i] = l[c];
}

// ---- every 8 x 8 diagonal block (strictly lower L, reciprocal diagonal) -> inverse of its Cholesky factor.
// Thread 8b + c computes column c of block b: X_cc = 1/L_cc, X_ic = -(1/L_ii) sum_{k=c}^{i-1} L_ik X_kc.
__device__ __forceinline__ void invert_diag_blocks(double *H, int N) {
  const int nblk = (N + 7) >> 3;
  const int b = threadIdx.x >> 3, c = threadIdx.x & 7;
  const int j0 = b << 3, nb = min(8, N - j0);
  double a[8][8];
  const bool act = b < nblk && c < nb;
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k <= i; ++k) a[i][k] = (i < nb && k >= c) ? H[tri(j0 + i) + j0 + k] : 0.0;
  }
  __syncthreads();
  if (act) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < c) { x[i] = 0.0; continue; }
      if (i == c) { x[i] = a[i][i]; continue; }
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k)
        if (k >= c) s = fma(a[i][k], x[k], s);
      x[i] = -a[i][i] * s;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= c && i < nb) H[tri(j0 + i) + j0 + c] = x[i];
  }
  __syncthreads();
}

// ---- blocked Cholesky of H[0..N) with rows [N, Nrows) carried as right-hand sides (their forward substitution is the
// panel solve).  On exit: off-diagonal part = L, 8 x 8 diagonal blocks = inverses of the diagonal blocks of L.
// Returns false (block-uniform) on a non-positive pivot.  S: staging area of 16 rows x ld.
template <bool LOOKAHEAD>
__device__ __forceinline__ bool chol2(double *H, int N, int Nrows, double *S, int ld, int *fail, long long *dbg = nullptr) {
  const int w = threadIdx.x >> 5;
  long long t0 = dbg ? clock64() : 0;
#define C2_MARK(i) do { if (dbg && threadIdx.x == 0) { long long t_ = clock64(); dbg[i] += t_ - t0; t0 = t_; } } while (0)
  __syncthreads();
  if (w == 0) chol_diag8(H, 0, min(8, N), fail);
  __syncthreads();
  if (*fail) return false;
  C2_MARK(0);
  for (int j0 = 0; j0 < N; j0 += 16) {
    const int nb0 = min(8, N - j0), j1 = j0 + 8;
    trsm_rows8(H, j0, nb0, j0 + nb0, Nrows, S, ld);
    __syncthreads();
    C2_MARK(1);
    if (j1 >= N) break;
    const int nb1 = min(8, N - j1), lo = j1 + nb1;
    syrk_col8(H, j1, Nrows, S, ld, w, IPM_NW);
    __syncthreads();
    C2_MARK(2);
    if (w == 0) chol_diag8(H, j1, nb1, fail);
    __syncthreads();
    if (*fail) return false;
    C2_MARK(0);
    trsm_rows8(H, j1, nb1, lo, Nrows, S + 8 * ld, ld);
    __syncthreads();
    C2_MARK(1);
    if (lo >= N) break;
    // trailing rank-16 update; with look-ahead warp 0 does the unit that holds the next diagonal block, then factors
    // that block while warps 1.. finish the update
    if (LOOKAHEAD) {
      if (w == 0) {
        syrk_rt(H, lo, Nrows, S, ld, 16, -1.0, 0, 1 << 20);
        __syncwarp();
        chol_diag8(H, lo, min(8, N - lo), fail);
      } else {
        syrk_rt(H, lo, Nrows, S, ld, 16, -1.0, w, IPM_NW - 1);
      }
      __syncthreads();
      C2_MARK(3);
    } else {
      syrk_rt(H, lo, Nrows, S, ld, 16, -1.0, w, IPM_NW);
      __syncthreads();
      C2_MARK(3);
      if (w == 0) chol_diag8(H, lo, min(8, N - lo), fail);
      __syncthreads();
      C2_MARK(0);
    }
    if (*fail) return false;
  }
  invert_diag_blocks(H, N);
  C2_MARK(4);
  return true;
}


// ---- 8 x 8 diagonal block at j0: Cholesky AND inverse of the factor in the registers of one thread.  The forward
// substitution of the identity (X = L^{-1}, row by row) is written into the same unrolled step as the pivot so its
// independent FMAs issue under the rsqrt latency of the pivot chain.  On exit the block holds X (lower triangular).
static __device__ __noinline__ void chol_diag8_inv(double *H, int j0, int nb, int *fail) {
  if ((threadIdx.x & 31) != 0) return;
  double a[8][8], x[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) a[i][k] = (i < nb) ? H[tri(j0 + i) + j0 + k] : ((i == k) ? 1.0 : 0.0);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = a[j][j];
    if (!(d > 0.0) || !isfinite(d)) ok = false;
    const double rs = rsqrt(d);
    // row j of X before scaling: t_c = [c == j] - sum_{k=c}^{j-1} L_jk X_kc   (L_jk, k < j, are final already)
    double t[8];
#pragma unroll
    for (int c = 0; c < j; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int k = c; k < j; ++k) acc = fma(a[j][k], x[k][c], acc);
      t[c] = -acc;
    }
#pragma unroll
    for (int c = 0; c < j; ++c) x[j][c] = t[c] * rs;
    x[j][j] = rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k)
      if (i < nb) H[tri(j0 + i) + j0 + k] = x[i][k];
  *fail = ok ? 0 : 1;
}

// ---- panel rows [lo, Nrows) below the diagonal block at j0 (which holds the INVERSE X of its factor): one warp per
// 8 rows, L_rows = A_rows * X' as two DMMAs; results to H and, transposed, to S[c][i] (zero padded).
__device__ __forceinline__ void trsm_dmma8(double *H, int j0, int nb, int lo, int Nrows, double *S, int ld) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int lo8 = lo & ~7, hi8 = (Nrows + 7) & ~7;
  double bv[2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int kk = 4 * ks + lc;
    bv[ks] = (lr < nb && kk <= lr) ? H[tri(j0 + lr) + j0 + kk] : 0.0;   // B[k][c] = X[c][k]
  }
  for (int i0 = lo8 + 8 * w; i0 < hi8; i0 += 8 * IPM_NW) {
    const int i = i0 + lr;
    const bool iv = i >= lo && i < Nrows;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int kk = 4 * ks + lc;
      const double av = (iv && kk < nb) ? H[tri(i) + j0 + kk] : 0.0;
      dmma(c0, c1, av, bv[ks]);
    }
    const int cc = 2 * lc;
    if (iv) {
      if (cc < nb) H[tri(i) + j0 + cc] = c0;
      if (cc + 1 < nb) H[tri(i) + j0 + cc + 1] = c1;
    }
    S[cc * ld + i] = (iv && cc < nb) ? c0 : 0.0;
    S[(cc + 1) * ld + i] = (iv && cc + 1 < nb) ? c1 : 0.0;
  }
}

// ---- chol2 with in-register block inverses and DMMA panel solves (no separate inversion pass).
// jstart > 0: the columns before jstart are final and the diagonal block at jstart is already factored and inverted.
__device__ __forceinline__ bool chol3(double *H, int N, int Nrows, double *S, int ld, int *fail, long long *dbg = nullptr,
                                      int jstart = 0) {
  const int w = threadIdx.x >> 5;
  long long t0 = dbg ? clock64() : 0;
  if (jstart == 0) {
    __syncthreads();
    if (w == 0) chol_diag8_inv(H, 0, min(8, N), fail);
    __syncthreads();
    if (*fail) return false;
  }
  C2_MARK(0);
  for (int j0 = jstart; j0 < N; j0 += 16) {
    const int nb0 = min(8, N - j0), j1 = j0 + 8;
    trsm_dmma8(H, j0, nb0, j0 + nb0, Nrows, S, ld);
    __syncthreads();
    C2_MARK(1);
    if (j1 >= N) break;
    const int nb1 = min(8, N - j1), lo = j1 + nb1;
    // warp 0 updates the next diagonal tile first and factors it while the others update the rest of the block column
    if (w == 0) {
      syrk_col8(H, j1, j1 + 8 < Nrows ? j1 + 8 : Nrows, S, ld, 0, 1);
      __syncwarp();
      chol_diag8_inv(H, j1, nb1, fail);
    } else {
      syrk_col8(H, j1 + 8, Nrows, S, ld, w - 1, IPM_NW - 1, j1);
    }
    __syncthreads();
    if (*fail) return false;
    C2_MARK(2);
    trsm_dmma8(H, j1, nb1, lo, Nrows, S + 8 * ld, ld);
    __syncthreads();
    C2_MARK(1);
    if (lo >= N) break;
    if (w == 0) {
      syrk_rt16_sub(H, lo, Nrows, S, ld, 0, 1 << 20);
      __syncwarp();
      chol_diag8_inv(H, lo, min(8, N - lo), fail);
    } else {
      syrk_rt16_sub(H, lo, Nrows, S, ld, w, IPM_NW - 1);
    }
    __syncthreads();
    C2_MARK(3);
    if (*fail) return false;
  }
  return true;
}

// =====================================================================================================================
// chol4: 16-column panels with the WHOLE 16 x 16 diagonal block factored one panel ahead.
//
// chol3 runs four CTA-wide phases per 16 columns (solve | 8-column update + diagonal block | solve | trailing update +
// diagonal block) and in two of them everybody waits for the pivot chain of an 8 x 8 diagonal block (~2 k cycles of one
// thread).  Here a panel is two phases:
//   panel phase     every warp takes row tiles below the diagonal block and does, for its own eight rows and without any
//                   CTA barrier in between, solve against X(j0) -> 8-column update of the second sub-panel -> solve
//                   against X(j1) (all three are row-local), leaving the sixteen solved columns transposed in S;
//   trailing phase  warps 1 .. NW-1 apply the rank-16 update; warp 0 updates the NEXT 16 x 16 diagonal block first and
//                   factors it completely under that update: diagonal tile, the tile below it, second diagonal tile.
// The pivot chain is thereby hidden as long as the trailing update is longer than it, and there are two CTA barriers
// per 16 columns instead of four.  Ragged ends (N not a multiple of 16; the carried right-hand-side rows share the last
// tile row with the last diagonal block) are left to chol3's loop, entered at the first panel that is not full.

// 8 x 8 diagonal block at j0 (full: 8 valid rows): Cholesky + inverse by ONE WARP.  Every lane runs the factorisation
// redundantly in registers (same values, no communication); lane c < 8 forms column c of X = L^-1 alongside, so the
// instruction stream of the warp holds the factorisation plus ONE column of the inverse instead of all eight (the
// one-thread version issues ~250 dependent-ish FP64 instructions; this one ~150).  A non-positive pivot SETS *fail
// (never clears it).
static __device__ __noinline__ void chol_diag8_invw(double *H, int j0, int *fail) {
  const int lane = threadIdx.x & 31, c = lane & 7;
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) a[i][k] = H[tri(j0 + i) + j0 + k];
  __syncwarp();
  bool ok = true;
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = a[j][j];
    if (!(d > 0.0) || !isfinite(d)) ok = false;
    const double rs = rsqrt(d);
    double t = (c == j) ? 1.0 : 0.0;            // row j of X, column c: [c == j] - sum_{k<j} L_jk X_kc
#pragma unroll
    for (int k = 0; k < j; ++k) t = fma(-a[j][k], x[k], t);
    x[j] = t * rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= c) H[tri(j0 + i) + j0 + c] = x[i];
  }
  if (lane == 0 && !ok) *fail = 1;
  __syncwarp();
}

// the rest of the 16 x 16 diagonal block at la (both 8 x 8 diagonal tiles full), given its first diagonal tile factored and
// inverted: tile (lb, la) := tile X(la)', tile (lb, lb) -= L(lb, la) L(lb, la)', factor + invert (lb, lb).  One whole warp.
__device__ __forceinline__ void chol_diag16_rest(double *H, int la, int *fail) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  const int lb = la + 8, cc = 2 * lc;
  double c0 = 0.0, c1 = 0.0;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int kk = 4 * ks + lc;
    const double av = H[tri(lb + lr) + la + kk];
    const double bv = (kk <= lr) ? H[tri(la + lr) + la + kk] : 0.0;     // B[k][c] = X[c][k]
    dmma(c0, c1, av, bv);
  }
  double *Tl = H + tri(lb + lr) + la;
  __syncwarp();
  Tl[cc] = c0; Tl[cc + 1] = c1;
  __syncwarp();
  const double a0 = Tl[lc], a1 = Tl[4 + lc];      // A[m][k] = L[lb+m][la+k] and B[k][n] = L[lb+n][la+k]: the same fragment
  double *Td = H + tri(lb + lr) + lb;
  double d0 = (cc <= lr) ? Td[cc] : 0.0, d1 = (cc + 1 <= lr) ? Td[cc + 1] : 0.0;
  dmma(d0, d1, -a0, a0);
  dmma(d0, d1, -a1, a1);
  if (cc <= lr) Td[cc] = d0;
  if (cc + 1 <= lr) Td[cc + 1] = d1;
  __syncwarp();
  chol_diag8_invw(H, lb, fail);
}

// panel phase of a full panel (16 columns from j0, diagonal block factored): rows [lo2, Nrows), lo2 = j0 + 16, one row
// tile per warp at a time; the solved columns go to H and, transposed, to S[0 .. 15][i] (zero for i >= Nrows).
// (t0, tstride: the calling warp takes the row tiles t0, t0 + tstride, ... counted from row lo2)
__device__ __forceinline__ void panel16_rows(double *H, int j0, int Nrows, double *S, int ld, int t0 = -1, int tstride = IPM_NW) {
  const int lane = threadIdx.x & 31, w = t0 >= 0 ? t0 : (int)(threadIdx.x >> 5);
  const int lr = lane >> 2, lc = lane & 3, cc = 2 * lc;
  const int j1 = j0 + 8, lo2 = j0 + 16, hi8 = (Nrows + 7) & ~7;
  double bx0[2], bx1[2], bl[2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int kk = 4 * ks + lc;
    bx0[ks] = (kk <= lr) ? H[tri(j0 + lr) + j0 + kk] : 0.0;
    bx1[ks] = (kk <= lr) ? H[tri(j1 + lr) + j1 + kk] : 0.0;
    bl[ks] = H[tri(j1 + lr) + j0 + kk];            // B[k][n] = L[j1 + n][j0 + k]
  }
  for (int i0 = lo2 + 8 * w; i0 < hi8; i0 += 8 * tstride) {
    const int i = i0 + lr;
    const bool iv = i < Nrows;
    double *Hi = H + tri(min(i, Nrows - 1));
    // (a) columns j0 .. j0+7: L = A X(j0)'
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) dmma(c0, c1, iv ? Hi[j0 + 4 * ks + lc] : 0.0, bx0[ks]);
    __syncwarp();
    if (iv) { Hi[j0 + cc] = c0; Hi[j0 + cc + 1] = c1; }
    S[cc * ld + i] = iv ? c0 : 0.0;
    S[(cc + 1) * ld + i] = iv ? c1 : 0.0;
    // the tile of the second sub-panel, as accumulator
    double d0 = iv ? Hi[j1 + cc] : 0.0, d1 = iv ? Hi[j1 + cc + 1] : 0.0;
    __syncwarp();
    // (b) columns j1 .. j1+7 -= L(i, j0..) L(j1.., j0..)'
    dmma(d0, d1, -S[lc * ld + i], bl[0]);
    dmma(d0, d1, -S[(4 + lc) * ld + i], bl[1]);
    if (iv) { Hi[j1 + cc] = d0; Hi[j1 + cc + 1] = d1; }
    __syncwarp();
    // (c) L = (updated tile) X(j1)'
    double e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) dmma(e0, e1, iv ? Hi[j1 + 4 * ks + lc] : 0.0, bx1[ks]);
    __syncwarp();
    if (iv) { Hi[j1 + cc] = e0; Hi[j1 + cc + 1] = e1; }
    S[(8 + cc) * ld + i] = iv ? e0 : 0.0;
    S[(9 + cc) * ld + i] = iv ? e1 : 0.0;
  }
}

__device__ __forceinline__ bool chol4(double *H, int N, int Nrows, double *S, int ld, int *fail, long long *dbg = nullptr) {
  const int w = threadIdx.x >> 5;
  if (N < 32) return chol3(H, N, Nrows, S, ld, fail, dbg);
  long long t0 = dbg ? clock64() : 0;
  if (threadIdx.x == 0) *fail = 0;
  __syncthreads();
  if (w == 0) { chol_diag8_invw(H, 0, fail); chol_diag16_rest(H, 0, fail); }
  __syncthreads();
  if (*fail) return false;
  C2_MARK(0);
  int j0 = 0;
  for (;;) {                                  // invariant: j0 + 16 <= N and the 16 x 16 diagonal block at j0 is factored
    const int lo = j0 + 16;
    const bool next_full = lo + 16 <= N;
#ifndef BNLP_CHOL4_NO_EARLY_CHAIN
    if (next_full) {
      // Warp 0 solves the two row tiles of the NEXT diagonal block itself and goes straight on to that block's update and
      // pivot chain; it only ARRIVES at the barrier that separates the other warps' panel rows from their trailing update
      // (named barrier 1: 15 warps wait, warp 0 does not), so in the late panels — where the chain is longer than the
      // trailing update — the chain no longer starts behind everybody's panel phase.
      if (w == 0) {
        panel16_rows(H, j0, Nrows, S, ld, 0, 1 << 20);
        panel16_rows(H, j0, Nrows, S, ld, 1, 1 << 20);
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"r"(IPM_NT) : "memory");
        syrk_rt16_sub(H, lo, Nrows, S, ld, 0, 1 << 20);
        __syncwarp();
        chol_diag8_invw(H, lo, fail);
        chol_diag16_rest(H, lo, fail);
      } else {
#ifndef BNLP_CHOL4_PLAIN_ORDER
        // work goes first to the warps that do NOT share warp 0's SM sub-partition (warps 4, 8, 12 do): in the late panels,
        // where there are fewer units than warps, the pivot chain then has its FP64 pipe to itself
        const int rk = (w & 3) ? w - (w >> 2) : (IPM_NW - IPM_NW / 4) + (w >> 2);       // 1 .. NW-1, a permutation
#else
        const int rk = w;
#endif
        panel16_rows(H, j0, Nrows, S, ld, 1 + rk, IPM_NW - 1);       // tiles 2, 3, ... over warps 1 .. NW-1
        asm volatile("bar.sync 1, %0;" ::"r"(IPM_NT) : "memory");
        syrk_rt16_sub(H, lo, Nrows, S, ld, rk, IPM_NW - 1);
      }
      __syncthreads();
      C2_MARK(3);
      if (*fail) return false;
      j0 = lo;
      continue;
    }
#endif
    panel16_rows(H, j0, Nrows, S, ld);
    __syncthreads();
    C2_MARK(1);
    if (lo >= N) return true;                 // (N a multiple of 16: only carried rows were left)
    if (w == 0) {
      syrk_rt16_sub(H, lo, Nrows, S, ld, 0, 1 << 20);
      __syncwarp();
      if (next_full) { chol_diag8_invw(H, lo, fail); chol_diag16_rest(H, lo, fail); }
      else chol_diag8_inv(H, lo, min(8, N - lo), fail);   // (*fail is 0 here: it was checked after the last barrier)
    } else {
      syrk_rt16_sub(H, lo, Nrows, S, ld, w, IPM_NW - 1);
    }
    __syncthreads();
    C2_MARK(3);
    if (*fail) return false;
    j0 = lo;
    if (!next_full) break;
  }
  return chol3(H, N, Nrows, S, ld, fail, dbg, j0);
}

}  // namespace BNLP_NS
