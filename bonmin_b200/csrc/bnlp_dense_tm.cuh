// Per-child dense kernels of the cooperative IPM kernel on a TILE-MAJOR condensed matrix (included by
// bnlp_ipm_coop.cuh after tri() and dmma()).
//
// Storage.  The lower triangle of the augmented matrix [H; rhs0'; rhs1'] is kept in shared memory as 8 x 8 tiles,
// tile (bi, bk), bk <= bi, at tri(bi) * 64 + bk * 64 doubles; inside a tile element (r, c) sits at r * 8 + (c ^ swz(r)),
// swz(r) = 4 for r in {2, 3, 6, 7}.  With that XOR
//   * an m8n8k4 OPERAND fragment (lane (lr, lc) reads [lr][4 ks + lc]) touches 16 distinct 8-byte banks per half warp,
//   * an ACCUMULATOR fragment (lane reads/writes [lr][2 lc], [lr][2 lc + 1]) is one aligned 16-byte access, 128
//     contiguous bytes per quarter warp,
// so the factor can be used as a DMMA operand where it lies (the packed row-major layout needed a transposed staging
// copy of every panel) and every read-modify-write of a tile is conflict free (the packed layout paid ~2.5 wavefronts
// per request there: ncu counted 242 M bank conflicts per launch, profiles/r1k_ipm_coop_c3_ncu_raw.txt).
// The LAST tile row is stored 4 rows high when it holds at most 4 rows (n = 200: rows 200 .. 202 = the phase-1 elastic
// and the two right-hand sides), which is what lets n = 200 fit 227 KB.
//
//   tm_syrk_stage   H (+)= M' M from staged rows (J' Sigma J), 2 x 4 tile units, accumulators written with 16-byte
//                   accesses; FIRST = store instead of add (no zeroing pass, no read)
//   chol_tm         right-looking Cholesky, 8-column panels, look-ahead: warp 0 updates and factors the next diagonal
//                   tile (one thread, registers) and warps 0 .. PW-1 solve the next panel while the other warps apply
//                   the current panel to the trailing tiles (units handed out by a shared-memory counter; the panel
//                   warps join when their chain is done).  One CTA barrier per panel.
//   backsolve_tm    L' x = w for the two right-hand sides, one warp each, no CTA barrier inside the block loop
#pragma once

namespace BNLP_NS {

struct TmLayout {
  int nt;       // tile rows
  int last_h;   // height of the last tile row: 8 or 4
};

__host__ __device__ inline TmLayout tm_layout(int nrows) {
  TmLayout L;
  L.nt = (nrows + 7) >> 3;
  const int rem = nrows - ((L.nt - 1) << 3);
  L.last_h = (rem <= 4) ? 4 : 8;
  return L;
}
__host__ __device__ inline int tm_size(const TmLayout &L) { return ((L.nt - 1) * L.nt / 2) * 64 + L.nt * 8 * L.last_h; }
__host__ __device__ inline int tm_tile(const TmLayout &L, int bi, int bk) {
  return ((bi * (bi + 1)) >> 1) * 64 + bk * ((bi == L.nt - 1) ? 8 * L.last_h : 64);
}
__host__ __device__ inline int tm_swz(int r) { return (r & 2) << 1; }
__host__ __device__ inline int tm_idx(const TmLayout &L, int i, int k) {
  return tm_tile(L, i >> 3, k >> 3) + ((i & 7) << 3) + ((k & 7) ^ tm_swz(i & 7));
}
// lane (lr, lc): is row lr of tile row bi stored?
__device__ __forceinline__ bool tm_row_ok(const TmLayout &L, int bi, int lr) { return bi != L.nt - 1 || lr < L.last_h; }

// operand fragment of k-step ks (columns 4 ks .. 4 ks + 3) of the tile at T
__device__ __forceinline__ double tm_frag(const double *T, int lr, int lc, int ks) {
  return T[(lr << 3) + ((4 * ks + lc) ^ tm_swz(lr))];
}
__device__ __forceinline__ double2 *tm_acc(double *T, int lr, int lc) {
  return reinterpret_cast<double2 *>(T + (lr << 3) + ((2 * lc) ^ tm_swz(lr)));
}

__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// unit u of a triangular tile region of m tile rows: row pair r2 (tile rows 2 r2, 2 r2 + 1), column quad cq (tile columns
// 4 cq .. 4 cq + 3), numbered from the BOTTOM row pair up so that the full 8-tile units are handed out first and the
// short diagonal ones fill the tail.
__device__ __forceinline__ int tm_units(int m) {
  const int nrp = (m + 1) >> 1;          // row pairs; pairs (2a, 2a+1) own a + 1 units each
  const int a = nrp >> 1;
  return a * (a + 1) + ((nrp & 1) ? a + 1 : 0);
}
__device__ __forceinline__ void tm_unit_decode(int m, int u, int &r2, int &cq) {
  const int nrp = (m + 1) >> 1;
  r2 = nrp - 1;
  for (;;) {
    const int c = (r2 >> 1) + 1;
    if (u < c) break;
    u -= c;
    --r2;
  }
  cq = u;
}

// ---- H(tile rows < ntr) (+)= M' M.  M: K staged rows of length ld (ld = 4 mod 8 doubles: conflict-free fragments), zero
// beyond the last real column.  Units are taken from *ctr (shared memory, zeroed by the caller before a CTA barrier).
// sgn (optional, K values of +-1 in shared memory): staged row c enters as sgn[c] * m_c m_c' (the A fragments are scaled).
template <bool FIRST>
__device__ __forceinline__ void tm_syrk_stage(double *H, const TmLayout &L, int ntr, const double *M, int ld, int K, int *ctr,
                                              const double *sgn = nullptr) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nunits = tm_units(ntr);
  for (;;) {
    int u = 0;
    if (lane == 0) u = atomicAdd(ctr, 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= nunits) break;
    int r2, cq;
    tm_unit_decode(ntr, u, r2, cq);
    const int bi0 = 2 * r2, bk0 = 4 * cq;
    const bool has1 = bi0 + 1 < ntr;
    const int ntc = min(4, bi0 + (has1 ? 1 : 0) - bk0 + 1);
    // the accumulators START from the stored tiles (all loads issued up front, under the first operand loads): a
    // load-add-store epilogue is serialised tile by tile, because the compiler cannot prove that the tiles are distinct
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[a][t][0] = acc[a][t][1] = 0.0;
        if (!FIRST && t < ntc && bk0 + t <= bi0 + a && bi0 + a < ntr && tm_row_ok(L, bi0 + a, lr)) {
          const double2 h = *tm_acc(H + tm_tile(L, bi0 + a, bk0 + t), lr, lc);
          acc[a][t][0] = h.x; acc[a][t][1] = h.y;
        }
      }
    const double *Ma = M + lc * ld + (bi0 << 3) + lr, *Mb = M + lc * ld + (bk0 << 3) + lr;
    if (ntc == 4 && has1 && bk0 + 3 <= bi0) {          // all eight tiles on or below the diagonal
#pragma unroll 2
      for (int k = 0; k < K; k += 4) {
        const double sg = sgn ? sgn[k + lc] : 1.0;
        const double a0 = sg * Ma[k * ld], a1 = sg * Ma[k * ld + 8];
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
        const double sg = sgn ? sgn[k + lc] : 1.0;
        const double a0 = sg * Ma[k * ld], a1 = has1 ? sg * Ma[k * ld + 8] : 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < ntc) {
            const double v = Mb[k * ld + 8 * t];
            if (bk0 + t <= bi0) dmma(acc[0][t][0], acc[0][t][1], a0, v);
            if (has1) dmma(acc[1][t][0], acc[1][t][1], a1, v);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int bi = bi0 + a;
      if (bi < ntr && tm_row_ok(L, bi, lr)) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int bk = bk0 + t;
          if (t < ntc && bk <= bi) {
            *tm_acc(H + tm_tile(L, bi, bk), lr, lc) = make_double2(acc[a][t][0], acc[a][t][1]);
          }
        }
      }
    }
  }
}

// ---- trailing unit of the factorisation: tiles (i, j), i in {bi0, bi0 + 1}, j in {bj0 .. bj0 + 3}, j <= i < nt:
// T(i, j) -= T(i, k) T(j, k)'   (operands read where the factor lies).
// ONE branch-free body for every unit: tiles above the diagonal / beyond the last tile row are computed on clamped
// addresses and not stored (the FP64 pipe has slack here, the pivot chain of the diagonal tiles is the floor), because
// with a branch per tile the loads, DMMAs and stores of one tile stay in its own basic block and a ragged unit costs
// 1 300 cycles instead of 450 (measured, scripts/micro/chol_tm_bench.cu).  A fragments are negated and the
// accumulators start from the stored tiles: T := T + (-A) B', one 16-byte store per tile.
__device__ __forceinline__ void tm_trail_unit(double *H, const TmLayout &L, int k, int bi0, int bj0) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt = L.nt;
  const int fo = (lr << 3) + (lc ^ tm_swz(lr)), f1 = ((4 + lc) ^ tm_swz(lr)) - (lc ^ tm_swz(lr)), ao = (lr << 3) + ((2 * lc) ^ tm_swz(lr));
  const bool thin_ok = lr < L.last_h;           // rows of the last tile row that exist
  int bi[2];
  bool rv[2];
  bi[0] = bi0; bi[1] = min(bi0 + 1, nt - 1);
  rv[0] = bi0 != nt - 1 || thin_ok;
  rv[1] = bi0 + 1 < nt && (bi0 + 1 != nt - 1 || thin_ok);
  double a[2][2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double *At = H + tm_tile(L, bi[q], k) + fo;
    a[q][0] = rv[q] ? -At[0] : 0.0;
    a[q][1] = rv[q] ? -At[f1] : 0.0;
  }
  double b[4][2], acc[2][4][2];
  bool ok[2][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int bj = min(bj0 + t, nt - 1);
    const bool bv = bj != nt - 1 || thin_ok;
    const double *Bt = H + tm_tile(L, bj, k) + fo;
    b[t][0] = bv ? Bt[0] : 0.0;
    b[t][1] = bv ? Bt[f1] : 0.0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      ok[q][t] = rv[q] && bj0 + t <= bi0 + q;
      const double2 *C = reinterpret_cast<const double2 *>(H + tm_tile(L, bi[q], min(bj0 + t, bi[q])) + ao);
      double2 h = make_double2(0.0, 0.0);
      if (ok[q][t]) h = *C;
      acc[q][t][0] = h.x; acc[q][t][1] = h.y;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) { dmma(acc[0][t][0], acc[0][t][1], a[0][0], b[t][0]); dmma(acc[1][t][0], acc[1][t][1], a[1][0], b[t][0]); }
#pragma unroll
  for (int t = 0; t < 4; ++t) { dmma(acc[0][t][0], acc[0][t][1], a[0][1], b[t][1]); dmma(acc[1][t][0], acc[1][t][1], a[1][1], b[t][1]); }
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (ok[q][t])
        *reinterpret_cast<double2 *>(H + tm_tile(L, bi[q], min(bj0 + t, bi[q])) + ao) = make_double2(acc[q][t][0], acc[q][t][1]);
}

// ---- T(i, k+1) -= T(i, k) T(k+1, k)'  for one tile row i >= k + 1 (the look-ahead update of the next panel)
__device__ __forceinline__ void tm_colupd_tile(double *H, const TmLayout &L, int k, int i) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const bool rv = tm_row_ok(L, i, lr), bv = tm_row_ok(L, k + 1, lr);
  const double *At = H + tm_tile(L, i, k), *Bt = H + tm_tile(L, k + 1, k);
  const double a0 = rv ? -tm_frag(At, lr, lc, 0) : 0.0, a1 = rv ? -tm_frag(At, lr, lc, 1) : 0.0;
  const double b0 = bv ? tm_frag(Bt, lr, lc, 0) : 0.0, b1 = bv ? tm_frag(Bt, lr, lc, 1) : 0.0;
  double2 *p = tm_acc(H + tm_tile(L, i, k + 1), lr, lc);
  double c0 = 0.0, c1 = 0.0;
  if (rv) { const double2 h = *p; c0 = h.x; c1 = h.y; }
  dmma(c0, c1, a0, b0);
  dmma(c0, c1, a1, b1);
  if (rv) *p = make_double2(c0, c1);
}

// ---- T(i, k) := T(i, k) X'   (X = inverse of the factor of diagonal tile k, lower triangular, stored in that tile)
__device__ __forceinline__ void tm_trsm_tile(double *H, const TmLayout &L, int k, int i) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const bool rv = tm_row_ok(L, i, lr), xv = tm_row_ok(L, k, lr);
  double *At = H + tm_tile(L, i, k);
  const double *X = H + tm_tile(L, k, k);
  const double a0 = rv ? tm_frag(At, lr, lc, 0) : 0.0, a1 = rv ? tm_frag(At, lr, lc, 1) : 0.0;
  // B[kk][c] = X[c][kk], zero above the diagonal of X
  const double b0 = (xv && lc <= lr) ? tm_frag(X, lr, lc, 0) : 0.0, b1 = (xv && 4 + lc <= lr) ? tm_frag(X, lr, lc, 1) : 0.0;
  double c0 = 0.0, c1 = 0.0;
  dmma(c0, c1, a0, b0);
  dmma(c0, c1, a1, b1);
  if (rv) *tm_acc(At, lr, lc) = make_double2(c0, c1);
}

// ---- diagonal tile: Cholesky of its leading nb x nb block by lane 0 (registers, right-looking: the FMAs of a step are
// independent and issue under the rsqrt latency of the pivot chain), rows nb .. nrows-1 of the tile (right-hand sides
// that share the tile) are forward-substituted against it; then lanes 0 .. 7 invert the factor column by column and the
// tile's leading block is overwritten with X = L^-1.  Called by ONE WHOLE WARP.  *fail = 1 on a non-positive pivot.
// ---- the same two tile operations for up to three tile rows i0, i0 + st, i0 + 2 st at once: every load is issued before
// the first DMMA (one tile at a time the load -> DMMA -> DMMA -> store chain of ~300 cycles is fully exposed, and the
// compiler may not overlap two tiles itself because it cannot prove that they are distinct)
__device__ __forceinline__ void tm_colupd_tiles3(double *H, const TmLayout &L, int k, int i0, int st) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt = L.nt;
  const double *Bt = H + tm_tile(L, k + 1, k);
  const bool bv = tm_row_ok(L, k + 1, lr);
  const double b0 = bv ? tm_frag(Bt, lr, lc, 0) : 0.0, b1 = bv ? tm_frag(Bt, lr, lc, 1) : 0.0;
  double a[3][2], c[3][2];
  bool rv[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int i = i0 + q * st;
    rv[q] = i < nt && tm_row_ok(L, i, lr);
    const int ii = min(i, nt - 1);
    const double *At = H + tm_tile(L, ii, k);
    a[q][0] = rv[q] ? -tm_frag(At, lr, lc, 0) : 0.0;
    a[q][1] = rv[q] ? -tm_frag(At, lr, lc, 1) : 0.0;
    c[q][0] = c[q][1] = 0.0;
    if (rv[q]) { const double2 h = *tm_acc(H + tm_tile(L, ii, k + 1), lr, lc); c[q][0] = h.x; c[q][1] = h.y; }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) dmma(c[q][0], c[q][1], a[q][0], b0);
#pragma unroll
  for (int q = 0; q < 3; ++q) dmma(c[q][0], c[q][1], a[q][1], b1);
#pragma unroll
  for (int q = 0; q < 3; ++q)
    if (rv[q]) *tm_acc(H + tm_tile(L, i0 + q * st, k + 1), lr, lc) = make_double2(c[q][0], c[q][1]);
}
__device__ __forceinline__ void tm_trsm_tiles3(double *H, const TmLayout &L, int k, int i0, int st) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt = L.nt;
  const double *X = H + tm_tile(L, k, k);
  const bool xv = tm_row_ok(L, k, lr);
  const double b0 = (xv && lc <= lr) ? tm_frag(X, lr, lc, 0) : 0.0, b1 = (xv && 4 + lc <= lr) ? tm_frag(X, lr, lc, 1) : 0.0;
  double a[3][2], c[3][2];
  bool rv[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int i = i0 + q * st;
    rv[q] = i < nt && tm_row_ok(L, i, lr);
    const double *At = H + tm_tile(L, min(i, nt - 1), k);
    a[q][0] = rv[q] ? tm_frag(At, lr, lc, 0) : 0.0;
    a[q][1] = rv[q] ? tm_frag(At, lr, lc, 1) : 0.0;
    c[q][0] = c[q][1] = 0.0;
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) dmma(c[q][0], c[q][1], a[q][0], b0);
#pragma unroll
  for (int q = 0; q < 3; ++q) dmma(c[q][0], c[q][1], a[q][1], b1);
#pragma unroll
  for (int q = 0; q < 3; ++q)
    if (rv[q]) *tm_acc(H + tm_tile(L, i0 + q * st, k), lr, lc) = make_double2(c[q][0], c[q][1]);
}

// ---- full 8 x 8 diagonal tile: Cholesky AND inverse of the factor in the registers of lane 0.  Row j of X = L^-1 is
// formed in the same unrolled step as pivot j, so its independent FMAs issue under the rsqrt latency of the pivot chain
// (the chain — rsqrt, scale, update of the next pivot — is ~92 cycles per pivot and is the floor of the factorisation).
// 16-byte loads / stores of whole tile rows; the upper triangle of X is stored as zeros.
static __device__ __noinline__ void tm_potrf_full(double *T, int *fail) {
  if ((threadIdx.x & 31) != 0) return;
  double a[8][8], x[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int p = 0; p < 8; p += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(T + (i << 3) + p);
      a[i][p ^ tm_swz(i)] = v.x; a[i][(p + 1) ^ tm_swz(i)] = v.y;
    }
  }
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = a[j][j];
    if (!(d > 0.0) || !isfinite(d)) ok = false;
    const double rs = rsqrt(d);
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
    for (int c = j + 1; c < 8; ++c) x[j][c] = 0.0;
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
    for (int p = 0; p < 8; p += 2)
      *reinterpret_cast<double2 *>(T + (i << 3) + p) = make_double2(x[i][p ^ tm_swz(i)], x[i][(p + 1) ^ tm_swz(i)]);
  *fail = ok ? 0 : 1;
}

template <bool FULL>
static __device__ __noinline__ void tm_potrf_t(double *T, int nb_, int nrows_, int *fail, long long *pd = nullptr) {
  const int nb = FULL ? 8 : nb_, nrows = FULL ? 8 : nrows_;   // FULL: every predicate below folds away
  const int lane = threadIdx.x & 31;
  long long p0 = pd ? clock64() : 0, p1 = 0, p2 = 0;
  if (lane == 0) {
    double a[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k <= i; ++k) a[i][k] = (i < nrows) ? T[(i << 3) + (k ^ tm_swz(i))] : ((i == k) ? 1.0 : 0.0);
    if (pd) p1 = clock64();
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < nb) {
        const double d = a[j][j];
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double rs = rsqrt(d);
        a[j][j] = rs;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i][j] *= rs;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
          for (int k = j + 1; k <= i; ++k)
            if (k < nb || k == i) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
      }
    }
    if (pd) p2 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k <= i; ++k)
        if (i < nrows && (k < nb)) T[(i << 3) + (k ^ tm_swz(i))] = a[i][k];
    *fail = ok ? 0 : 1;
  }
  __syncwarp();
  const long long p3 = pd ? clock64() : 0;
  // inverse of the nb x nb factor (strictly lower part = L, diagonal = 1 / L_kk): lane c owns column c of X
  // (every lane runs the arithmetic — the loads are broadcasts — and lanes 0 .. 7 store after a warp barrier)
  const int c = lane & 7;
  double s[8], x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double lkk = (k < nb) ? T[(k << 3) + (k ^ tm_swz(k))] : 1.0;
    const double xk = (k == c) ? lkk : ((k > c) ? -lkk * s[k] : 0.0);
    x[k] = xk;
#pragma unroll
    for (int i = k + 1; i < 8; ++i)
      if (i < nb) s[i] = fma(T[(i << 3) + (k ^ tm_swz(i))], xk, s[i]);
  }
  __syncwarp();
  if (lane < 8 && c < nb) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= c && i < nb) T[(i << 3) + (c ^ tm_swz(i))] = x[i];
  }
  __syncwarp();
  if (pd && lane == 0) { const long long p4 = clock64(); pd[0] = p1 - p0; pd[1] = p2 - p1; pd[2] = p3 - p2; pd[3] = p4 - p3; }
}
__device__ __forceinline__ void tm_potrf(double *T, int nb, int nrows, int *fail, long long *pd = nullptr) {
  if (nb == 8 && nrows == 8) { tm_potrf_full(T, fail); __syncwarp(); }
  else tm_potrf_t<false>(T, nb, nrows, fail, pd);
}

// ---- blocked Cholesky of rows/columns [0, N) of the tile-major matrix; rows [N, Nrows) are carried along (right-hand
// sides: their forward substitution is the panel solve).  On exit the diagonal tiles hold the INVERSES of the diagonal
// blocks of L, everything below holds L.  Returns false (block-uniform) on a non-positive pivot.
// ctr: two shared-memory ints.  Named barrier 1 is used by the panel warps.
// dbg (optional, shared memory): cycles of warp 0 in [0] diagonal update + factor, [1] panel barrier, [2] panel solve,
// [3] trailing units, [4] CTA barrier; of the last warp in [5] trailing units, [6] CTA barrier.
__device__ __forceinline__ bool chol_tm(double *H, const TmLayout &L, int N, int Nrows, int *fail, int *ctr,
                                        long long *dbg = nullptr) {
  // Per panel k (column k final on entry):
  //   phase 1  warp 0: update and factor the next diagonal tile (the pivot chain: ~1.9 k cycles, the floor of the whole
  //            factorisation).  It runs ALONE on SM sub-partition 0 (warps 4, 8, 12 wait): the chain is DFMA latency, and
  //            behind the DMMAs of other warps of its sub-partition every step of it queues for the FP64 pipe.
  //            The warps of sub-partitions 1-3 take units from a counter: first the rest of column k+1 (look-ahead
  //            update, three tiles per unit), then the trailing tiles (2 x 4 tiles per unit).
  //   phase 2  every warp solves its tiles of column k+1 against the new diagonal block.
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool bulk = (w & 3) != 0;
  const int nt = L.nt, ntN = (N + 7) >> 3;
  long long tq = 0;
#define TM_MARK(i) do { if (dbg && lane == 0) { const long long t_ = clock64(); dbg[i] += t_ - tq; tq = t_; } } while (0)
  __syncthreads();
  if (w == 0) tm_potrf(H + tm_tile(L, 0, 0), min(8, N), min(8, Nrows), fail);
  if (threadIdx.x == 32) ctr[0] = 0;
  __syncthreads();
  if (*fail) return false;
  for (int i = 1 + w; i < nt; i += 3 * IPM_NW) tm_trsm_tiles3(H, L, 0, i, IPM_NW);
  __syncthreads();
  if (dbg) tq = clock64();
  for (int k = 0; k + 1 < ntN; ++k) {
    const int c1 = k + 1;
    int *cnt = ctr + (k & 1);
    if (threadIdx.x == 32) ctr[(k + 1) & 1] = 0;
    if (w == 0) {
      tm_colupd_tile(H, L, k, c1);
      __syncwarp();
      tm_potrf(H + tm_tile(L, c1, c1), min(8, N - 8 * c1), min(8, Nrows - 8 * c1), fail);
      TM_MARK(0);
    } else if (bulk) {
      const int ncol = nt - (c1 + 1);               // tiles of column k+1 below its diagonal
      const int ncu = (ncol + 2) / 3;
      const int m = nt - (k + 2);
      const int nunits = ncu + (m > 0 ? tm_units(m) : 0);
      for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(cnt, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nunits) break;
        if (u < ncu) {
          tm_colupd_tiles3(H, L, k, c1 + 1 + u, ncu);
        } else {
          int r2, cq;
          tm_unit_decode(m, u - ncu, r2, cq);
          tm_trail_unit(H, L, k, k + 2 + 2 * r2, k + 2 + 4 * cq);
        }
      }
      if (w == IPM_NW - 1) TM_MARK(5);
    }
    __syncthreads();
    if (dbg && threadIdx.x == 0 && k < 28) dbg[16 + 2 * k] = clock64();
    if (w == 0) TM_MARK(1);
    if (w == IPM_NW - 1) TM_MARK(6);
    if (*fail) return false;
    if (dbg && k == 2 && lane == 0) dbg[80 + 2 * w] = clock64();
    for (int i = c1 + 1 + w; i < nt; i += 3 * IPM_NW) tm_trsm_tiles3(H, L, c1, i, IPM_NW);
    if (dbg && k == 2 && lane == 0) dbg[81 + 2 * w] = clock64();
    if (w == 0) TM_MARK(2);
    if (w == IPM_NW - 1) TM_MARK(3);
    __syncthreads();
    if (dbg && threadIdx.x == 0 && k < 28) dbg[17 + 2 * k] = clock64();
    if (w == 0) TM_MARK(4);
    if (w == IPM_NW - 1) TM_MARK(7);
  }
#undef TM_MARK
  return true;
}

// =====================================================================================================================
// LEFT-LOOKING variant.  The right-looking update above re-reads and re-writes every trailing tile once per panel: 88
// shared-memory wavefronts per 16 DMMAs, i.e. the trailing updates are bound by shared-memory bandwidth (~40 % of the
// FP64 pipe measured, scripts/micro/chol_tm_bench.cu) and the early panels are not hidden under the pivot chain.
// Left-looking, a tile of block column j accumulates ALL its updates sum_{k<j} L(i,k) L(j,k)' in registers and is
// written once: 351 tile writes instead of ~3 000 read-modify-writes for n = 200.
//
// R tile rows i0, i0 + st (R = 2) of block column j:  T(i, j) := T(i, j) - sum_{k<j} T(i, k) T(j, k)'.
// Two accumulator sets per tile (k-step 0 / 1 of every k) halve the dependent DMMA chain.
template <int R>
__device__ __forceinline__ void tm_ll_update(double *H, const TmLayout &L, int j, int i0, int st) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt = L.nt;
  const int fo = (lr << 3) + (lc ^ tm_swz(lr)), f1 = ((4 + lc) ^ tm_swz(lr)) - (lc ^ tm_swz(lr)), ao = (lr << 3) + ((2 * lc) ^ tm_swz(lr));
  const bool thin_ok = lr < L.last_h;
  const double *Brow = H + tm_tile(L, j, 0) + fo;
  const int bst = (j == nt - 1) ? 8 * L.last_h : 64;
  const bool bv = j != nt - 1 || thin_ok;
  const double *Arow[R];
  int ast[R];
  bool rv[R];
  double s0[R][2], s1[R][2];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * st;
    const int ii = min(i, nt - 1);
    rv[r] = i < nt && (ii != nt - 1 || thin_ok);
    Arow[r] = H + tm_tile(L, ii, 0) + fo;
    ast[r] = (ii == nt - 1) ? 8 * L.last_h : 64;
    s0[r][0] = s0[r][1] = s1[r][0] = s1[r][1] = 0.0;
  }
#pragma unroll 4
  for (int k = 0; k < j; ++k) {
    const double b0 = bv ? Brow[k * bst] : 0.0, b1 = bv ? Brow[k * bst + f1] : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double a0 = rv[r] ? Arow[r][k * ast[r]] : 0.0, a1 = rv[r] ? Arow[r][k * ast[r] + f1] : 0.0;
      dmma(s0[r][0], s0[r][1], a0, b0);
      dmma(s1[r][0], s1[r][1], a1, b1);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (rv[r]) {
      double2 *C = reinterpret_cast<double2 *>(H + tm_tile(L, i0 + r * st, j) + ao);
      double2 h = *C;
      h.x -= s0[r][0] + s1[r][0]; h.y -= s0[r][1] + s1[r][1];
      *C = h;
    }
  }
}

// ---- the same update for k in [k0, k1) and for NC block columns j .. j + NC - 1 alongside (the caller passes only rows
// i >= j + NC - 1, so every tile written exists): T(i, j + c) -= sum_k T(i, k) T(j + c, k)'.  Every A tile T(i, k) that crosses
// L2 -> SM then feeds the DMMAs of NC block columns: the left-looking factorisation of a large model is bound by exactly that traffic
// (11 MB per factorisation and CTA at n = 400, 148 CTAs at once).
template <int R, int NC>
__device__ __forceinline__ void tm_ll_update2(double *H, const TmLayout &L, int j, int i0, int st, int k0, int k1) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt = L.nt;
  const int fo = (lr << 3) + (lc ^ tm_swz(lr)), f1 = ((4 + lc) ^ tm_swz(lr)) - (lc ^ tm_swz(lr)), ao = (lr << 3) + ((2 * lc) ^ tm_swz(lr));
  const bool thin_ok = lr < L.last_h;
  const double *Brow[NC];
  int bst[NC];
  bool bv[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    Brow[c] = H + tm_tile(L, j + c, 0) + fo;
    bst[c] = (j + c == nt - 1) ? 8 * L.last_h : 64;
    bv[c] = j + c != nt - 1 || thin_ok;
  }
  const double *Arow[R];
  int ast[R];
  bool rv[R];
  double s0[R][NC][2], s1[R][NC][2];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * st;
    const int ii = min(i, nt - 1);
    rv[r] = i < nt && (ii != nt - 1 || thin_ok);
    Arow[r] = H + tm_tile(L, ii, 0) + fo;
    ast[r] = (ii == nt - 1) ? 8 * L.last_h : 64;
#pragma unroll
    for (int c = 0; c < NC; ++c) s0[r][c][0] = s0[r][c][1] = s1[r][c][0] = s1[r][c][1] = 0.0;
  }
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    double b0[NC], b1[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { b0[c] = bv[c] ? Brow[c][k * bst[c]] : 0.0; b1[c] = bv[c] ? Brow[c][k * bst[c] + f1] : 0.0; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double a0 = rv[r] ? Arow[r][k * ast[r]] : 0.0, a1 = rv[r] ? Arow[r][k * ast[r] + f1] : 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        dmma(s0[r][c][0], s0[r][c][1], a0, b0[c]);
        dmma(s1[r][c][0], s1[r][c][1], a1, b1[c]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (rv[r]) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double2 *C = reinterpret_cast<double2 *>(H + tm_tile(L, i0 + r * st, j + c) + ao);
        double2 h = *C;
        h.x -= s0[r][c][0] + s1[r][c][0]; h.y -= s0[r][c][1] + s1[r][c][1];
        *C = h;
      }
    }
  }
}

// ---- left-looking blocked Cholesky, same contract as chol_tm (rows [N, Nrows) carried, diagonal tiles end up as the
// inverses of the diagonal blocks of L).  Per block column j: warp 0 updates and factors the diagonal tile (the pivot
// chain), warps 1 .. NW-1 update the tiles below it; CTA barrier; every warp solves its tiles against the new diagonal
// block; CTA barrier.
__device__ __forceinline__ bool chol_tm_ll(double *H, const TmLayout &L, int N, int Nrows, int *fail, long long *dbg = nullptr) {
  const int w = threadIdx.x >> 5;
  const int nt = L.nt, ntN = (N + 7) >> 3;
  // (timestamps: the clock is read after a volatile shared-memory load, because BAR.SYNC.DEFER_BLOCKING lets a warp run on
  // past the barrier until its next memory instruction)
#define LL_STAMP(slot) do { if (dbg && (threadIdx.x & 31) == 0 && (w == 0 || w == IPM_NW - 1)) { const int f_ = *(volatile int *)fail; dbg[(slot) + (w ? 1 : 0)] = clock64() + (f_ & 0); } } while (0)
  __syncthreads();
  for (int j = 0; j < ntN; ++j) {
    if (j < 8) LL_STAMP(8 * j);
    // block columns in PAIRS: an even column applies its terms k < j and, alongside (same A tiles), the terms k < j of
    // column j + 1; the odd column that follows only has the term k = j left.  Measured at C5 / C4 (relaxations/s): one
    // column at a time 6.3 k / 1.80 k, pairs 7.2 k / 1.99 k, groups of four (one row per warp) 6.15 k / 1.77 k.
    const bool even = (j & 1) == 0;
    const int k0 = even ? 0 : j - 1, k1 = j;
    const bool two = even && j + 1 < ntN;
    if (w == 0) {
      if (k1 > k0) { tm_ll_update2<1, 1>(H, L, j, j, 0, k0, k1); __syncwarp(); }
      tm_potrf(H + tm_tile(L, j, j), min(8, N - 8 * j), min(8, Nrows - 8 * j), fail);
    } else if (k1 > k0) {
      for (int i = j + w; i < nt; i += 2 * (IPM_NW - 1)) {
        const bool pair = i + (IPM_NW - 1) < nt;
        if (two) { if (pair) tm_ll_update2<2, 2>(H, L, j, i, IPM_NW - 1, k0, k1); else tm_ll_update2<1, 2>(H, L, j, i, 0, k0, k1); }
        else { if (pair) tm_ll_update2<2, 1>(H, L, j, i, IPM_NW - 1, k0, k1); else tm_ll_update2<1, 1>(H, L, j, i, 0, k0, k1); }
      }
    }
    if (j < 8) LL_STAMP(8 * j + 2);
    __syncthreads();
    if (*fail) return false;
    if (j < 8) LL_STAMP(8 * j + 4);
    for (int i = j + 1 + w; i < nt; i += 3 * IPM_NW) tm_trsm_tiles3(H, L, j, i, IPM_NW);
    __syncthreads();
  }
#undef LL_STAMP
  return true;
}

// update step of the back-substitution for JN groups of 32 columns: v[j] -= sum_q L[i0 + q][k] x_q, k = lane + 32 j < i0.
// Straight-line (addresses clamped instead of branches), so the JN independent FMA chains interleave.
template <int JN>
__device__ __forceinline__ void bs_update(const double *H, const TmLayout &L, int bi, int i0, int nb, int lane, const double (&x)[8],
                                          double (&v)[16], double &keep) {
  const double *Trow = H + tm_tile(L, bi, 0);
  const int tstride = (bi == L.nt - 1) ? 8 * L.last_h : 64;
#pragma unroll
  for (int j0 = 0; j0 < JN; j0 += 4) {      // four columns per lane in flight (32 loads, then four independent FMA chains)
    double t[4][8];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (j0 + jj < JN) {
        const int k = min(lane + 32 * (j0 + jj), i0 - 1);
        const double *T = Trow + (k >> 3) * tstride;
        const int kc = k & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) t[jj][q] = (q < nb) ? T[(q << 3) + (kc ^ tm_swz(q))] : 0.0;
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (j0 + jj < JN) {
        double a = v[j0 + jj];
#pragma unroll
        for (int q = 0; q < 8; ++q) a = fma(-t[jj][q], x[q], a);
        const int k = lane + 32 * (j0 + jj);
        if (k < i0) {
          v[j0 + jj] = a;
          if (k >= i0 - 8) keep = a;
        }
      }
    }
  }
}

// ---- backward substitution L' x = w for the two right-hand sides carried in rows N and N + 1 (after chol_tm they hold
// L^-1 rhs).  The rows are copied to the contiguous vectors b0 / b1 (shared memory, >= 8 * ceil(N / 8) doubles each),
// which hold the solutions on exit.  One warp per right-hand side, no CTA barrier inside the block loop.
// GLOBAL_H: H lives in global memory — the tile rows the next two blocks will read are prefetched into L2 (a block's chain
// otherwise starts with a DRAM-latency load when the per-CTA blocks of a large model do not fit L2).
template <bool GLOBAL_H = false>
__device__ __forceinline__ void backsolve_tm(const double *H, const TmLayout &L, int N, double *b0, double *b1) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ntN = (N + 7) >> 3;
  for (int e = threadIdx.x; e < 2 * 8 * ntN; e += IPM_NT) {
    const int r = e / (8 * ntN), k = e - r * 8 * ntN;
    (r ? b1 : b0)[k] = (k < N) ? H[tm_idx(L, N + r, k)] : 0.0;
  }
  __syncthreads();
  if (w < 2) {
    // lane l keeps entries l, l + 32, ... of its right-hand side in registers; per block the eight solved entries go
    // through shared memory, everything else is independent FMA chains (nothing in the loop is stored to and re-read)
    double *b = w ? b1 : b0;
    constexpr int J = 16;                     // N <= 512
    double v[J];
#pragma unroll
    for (int j = 0; j < J; ++j) v[j] = (lane + 32 * j < 8 * ntN) ? b[lane + 32 * j] : 0.0;
    for (int bi = ntN - 1; bi >= 0; --bi) {
      const int i0 = bi << 3, nb = min(8, N - i0);
      const double *X = H + tm_tile(L, bi, bi);
      if (GLOBAL_H && w == 0 && bi >= 2) {
        const char *row = reinterpret_cast<const char *>(H + tm_tile(L, bi - 2, 0));
        const int bytes = ((bi - 1) * 64 + bi * 64) * 8;       // tile rows bi-2 and bi-1 are contiguous
        for (int off = lane * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
      }
      // b[i0 .. i0 + 7] is current in shared memory (initially, then written back by the previous block's update)
      double xc = 0.0;
      if (lane < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k >= lane && k < nb) xc = fma(X[(k << 3) + (lane ^ tm_swz(k))], b[i0 + k], xc);
      }
      __syncwarp();
      if (lane < 8) b[i0 + lane] = xc;        // zero beyond nb
      __syncwarp();
      double x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = b[i0 + q];
      double keep = 0.0;
      switch ((i0 + 31) >> 5) {     // 32-column groups that still hold unsolved entries (warp-uniform)
        case 1: bs_update<1>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 2: bs_update<2>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 3: bs_update<3>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 4: bs_update<4>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 5: bs_update<5>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 6: bs_update<6>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 7: bs_update<7>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 8: bs_update<8>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 9: bs_update<9>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 10: bs_update<10>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 11: bs_update<11>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 12: bs_update<12>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 13: bs_update<13>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 14: bs_update<14>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 15: bs_update<15>(H, L, bi, i0, nb, lane, x, v, keep); break;
        case 16: bs_update<16>(H, L, bi, i0, nb, lane, x, v, keep); break;
        default: break;
      }
      // the next block's entries go back to shared memory (one store, after the loop: a store inside it would order
      // every later load behind it)
      const int kst = i0 - 8 + ((lane - (i0 - 8)) & 31);
      if (i0 >= 8 && kst < i0) b[kst] = keep;
      __syncwarp();
    }
  }
  __syncthreads();
}

// =====================================================================================================================
// Back-substitution with 32 x 32 INVERSE diagonal blocks on the tile-major matrix (the large-model path: H in global memory).
//
// backsolve_tm above keeps one warp per right-hand side busy and everything else idle: at n = 400 its 51 block steps are 51
// chains of L2 round trips, 544 k cycles per iteration (23 % of the C5 iteration, profiles/r2c notes).  Here the diagonal
// 32 x 32 blocks made of full tiles are inverted in place first (one warp per block, all blocks at once, 8 x 8 tile products
// on DMMA: Inv_pq = -X_p sum_{r=q}^{p-1} L_pr Inv_rq, column by column), and a block step is then one dot-product stage
// (64 dot products, eight lanes each) and one update stage in which EVERY thread owns an entry of both right-hand sides
// and requests its 32 coefficients in batches of eight: 13 block steps at n = 400, all 512 threads loading.

// C += sgn * A B for 8 x 8 tiles of the tile-major matrix (a_lo / b_lo: the operand is lower triangular)
__device__ __forceinline__ void tm_tile_mm(const double *Ta, bool a_lo, const double *Tb, bool b_lo, double sgn, double &c0, double &c1) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int kk = 4 * ks + lc;
    const double av = (!a_lo || kk <= lr) ? Ta[(lr << 3) + (kk ^ tm_swz(lr))] : 0.0;      // A[m = lr][k = kk]
    const double bv = (!b_lo || lr <= kk) ? Tb[(kk << 3) + (lr ^ tm_swz(kk))] : 0.0;      // B[k = kk][n = lr]
    dmma(c0, c1, sgn * av, bv);
  }
}

// in-place inverse of the 32 x 32 diagonal block at tile row 4 b (four full tile rows; diagonal tiles hold X_p).  One warp.
__device__ __forceinline__ void tm_invert_block32(double *H, const TmLayout &L, int b) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  const int t0 = 4 * b;
  for (int q = 0; q < 3; ++q) {
    for (int p = q + 1; p < 4; ++p) {
      double s0 = 0.0, s1 = 0.0;
      for (int r = q; r < p; ++r)
        tm_tile_mm(H + tm_tile(L, t0 + p, t0 + r), false, H + tm_tile(L, t0 + r, t0 + q), r == q, 1.0, s0, s1);
      double *T = H + tm_tile(L, t0 + p, t0 + q);
      __syncwarp();
      *tm_acc(T, lr, lc) = make_double2(s0, s1);
      __syncwarp();
      double c0 = 0.0, c1 = 0.0;
      tm_tile_mm(H + tm_tile(L, t0 + p, t0 + p), true, T, false, -1.0, c0, c1);
      __syncwarp();
      *tm_acc(T, lr, lc) = make_double2(c0, c1);
      __syncwarp();
    }
  }
}

// L' x = w for the two right-hand sides carried in rows N and N + 1; b0 / b1 (shared memory, >= 8 ceil(N / 8) doubles each)
// hold the solutions on exit; xs: 64 doubles of shared scratch.  Same contract as backsolve_tm.
__device__ __forceinline__ void backsolve32_tm(double *H, const TmLayout &L, int N, double *b0, double *b1, double *xs) {
  const int tid = threadIdx.x, w = tid >> 5;
  const int ntN = (N + 7) >> 3, nfull = N >> 5;
  for (int b = w; b < nfull; b += IPM_NW) tm_invert_block32(H, L, b);
  for (int e = tid; e < 2 * 8 * ntN; e += IPM_NT) {
    const int r = e / (8 * ntN), k = e - r * 8 * ntN;
    (r ? b1 : b0)[k] = (k < N) ? H[tm_idx(L, N + r, k)] : 0.0;
  }
  __syncthreads();
  // ragged tail above the last full block: 8-row steps
  for (int bi = ntN - 1; bi >= 4 * nfull; --bi) {
    const int i0 = bi << 3, nb = min(8, N - i0);
    if (tid < 16) {
      const int rhs = tid >> 3, c = tid & 7;
      const double *bv = rhs ? b1 : b0;
      const double *X = H + tm_tile(L, bi, bi);
      double xc = 0.0;
      if (c < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nb && k >= c) xc = fma(X[(k << 3) + (c ^ tm_swz(k))], bv[i0 + k], xc);
      }
      xs[tid] = xc;
    }
    __syncthreads();
    for (int kk = tid; kk < i0 + 8; kk += IPM_NT) {
      if (kk < i0) {
        const double *T = H + tm_tile(L, bi, kk >> 3);
        const int kc = kk & 7;
        double v0 = b0[kk], v1 = b1[kk];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nb) {
            const double l = T[(c << 3) + (kc ^ tm_swz(c))];
            v0 = fma(-l, xs[c], v0);
            v1 = fma(-l, xs[8 + c], v1);
          }
        }
        b0[kk] = v0; b1[kk] = v1;
      } else if (kk < i0 + nb) {
        b0[kk] = xs[kk - i0]; b1[kk] = xs[8 + kk - i0];
      }
    }
    __syncthreads();
  }
  constexpr int LPD = IPM_NT / 64;                 // lanes per dot product (64 dot products: 32 entries x 2 right-hand sides)
  const int dot = tid / LPD, part = tid % LPD, drhs = dot >> 5, dc = dot & 31;
  for (int b = nfull - 1; b >= 0; --b) {
    const int i0 = 32 * b, t0 = 4 * b;
    {
      const double *wv = drhs ? b1 : b0;
      double a = 0.0;
#pragma unroll
      for (int u = 0; u < 32 / LPD; ++u) {
        const int k = part + LPD * u;                // x[c] = sum_{k >= c} Inv[k][c] w[k]
        if (k >= dc) {
          const double *T = H + tm_tile(L, t0 + (k >> 3), t0 + (dc >> 3));
          a = fma(T[((k & 7) << 3) + ((dc & 7) ^ tm_swz(k & 7))], wv[i0 + k], a);
        }
      }
#pragma unroll
      for (int o = LPD / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (part == 0) xs[dot] = a;
    }
    __syncthreads();
    for (int kk = tid; kk < i0 + 32; kk += IPM_NT) {
      if (kk < i0) {
        const int kt = kk >> 3, kc = kk & 7;
        double v0 = b0[kk], v1 = b1[kk];
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {        // eight coefficients requested before the first FMA
          const double *T = H + tm_tile(L, t0 + (c0 >> 3), kt);
          double l[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) l[c] = T[(c << 3) + (kc ^ tm_swz(c))];
#pragma unroll
          for (int c = 0; c < 8; ++c) { v0 = fma(-l[c], xs[c0 + c], v0); v1 = fma(-l[c], xs[32 + c0 + c], v1); }
        }
        b0[kk] = v0; b1[kk] = v1;
      } else {
        b0[kk] = xs[kk - i0]; b1[kk] = xs[32 + kk - i0];
      }
    }
    __syncthreads();
  }
}

}  // namespace BNLP_NS
