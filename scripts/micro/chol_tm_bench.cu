// Micro-benchmark + correctness check of the tile-major dense kernels (bnlp_dense_tm.cuh) against the packed ones
// (bnlp_dense.cuh) and a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o chol_tm_bench.bin chol_tm_bench.cu
//   ./chol_tm_bench.bin [N] [CTAs]
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
namespace bnlp { constexpr double kInfD = 1e19; }
#include "../../bonmin_b200/csrc/bnlp_ipm_coop.cuh"
#include "../../bonmin_b200/csrc/bnlp_dense_tm.cuh"
using namespace bnlp;

__global__ void __launch_bounds__(IPM_NT, 1) bench_tm(int N, const double *H0, const double *M0, double *Hout, double *Xout,
                                                      double *Sout, long long *out, int usedbg) {
  extern __shared__ double sm[];
  const int Nrows = N + 2;
  const TmLayout L = tm_layout(Nrows);
  const int lds = ((N + 2 + 7) & ~7) + 4;
  const int P = tm_size(L);
  double *H = sm, *stage = sm + P;
  __shared__ int fail;
  __shared__ int ctr[2];
  __shared__ long long dbg[112];
  if (threadIdx.x < 112) dbg[threadIdx.x] = 0;
  for (int e = threadIdx.x; e < P; e += IPM_NT) H[e] = H0[e];
  for (int e = threadIdx.x; e < 16 * lds; e += IPM_NT) stage[e] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  const bool ok = (usedbg >= 2) ? chol_tm_ll(H, L, N, Nrows, &fail, usedbg == 3 ? dbg : nullptr) : chol_tm(H, L, N, Nrows, &fail, ctr, usedbg ? dbg : nullptr);
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0) for (int e = threadIdx.x; e < P; e += IPM_NT) Hout[e] = H[e];
  __syncthreads();
  long long t2 = clock64();
  backsolve_tm(H, L, N, stage, stage + lds);
  __syncthreads();
  long long t3 = clock64();
  if (blockIdx.x == 0) for (int e = threadIdx.x; e < 2 * N; e += IPM_NT) Xout[e] = stage[(e / N) * lds + e % N];
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = t3 - t2; for (int q = 0; q < 8; ++q) out[8 + q] = dbg[q]; for (int q = 0; q < 56; ++q) out[32 + q] = dbg[16 + q] - t0; for (int q = 0; q < 32; ++q) out[88 + q] = dbg[80 + q] - t0; if (usedbg == 3) for (int q = 0; q < 64; ++q) out[32 + q] = dbg[q] - t0; }
  __syncthreads();
  // J' Sigma J from 16 staged rows: first chunk stores, four more accumulate
  for (int e = threadIdx.x; e < 16 * lds; e += IPM_NT) stage[e] = M0[e];
  if (threadIdx.x == 0) ctr[0] = 0;
  __syncthreads();
  const int ntr = (N + 7) >> 3;
  long long s0 = clock64();
  tm_syrk_stage<true>(H, L, ntr, stage, lds, 16, ctr);
  __syncthreads();
  long long s1 = clock64();
  for (int rep = 0; rep < 4; ++rep) {
    if (threadIdx.x == 0) ctr[0] = 0;
    __syncthreads();
    tm_syrk_stage<false>(H, L, ntr, stage, lds, 16, ctr);
    __syncthreads();
  }
  long long s2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[3] = s1 - s0; out[4] = (s2 - s1) / 4; }
  if (blockIdx.x == 0) for (int e = threadIdx.x; e < P; e += IPM_NT) Sout[e] = H[e];
  // primitive probes (warp 0 only where it says so)
  __syncthreads();
  for (int e = threadIdx.x; e < P; e += IPM_NT) H[e] = H0[e];
  __syncthreads();
  long long q0 = clock64();
  if (threadIdx.x < 32) tm_potrf(H + tm_tile(L, 0, 0), 8, 8, &fail, dbg);
  long long q1 = clock64();
  if (threadIdx.x < 32) tm_trsm_tiles3(H, L, 0, 1, 3);
  long long q2 = clock64();
  if (threadIdx.x < 32) tm_trail_unit(H, L, 0, 6, 2);
  long long q3 = clock64();
  int u = 0;
  if (threadIdx.x == 0) u = atomicAdd(ctr, 1);
  u = __shfl_sync(0xffffffffu, u, 0);
  long long q4 = clock64();
  if (threadIdx.x < 32) tm_colupd_tile(H, L, 0, 1);
  long long q5 = clock64();
  if (threadIdx.x < 128) named_barrier(1, 128);
  long long q6 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[16] = q1 - q0; out[17] = q2 - q1; out[18] = q3 - q2; out[19] = q4 - q3 + (u & 0); out[20] = q5 - q4; out[21] = q6 - q5; for (int q = 0; q < 4; ++q) out[22 + q] = dbg[q]; }
}

// the packed kernels of round 1, same inputs
__global__ void __launch_bounds__(IPM_NT, 1) bench_packed(int N, const double *H0, long long *out) {
  extern __shared__ double sm[];
  const int Nrows = N + 2;
  const int lds = ((N + 2 + 7) & ~7) + 4;
  const int P = tri(Nrows);
  double *H = sm, *stage = sm + ((P + 1) & ~1);
  __shared__ int fail;
  __shared__ double xs[16];
  for (int e = threadIdx.x; e < P; e += IPM_NT) H[e] = H0[e];
  for (int e = threadIdx.x; e < 17 * lds; e += IPM_NT) stage[e] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  const bool ok = chol3(H, N, Nrows, stage, lds, &fail, nullptr);
  __syncthreads();
  long long t1 = clock64();
  backsolve_inv(H, N, xs);
  __syncthreads();
  long long t2 = clock64();
  for (int e = threadIdx.x; e < 17 * lds; e += IPM_NT) stage[e] = 1e-3 * (e % 13);
  __syncthreads();
  long long t3 = clock64();
  syrk_dmma(H, 0, N, stage, lds, 16, 1.0);
  __syncthreads();
  long long t4 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = t2 - t1; out[3] = t4 - t3; }
}

// only the left-looking factorisation + back-substitution (small code: is the big kernel paying for instruction fetch?)
__global__ void __launch_bounds__(IPM_NT, 1) bench_ll(int N, const double *H0, long long *out) {
  extern __shared__ double sm[];
  const int Nrows = N + 2;
  const TmLayout L = tm_layout(Nrows);
  const int lds = ((N + 2 + 7) & ~7) + 4;
  const int P = tm_size(L);
  double *H = sm, *stage = sm + P;
  __shared__ int fail;
  for (int rep = 0; rep < 4; ++rep) {
    for (int e = threadIdx.x; e < P; e += IPM_NT) H[e] = H0[e];
    __syncthreads();
    long long t0 = clock64();
    const bool ok = chol_tm_ll(H, L, N, Nrows, &fail);
    __syncthreads();
    long long t1 = clock64();
    backsolve_tm(H, L, N, stage, stage + lds);
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = t2 - t1; }
  }
}

// the same routines with H in GLOBAL memory (one block of P doubles per CTA, L2/L1 resident): the large-n path
__global__ void __launch_bounds__(IPM_NT, 1) bench_ll_global(int N, const double *H0, double *Hg, const double *M0, double *Hout,
                                                             double *Xout, long long *out) {
  extern __shared__ double sm[];
  const int Nrows = N + 2;
  const TmLayout L = tm_layout(Nrows);
  const int lds = ((N + 2 + 7) & ~7) + 4;
  const int P = tm_size(L);
  double *H = Hg + (size_t)blockIdx.x * P, *stage = sm;
  __shared__ int fail;
  __shared__ int ctr[2];
  for (int rep = 0; rep < 2; ++rep) {
    for (int e = threadIdx.x; e < P; e += IPM_NT) H[e] = H0[e];
    __syncthreads();
    long long t0 = clock64();
    const bool ok = chol_tm_ll(H, L, N, Nrows, &fail);
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && rep == 1) for (int e = threadIdx.x; e < P; e += IPM_NT) Hout[e] = H[e];
    __syncthreads();
    long long t2 = clock64();
    backsolve_tm(H, L, N, stage, stage + lds);
    __syncthreads();
    long long t3 = clock64();
    if (blockIdx.x == 0 && rep == 1) for (int e = threadIdx.x; e < 2 * N; e += IPM_NT) Xout[e] = stage[(e / N) * lds + e % N];
    __syncthreads();
    for (int e = threadIdx.x; e < 16 * lds; e += IPM_NT) stage[e] = M0[e];
    if (threadIdx.x == 0) ctr[0] = 0;
    __syncthreads();
    long long t4 = clock64();
    tm_syrk_stage<false>(H, L, (N + 7) >> 3, stage, lds, 16, ctr);
    __syncthreads();
    long long t5 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = t3 - t2; out[3] = t5 - t4; }
  }
}

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 201;
  const int NB = argc > 2 ? atoi(argv[2]) : 1;
  const int Nrows = N + 2, Pp = Nrows * (Nrows + 1) / 2;
  const TmLayout L = tm_layout(Nrows);
  const int P = tm_size(L);
  const int lds = ((N + 2 + 7) & ~7) + 4;
  std::vector<double> G((size_t)N * N), A((size_t)Nrows * Nrows, 0.0), Hp(Pp), Ht(P, 0.0);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0 - 0.5; };
  for (auto &g : G) g = rnd();
  for (int i = 0; i < N; ++i)
    for (int k = 0; k <= i; ++k) {
      double a = (i == k) ? 0.5 * N : 0.0;
      for (int c = 0; c < N; ++c) a += G[(size_t)i * N + c] * G[(size_t)k * N + c];
      A[(size_t)i * Nrows + k] = a;
    }
  for (int r = N; r < Nrows; ++r)
    for (int k = 0; k < N; ++k) A[(size_t)r * Nrows + k] = rnd();
  for (int i = 0; i < Nrows; ++i)
    for (int k = 0; k <= i; ++k) {
      Hp[i * (i + 1) / 2 + k] = A[(size_t)i * Nrows + k];
      Ht[tm_idx(L, i, k)] = A[(size_t)i * Nrows + k];
    }
  std::vector<double> Lh = A;
  for (int j = 0; j < N; ++j) {
    double d = Lh[(size_t)j * Nrows + j];
    for (int c = 0; c < j; ++c) d -= Lh[(size_t)j * Nrows + c] * Lh[(size_t)j * Nrows + c];
    d = std::sqrt(d);
    Lh[(size_t)j * Nrows + j] = d;
    for (int i = j + 1; i < Nrows; ++i) {
      double v = Lh[(size_t)i * Nrows + j];
      for (int c = 0; c < j; ++c) v -= Lh[(size_t)i * Nrows + c] * Lh[(size_t)j * Nrows + c];
      Lh[(size_t)i * Nrows + j] = v / d;
    }
  }
  std::vector<double> X(2 * N);
  for (int r = 0; r < 2; ++r)
    for (int i = N - 1; i >= 0; --i) {
      double v = Lh[(size_t)(N + r) * Nrows + i];
      for (int k = i + 1; k < N; ++k) v -= Lh[(size_t)k * Nrows + i] * X[r * N + k];
      X[r * N + i] = v / Lh[(size_t)i * Nrows + i];
    }
  // staged rows for the syrk test: 16 x lds, zero beyond N
  std::vector<double> M0((size_t)16 * lds, 0.0);
  for (int c = 0; c < 16; ++c)
    for (int i = 0; i < N; ++i) M0[(size_t)c * lds + i] = rnd();
  double *dH0, *dHt, *dM0, *dHo, *dXo, *dSo; long long *dout;
  cudaMalloc(&dH0, Pp * 8); cudaMalloc(&dHt, P * 8); cudaMalloc(&dM0, M0.size() * 8); cudaMalloc(&dHo, P * 8);
  cudaMalloc(&dXo, 2 * N * 8); cudaMalloc(&dSo, P * 8); cudaMalloc(&dout, 1024);
  cudaMemcpy(dH0, Hp.data(), Pp * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dHt, Ht.data(), P * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dM0, M0.data(), M0.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem_tm = (size_t)(P + 16 * lds) * 8, smem_p = (size_t)(((Pp + 1) & ~1) + 17 * lds) * 8;
  cudaFuncSetAttribute(bench_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tm);
  cudaFuncSetAttribute(bench_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
  if (argc > 3 && argv[3][0] == 'g') {
    double *Hg; cudaMalloc(&Hg, (size_t)NB * P * 8);
    const size_t smem_g = (size_t)17 * lds * 8;
    cudaFuncSetAttribute(bench_ll_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g);
    long long h2[8];
    for (int rep = 0; rep < 2; ++rep) { bench_ll_global<<<NB, IPM_NT, smem_g>>>(N, dHt, Hg, dM0, dHo, dXo, dout); cudaDeviceSynchronize(); }
    cudaMemcpy(h2, dout, 64, cudaMemcpyDeviceToHost);
    std::vector<double> Ho(P), Xo(2 * N);
    cudaMemcpy(Ho.data(), dHo, P * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Xo.data(), dXo, 2 * N * 8, cudaMemcpyDeviceToHost);
    double errl = 0, errx = 0;
    for (int i = 0; i < Nrows; ++i)
      for (int k = 0; k < (i & ~7) && k < N; ++k) errl = fmax(errl, fabs(Ho[tm_idx(L, i, k)] - Lh[(size_t)i * Nrows + k]));
    for (int r = 0; r < 2; ++r)
      for (int i = 0; i < N; ++i) errx = fmax(errx, fabs(Xo[r * N + i] - X[r * N + i]) / (1e-300 + fabs(X[r * N + i])));
    printf("H in GLOBAL memory, N=%d, %d CTAs: chol_tm_ll %lld cycles (ok=%lld), backsolve_tm %lld, rank-16 syrk (add) %lld | max|L err| %.2e max rel x err %.2e (%s)\n",
           N, NB, h2[0], h2[1], h2[2], h2[3], errl, errx, cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (argc > 3) {
    cudaFuncSetAttribute(bench_ll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tm);
    long long h2[8];
    for (int rep = 0; rep < 2; ++rep) { bench_ll<<<NB, IPM_NT, smem_tm>>>(N, dHt, dout); cudaDeviceSynchronize(); }
    cudaMemcpy(h2, dout, 64, cudaMemcpyDeviceToHost);
    printf("bench_ll alone: chol_tm_ll %lld cycles (ok=%lld), backsolve_tm %lld (%s)\n", h2[0], h2[1], h2[2], cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  printf("N=%d tile rows %d (last %d high): tile-major %d doubles (%.1f KB), packed %d doubles (%.1f KB); %d CTAs\n", N, L.nt,
         L.last_h, P, P * 8 / 1024.0, Pp, Pp * 8 / 1024.0, NB);
  long long h[128];
  for (int rep = 0; rep < 3; ++rep) { bench_packed<<<NB, IPM_NT, smem_p>>>(N, dH0, dout); cudaDeviceSynchronize(); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error (packed): %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, dout, 256, cudaMemcpyDeviceToHost);
  printf("packed (round 1): chol3 %lld cycles (ok=%lld), backsolve %lld, rank-16 syrk %lld\n", h[0], h[1], h[2], h[3]);
  for (int rep = 0; rep < 3; ++rep) { bench_tm<<<NB, IPM_NT, smem_tm>>>(N, dHt, dM0, dHo, dXo, dSo, dout, 1); cudaDeviceSynchronize(); }
  cudaMemcpy(h, dout, 256, cudaMemcpyDeviceToHost);
  printf("  (timed build) chol_tm %lld | warp 0: diag update+factor %lld, panel barrier %lld, panel solve %lld, trailing %lld, CTA barrier %lld | last warp: trailing %lld, CTA barrier %lld, trsm %lld, barrier 2 %lld\n",
         h[0], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[11], h[15]);
  { long long hh[128]; cudaMemcpy(hh, dout, 1024, cudaMemcpyDeviceToHost); printf("  barrier exits (cycles since start; after phase 1 / after phase 2):"); for (int q = 0; q < 25; ++q) printf(" %lld/%lld", hh[32 + 2 * q], hh[33 + 2 * q]); printf("\n  trsm phase of panel 2, per warp (start-end):"); for (int q = 0; q < 16; ++q) printf(" %lld-%lld", hh[88 + 2 * q], hh[89 + 2 * q]); printf("\n"); }
  printf("  probes (warp 0 alone): potrf %lld | trsm_tiles3 %lld | one 8-tile trailing unit %lld | counter fetch %lld | colupd tile %lld | 4-warp barrier %lld || potrf: load %lld factor %lld store+sync %lld inverse %lld\n", h[16], h[17], h[18], h[19], h[20], h[21], h[22], h[23], h[24], h[25]);
  for (int rep = 0; rep < 3; ++rep) { bench_tm<<<NB, IPM_NT, smem_tm>>>(N, dHt, dM0, dHo, dXo, dSo, dout, 2); cudaDeviceSynchronize(); }
  { bench_tm<<<NB, IPM_NT, smem_tm>>>(N, dHt, dM0, dHo, dXo, dSo, dout, 3); cudaDeviceSynchronize();
    long long hh[128]; cudaMemcpy(hh, dout, 1024, cudaMemcpyDeviceToHost);
    printf("  left-looking, block columns 0-7 (cycles since start) [warp0: start, potrf start, potrf end, after barrier | last warp: start, update end, after barrier]\n");
    for (int j = 0; j < 8; ++j) printf("    j=%d  w0 %lld %lld %lld %lld | w15 %lld %lld %lld\n", j, hh[32 + 8 * j], hh[32 + 8 * j + 6], hh[32 + 8 * j + 2], hh[32 + 8 * j + 4], hh[32 + 8 * j + 1], hh[32 + 8 * j + 3], hh[32 + 8 * j + 5]);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error (tile-major): %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, dout, 256, cudaMemcpyDeviceToHost);
  std::vector<double> Ho(P), Xo(2 * N), So(P);
  cudaMemcpy(Ho.data(), dHo, P * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(Xo.data(), dXo, 2 * N * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(So.data(), dSo, P * 8, cudaMemcpyDeviceToHost);
  double errl = 0, errx = 0, errs = 0;
  for (int i = 0; i < Nrows; ++i)
    for (int k = 0; k < (i & ~7) && k < N; ++k) errl = fmax(errl, fabs(Ho[tm_idx(L, i, k)] - Lh[(size_t)i * Nrows + k]));
  for (int r = 0; r < 2; ++r)
    for (int i = 0; i < N; ++i) errx = fmax(errx, fabs(Xo[r * N + i] - X[r * N + i]) / (1e-300 + fabs(X[r * N + i])));
  for (int i = 0; i < N; ++i)
    for (int k = 0; k <= i; ++k) {
      double ref = 0;
      for (int c = 0; c < 16; ++c) ref += M0[(size_t)c * lds + i] * M0[(size_t)c * lds + k];
      errs = fmax(errs, fabs(So[tm_idx(L, i, k)] - 5.0 * ref));
    }
  printf("tile-major LEFT-LOOKING: chol_tm_ll %lld cycles (ok=%lld), backsolve %lld, rank-16 syrk store %lld / add %lld | max|L err| %.2e, max rel x err %.2e, max|syrk err| %.2e\n",
         h[0], h[1], h[2], h[3], h[4], errl, errx, errs);
  return (errl < 1e-10 && errx < 1e-9 && errs < 1e-10 && h[1] == 1) ? 0 : 2;
}
