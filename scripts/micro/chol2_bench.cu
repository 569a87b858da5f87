// Micro-benchmark + correctness check of the per-child dense kernels (bnlp_dense.cuh) against the previous versions
// and a host reference: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o chol2_bench chol2_bench.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
namespace bnlp { constexpr double kInfD = 1e19; }
#include "../../bonmin_b200/csrc/bnlp_ipm_coop.cuh"
using namespace bnlp;

__global__ void __launch_bounds__(512, 1) bench(int N, const double *H0, double *Hout, long long *out, int which, int usedbg) {
  extern __shared__ double sm[];
  const int Nrows = N + 2;
  const int lds = ((N + 2 + 7) & ~7) + 4;
  const int P = tri(Nrows);
  double *H = sm, *stage = sm + ((P + 1) & ~1);
  __shared__ int fail;
  __shared__ double xs[16];
  __shared__ double xs64[64];
  for (int e = threadIdx.x; e < P; e += 512) H[e] = H0[e];
  for (int e = threadIdx.x; e < 17 * lds; e += 512) stage[e] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  bool ok = true;
  if (which == 0) ok = chol_dmma(H, N, Nrows, stage, lds, &fail, nullptr);
  long long dbgl[5] = {0, 0, 0, 0, 0};
  long long *dbg = usedbg ? dbgl : nullptr;
  if (which == 1) ok = chol2<false>(H, N, Nrows, stage, lds, &fail, dbg);
  else if (which == 2) ok = chol2<true>(H, N, Nrows, stage, lds, &fail, dbg);
  else if (which == 3) ok = chol3(H, N, Nrows, stage, lds, &fail, dbg);
  else if (which == 4) ok = chol4(H, N, Nrows, stage, lds, &fail, dbg);
  if (threadIdx.x == 0 && blockIdx.x == 0) for (int q = 0; q < 5; ++q) out[8 + q] = dbgl[q];
  __syncthreads();
  long long t1 = clock64();
  if (which == 4) backsolve32(H, N, xs64); else backsolve_inv(H, N, xs);
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = t2 - t1; }
  if (blockIdx.x == 0) for (int e = threadIdx.x; e < P; e += 512) Hout[e] = H[e];
  // syrk timings (old vs new), full rank-16 update
  for (int e = threadIdx.x; e < 17 * lds; e += 512) stage[e] = 1e-3 * (e % 13);
  __syncthreads();
  t0 = clock64();
  syrk_dmma(H, 0, N, stage, lds, 16, 1.0);
  __syncthreads();
  t1 = clock64();
  syrk_rt(H, 0, N, stage, lds, 16, 1.0, threadIdx.x >> 5, IPM_NW);
  __syncthreads();
  t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[3] = t1 - t0; out[4] = t2 - t1; }
  // small-phase probes
  for (int e = threadIdx.x; e < P; e += 512) H[e] = H0[e];
  __syncthreads();
  long long q0 = clock64(); __syncthreads(); long long q1 = clock64();
  if (threadIdx.x < 32) chol_diag8_inv(H, 0, 8, &fail);
  long long q2 = clock64(); __syncthreads(); long long q3 = clock64();
  trsm_dmma8(H, 0, 8, 8, Nrows, stage, lds);
  long long q4 = clock64(); __syncthreads(); long long q5 = clock64();
  syrk_col8(H, 8, Nrows, stage, lds, threadIdx.x >> 5, IPM_NW);
  long long q6 = clock64(); __syncthreads(); long long q7 = clock64();
  trsm_rows8(H, 8, 8, 16, Nrows, stage + 8 * lds, lds);
  long long q8 = clock64(); __syncthreads(); long long q9 = clock64();
  // chol4 pieces, one warp alone: diagonal tile by a whole warp, rest of the 16 x 16 block, warp 0's trailing unit
  for (int e = threadIdx.x; e < P; e += 512) H[e] = H0[e];
  __syncthreads();
  long long r0 = clock64();
  if (threadIdx.x < 32) chol_diag8_invw(H, 0, &fail);
  long long r1 = clock64();
  if (threadIdx.x < 32) chol_diag16_rest(H, 0, &fail);
  long long r2 = clock64();
  if (threadIdx.x < 32) syrk_rt(H, 16, Nrows, stage, lds, 16, -1.0, 0, 1 << 20);
  long long r3 = clock64();
  __syncthreads();
  long long r4 = clock64();
  panel16_rows(H, 0, Nrows, stage, lds);
  long long r5 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[19] = r1 - r0; out[20] = r2 - r1; out[21] = r3 - r2; out[22] = r5 - r4; }
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[5] = q1 - q0; out[6] = q2 - q1; out[7] = q3 - q2; out[13] = q4 - q3; out[14] = q5 - q4; out[15] = q6 - q5; out[16] = q7 - q6; out[17] = q8 - q7; out[18] = q9 - q8; }
}

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 201;
  const int NB = argc > 2 ? atoi(argv[2]) : 1;
  const int Nrows = N + 2, P = Nrows * (Nrows + 1) / 2;
  std::vector<double> G((size_t)N * N), A((size_t)Nrows * Nrows, 0.0), Hp(P);
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
    for (int k = 0; k <= i; ++k) Hp[i * (i + 1) / 2 + k] = A[(size_t)i * Nrows + k];
  // host reference: L and forward-substituted rhs rows, then back-substitution
  std::vector<double> L = A;
  for (int j = 0; j < N; ++j) {
    double d = L[(size_t)j * Nrows + j];
    for (int c = 0; c < j; ++c) d -= L[(size_t)j * Nrows + c] * L[(size_t)j * Nrows + c];
    d = std::sqrt(d);
    L[(size_t)j * Nrows + j] = d;
    for (int i = j + 1; i < Nrows; ++i) {
      double v = L[(size_t)i * Nrows + j];
      for (int c = 0; c < j; ++c) v -= L[(size_t)i * Nrows + c] * L[(size_t)j * Nrows + c];
      L[(size_t)i * Nrows + j] = v / d;
    }
  }
  std::vector<double> X(2 * N);
  for (int r = 0; r < 2; ++r)
    for (int i = N - 1; i >= 0; --i) {
      double v = L[(size_t)(N + r) * Nrows + i];
      for (int k = i + 1; k < N; ++k) v -= L[(size_t)k * Nrows + i] * X[r * N + k];
      X[r * N + i] = v / L[(size_t)i * Nrows + i];
    }
  double *dH0, *dHo; long long *dout;
  cudaMalloc(&dH0, P * 8); cudaMalloc(&dHo, P * 8); cudaMalloc(&dout, 256);
  cudaMemcpy(dH0, Hp.data(), P * 8, cudaMemcpyHostToDevice);
  const int lds = ((N + 2 + 7) & ~7) + 4;
  size_t smem = (size_t)(((P + 1) & ~1) + 17 * lds) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char *nm[5] = {"chol_dmma (old)", "chol2", "chol2 look-ahead", "chol3", "chol4"};
  for (int usedbg = 0; usedbg < 2; ++usedbg)
  for (int which = 0; which < 5; ++which) {
    long long h[32];
    for (int rep = 0; rep < 3; ++rep) { bench<<<NB, 512, smem>>>(N, dH0, dHo, dout, which, usedbg); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dout, 256, cudaMemcpyDeviceToHost);
    std::vector<double> Ho(P);
    cudaMemcpy(Ho.data(), dHo, P * 8, cudaMemcpyDeviceToHost);
    // after backsolve the rhs rows hold the solutions x
    double errx = 0, errl = 0;
    for (int r = 0; r < 2; ++r)
      for (int i = 0; i < N; ++i) errx = fmax(errx, fabs(Ho[(N + r) * (N + r + 1) / 2 + i] - X[r * N + i]) / (1e-300 + fabs(X[r * N + i])));
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < (i & ~7); ++k) errl = fmax(errl, fabs(Ho[i * (i + 1) / 2 + k] - L[(size_t)i * Nrows + k]));
    printf("%-18s N=%d: factor %lld cycles (ok=%lld), backsolve %lld cycles, max|L err| %.2e, max rel x err %.2e | syrk16 old %lld new %lld\n",
           nm[which], N, h[0], h[1], h[2], errl, errx, h[3], h[4]);
    if (!which && !usedbg) printf("  probes (thread 0): sync %lld | diag8_inv %lld +sync %lld | trsm_dmma8 %lld +sync %lld | col8 %lld +sync %lld | trsm_rows8 %lld +sync %lld\n", h[5], h[6], h[7], h[13], h[14], h[15], h[16], h[17], h[18]);
    if (!which && !usedbg) printf("  chol4 probes (warp 0 alone): diag8_invw %lld | diag16_rest (incl. second diag8_invw) %lld | trailing unit 0 %lld | panel16_rows (thread 0) %lld\n", h[19], h[20], h[21], h[22]);
    if (which) printf("    diag %lld trsm %lld col8 %lld trailing %lld invert %lld\n", h[8], h[9], h[10], h[11], h[12]);
  }
  return 0;
}
