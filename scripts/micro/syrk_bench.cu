#include <cstdio>
#include <cuda_runtime.h>
namespace bnlp { constexpr double kInfD = 1e19; }
#include "../../bonmin_b200/csrc/bnlp_ipm_coop.cuh"
using namespace bnlp;
__global__ void __launch_bounds__(512, 1) bench(int N, long long *out) {
  extern __shared__ double sm[];
  const int lds = ((N + 2 + 7) & ~7) + 4;
  double *H = sm, *stage = sm + ((tri(N + 2) + 1) & ~1);
  for (int e = threadIdx.x; e < tri(N + 2); e += 512) H[e] = 0.0;
  for (int e = threadIdx.x; e < 17 * lds; e += 512) stage[e] = 1e-3 * (e % 13);
  __syncthreads();
  int idx = 0;
  for (int K = 4; K <= 16; K *= 2) {
    for (int lo = 0; lo <= 160; lo += 80) {
      __syncthreads();
      long long t0 = clock64();
      syrk_dmma(H, lo, N, stage, lds, K, 1.0);
      __syncthreads();
      long long t1 = clock64();
      if (threadIdx.x == 0) out[idx] = t1 - t0;
      idx++;
    }
  }
  // empty call (barrier only) and a single-tile call
  __syncthreads(); long long t0 = clock64(); __syncthreads(); long long t1 = clock64();
  if (threadIdx.x == 0) out[idx++] = t1 - t0;
  __syncthreads(); t0 = clock64(); syrk_dmma(H, 192, N, stage, lds, 8, 1.0); __syncthreads(); t1 = clock64();
  if (threadIdx.x == 0) out[idx++] = t1 - t0;
  __syncthreads(); t0 = clock64(); syrk_dmma(H, 8, N, stage, lds, 8, 1.0, 1, 1); __syncthreads(); t1 = clock64();
  if (threadIdx.x == 0) out[idx++] = t1 - t0;
}
int main() {
  const int N = 201;
  long long *d, h[16];
  cudaMalloc(&d, 128);
  size_t smem = (size_t)(((N + 2) * (N + 3) / 2 + 2) + 17 * 216) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) { bench<<<1, 512, smem>>>(N, d); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  int i = 0;
  for (int K = 4; K <= 16; K *= 2) for (int lo = 0; lo <= 160; lo += 80) { printf("syrk K=%2d lo=%3d : %lld cycles\n", K, lo, h[i]); i++; }
  printf("barrier only: %lld ; syrk lo=192 (3 tiles) K=8: %lld ; one block column (25 tiles) K=8: %lld\n", h[i], h[i + 1], h[i + 2]);
  return 0;
}
