// single-warp dependent-chain latencies on sm_100a: DFMA, DMUL, rsqrt(double), 1/x, LDS(8B), DMMA m8n8k4, shfl f64
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *t, double seed, int nwarps_active) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  if ((threadIdx.x >> 5) >= nwarps_active) return;
  const int N = 256;
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001, acc = 0;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) t[0] = (t1 - t0); acc += x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) t[1] = (t1 - t0); acc += x;
  x = 2.0 + threadIdx.x * 1e-9;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) t[2] = (t1 - t0); acc += x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) t[3] = (t1 - t0); acc += x;
  int idx = threadIdx.x & 1023;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx];
  t1 = clock64(); if (threadIdx.x == 0) t[4] = (t1 - t0); acc += idx;
  double c0 = 0, c1 = 0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma(c0, c1, y, y);
  t1 = clock64(); if (threadIdx.x == 0) t[5] = (t1 - t0); acc += c0 + c1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) t[6] = (t1 - t0); acc += x;
  // independent DFMAs from one warp: issue rate
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) { a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9); a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9); }
  t1 = clock64(); if (threadIdx.x == 0) t[7] = (t1 - t0); acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) t[8] = (t1 - t0); acc += x;
  float xf = (float)x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) xf = rsqrtf(xf) + 1.5f;
  t1 = clock64(); if (threadIdx.x == 0) t[9] = (t1 - t0); acc += xf;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) t[10] = (t1 - t0);
  out[threadIdx.x] = acc;
}
int main() {
  double *o; long long *t, h[16];
  cudaMalloc(&o, 8 * 512); cudaMalloc(&t, 128);
  const char *nm[11] = {"DFMA dep", "DMUL dep", "rsqrt(double)+add dep", "1/x+add dep", "LDS.64 dep (+cvt)", "DMMA dep", "shfl f64 + add dep", "8 indep DFMA (per 8)", "sqrt(double)+add dep", "rsqrtf+add dep", "__syncthreads"};
  for (int threads : {32, 512}) {
    for (int rep = 0; rep < 2; ++rep) { k<<<1, threads>>>(o, t, 1.0, 16); cudaDeviceSynchronize(); }
    cudaMemcpy(h, t, 128, cudaMemcpyDeviceToHost);
    printf("threads=%d\n", threads);
    for (int i = 0; i < 11; ++i) printf("  %-24s %.1f cycles\n", nm[i], h[i] / 256.0);
  }
  return 0;
}
