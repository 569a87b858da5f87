// micro-benchmark: FP64 mma.sync m8n8k4 (DMMA) latency / throughput vs DFMA on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k_dmma(double *out, int iters, long long *cyc) {
  double c[CH][2];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void k_dfma(double *out, int iters, long long *cyc) {
  double c[CH];
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-9;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double *out, int iters, long long *cyc) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rsqrt(double *out, int iters, long long *cyc) {
  double v = 2.0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = rsqrt(v) + 2.0;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int it = 4096;
#define RUN(name, kern, blocks, threads, per_iter)                                                       \
  kern<<<blocks, threads>>>(out, it, cyc); cudaDeviceSynchronize();                                      \
  kern<<<blocks, threads>>>(out, it, cyc); cudaDeviceSynchronize();                                      \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                        \
  printf("%-34s blocks=%3d thr=%4d : %.2f cycles per loop iteration (%d ops/iter/warp)\n", name, blocks, threads, (double)h / it, per_iter);
  RUN("dmma 1 chain (latency)", k_dmma<1>, 1, 32, 1)
  RUN("dmma 4 chains", k_dmma<4>, 1, 32, 4)
  RUN("dmma 8 chains 1 warp", k_dmma<8>, 1, 32, 8)
  RUN("dmma 8 chains 4 warps", k_dmma<8>, 1, 128, 8)
  RUN("dmma 8 chains 16 warps", k_dmma<8>, 1, 512, 8)
  RUN("dmma 8 chains 16 warps x148", k_dmma<8>, 148, 512, 8)
  RUN("dfma 1 chain (latency)", k_dfma<1>, 1, 32, 1)
  RUN("dfma 8 chains 1 warp", k_dfma<8>, 1, 32, 8)
  RUN("dfma 8 chains 16 warps", k_dfma<8>, 1, 512, 8)
  RUN("dfma 8 chains 16 warps x148", k_dfma<8>, 148, 512, 8)
  RUN("shfl f64 dependent", k_shfl, 1, 32, 1)
  RUN("rsqrt f64 dependent", k_rsqrt, 1, 32, 1)
  return 0;
}
