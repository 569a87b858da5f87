// The cooperative IPM kernel built a second time with 256 threads per CTA, for models small enough that two CTAs fit
// one SM (two children per SM interleave their latency chains).  The device headers are compiled into namespace
// bnlp256 (BNLP_NS); the argument structs have the same layout as in namespace bnlp.
#include <cuda_runtime.h>

#include "../../include/bnlp.h"
#define BNLP_NT 256
#define BNLP_NS bnlp256
#include "bnlp_internal.h"
namespace bnlp256 {
constexpr double kInfD = 1e19;
using bnlp::HostModel;
}
#include "bnlp_ipm_coop.cuh"

cudaError_t bnlp_coop256_set_smem(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(bnlp256::ipm_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bnlp256::ipm_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t bnlp_coop256_launch(bool timers, int ctas, size_t smem, cudaStream_t stream, void *M, void *A, void *CB) {
  void *args[] = {M, A, CB};
  return cudaLaunchCooperativeKernel(timers ? (void *)bnlp256::ipm_coop_kernel<true> : (void *)bnlp256::ipm_coop_kernel<false>,
                                     dim3(ctas), dim3(256), args, smem, stream);
}
int bnlp_coop256_max_ctas_per_sm(size_t smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bnlp256::ipm_coop_kernel<false>, 256, smem) != cudaSuccess) return 0;
  return nb;
}
