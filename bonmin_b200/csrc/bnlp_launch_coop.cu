// Translation unit of the cooperative IPM kernel (bnlp_ipm_coop.cuh): instantiations + launchers.
#include <cuda_runtime.h>

#include "../../include/bnlp.h"
#include "bnlp_internal.h"
namespace bnlp {
constexpr double kInfD = 1e19;
}
#include "bnlp_ipm_coop.cuh"

namespace bnlp {
cudaError_t ipm_coop_set_smem(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(ipm_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ipm_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t ipm_coop_launch(bool timers, int ctas, size_t smem, cudaStream_t stream, DevModel &M, IpmArgs &A, CoopBufs &CB) {
  void *args[] = {(void *)&M, (void *)&A, (void *)&CB};
  return cudaLaunchCooperativeKernel(timers ? (void *)ipm_coop_kernel<true> : (void *)ipm_coop_kernel<false>, dim3(ctas),
                                     dim3(IPM_NT), args, smem, stream);
}
}  // namespace bnlp
