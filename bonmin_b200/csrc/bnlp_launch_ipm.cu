// Translation unit of the general batched IPM kernel (bnlp_ipm.cuh): instantiations + launchers.
#include <cuda_runtime.h>

#include "../../include/bnlp.h"
#include "bnlp_internal.h"
namespace bnlp {
constexpr double kInfD = 1e19;
}
#include "bnlp_ipm.cuh"

namespace bnlp {
cudaError_t ipm_general_set_smem(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(ipm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ipm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t ipm_general_launch(bool wide, int ctas, size_t smem, cudaStream_t stream, const DevModel &M, const IpmArgs &A) {
  if (wide) ipm_kernel<16><<<ctas, IPM_NT, smem, stream>>>(M, A);
  else ipm_kernel<8><<<ctas, IPM_NT, smem, stream>>>(M, A);
  return cudaGetLastError();
}
}  // namespace bnlp
