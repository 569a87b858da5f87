// The WIDE build of the element-term machine: the general IPM kernel and the evaluation kernel compiled a second time in
// namespace bnlpw with BNLP_TV = BNLP_MAX_WIDE_TERM_VARS (24) local variables per term and a postfix stack of
// BNLP_MAX_WIDE_TERM_STACK entries.  A model gets here only when one of its terms couples more than BNLP_MAX_TERM_VARS
// variables (term_width_class, bnlp_model.cpp): a stack entry of the second-order forward mode is then 1 + 24 + 300 doubles
// instead of 28, so the narrow build stays what every other model runs on.  The argument structs have the same layout in
// both namespaces (nothing in them depends on TV); the per-term record stride and the Hessian-position table are built by
// the host with the same capacity (tv_cap in bnlp_api.cu / bnlp_api_eval.inc).
#include <cuda_runtime.h>

#include "../../include/bnlp.h"
#define BNLP_NS bnlpw
#define BNLP_TV BNLP_MAX_WIDE_TERM_VARS
#define BNLP_TS BNLP_MAX_WIDE_TERM_STACK
#define BNLP_EVAL_KERNELS      // eval_kernel (only that one depends on the term capacity) is instantiated here too
#include "bnlp_internal.h"
namespace bnlpw {
constexpr double kInfD = 1e19;
using bnlp::HostModel;
}
#include "bnlp_ipm.cuh"

cudaError_t bnlpw_ipm_set_smem(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(bnlpw::ipm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bnlpw::ipm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t bnlpw_ipm_launch(bool big, int ctas, size_t smem, cudaStream_t stream, const void *M, const void *A) {
  const bnlpw::DevModel &m = *static_cast<const bnlpw::DevModel *>(M);
  const bnlpw::IpmArgs &a = *static_cast<const bnlpw::IpmArgs *>(A);
  if (big) bnlpw::ipm_kernel<16><<<ctas, bnlpw::IPM_NT, smem, stream>>>(m, a);
  else bnlpw::ipm_kernel<8><<<ctas, bnlpw::IPM_NT, smem, stream>>>(m, a);
  return cudaGetLastError();
}
cudaError_t bnlpw_eval_launch(int nc, size_t smem, cudaStream_t stream, const void *ev, const void *args) {
  bnlpw::eval_kernel<<<nc, bnlpw::EV_NT, smem, stream>>>(*static_cast<const bnlpw::EvalModelDev *>(ev),
                                                         *static_cast<const bnlpw::EvalArgs *>(args));
  return cudaGetLastError();
}
