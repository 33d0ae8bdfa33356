// admm_kernel instantiations with steering-rate rows, horizons 32..63 (two warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w2r(const KParams& p, cudaStream_t stream) {
  return (p.N == 63) ? launch_one<6, 2, true, true>(p, stream) : launch_one<6, 2, false, true>(p, stream);
}
}  // namespace f110
