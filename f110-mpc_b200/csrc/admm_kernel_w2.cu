// admm_kernel instantiations for horizons 32..63 (two warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w2(const KParams& p, cudaStream_t stream) {
  return (p.N == 63) ? launch_one<6, 2, true>(p, stream) : launch_one<6, 2, false>(p, stream);
}
}  // namespace f110
