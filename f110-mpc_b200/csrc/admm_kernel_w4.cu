// admm_kernel instantiations for horizons 64..127 (four warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w4(const KParams& p, cudaStream_t stream) {
  return (p.N == 127) ? launch_one<7, 4, true>(p, stream) : launch_one<7, 4, false>(p, stream);
}
}  // namespace f110
