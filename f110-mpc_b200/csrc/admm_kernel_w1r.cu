// admm_kernel instantiations with steering-rate rows, horizons 1..31 (one warp per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w1r(const KParams& p, cudaStream_t stream, int nlev) {
  switch (nlev) {
    case 1: return launch_one<1, 1, false, true>(p, stream);
    case 2: return launch_one<2, 1, false, true>(p, stream);
    case 3: return launch_one<3, 1, false, true>(p, stream);
    case 4: return launch_one<4, 1, false, true>(p, stream);
    case 5: return (p.N == 31) ? launch_one<5, 1, true, true>(p, stream) : launch_one<5, 1, false, true>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace f110
