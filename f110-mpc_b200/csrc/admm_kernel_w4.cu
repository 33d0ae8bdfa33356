// admm_kernel instantiations for horizons 64..127 (four warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w4(const KParams& p, cudaStream_t stream) {
  // the tensor-memory variant (F110_NO_TMEM=1 selects the shared-memory kernel, for A/B measurements)
  static const bool no_tmem = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  if (!no_tmem) return (p.N == 127) ? launch_tmw<7, 4, true>(p, stream) : launch_tmw<7, 4, false>(p, stream);
  return (p.N == 127) ? launch_one<7, 4, true>(p, stream) : launch_one<7, 4, false>(p, stream);
}
}  // namespace f110
