// admm_kernel instantiations for horizons 32..63 (two warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w2(const KParams& p, cudaStream_t stream) {
  // the tensor-memory variant (F110_NO_TMEM=1 selects the shared-memory kernel, for A/B measurements)
  static const bool no_tmem = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  if (!no_tmem) return (p.N == 63) ? launch_tmw<6, 2, true>(p, stream) : launch_tmw<6, 2, false>(p, stream);
  return (p.N == 63) ? launch_one<6, 2, true>(p, stream) : launch_one<6, 2, false>(p, stream);
}
}  // namespace f110
