// admm_kernel instantiations with steering-rate rows, horizons 32..63 (two warps per QP).
#include "admm_kernel_impl.cuh"

namespace f110 {
cudaError_t launch_admm_w2r(const KParams& p, cudaStream_t stream) {
  // the tensor-memory variant: four of the five two-sided PCR levels in the strip, two CTAs of two QPs per SM instead of two QPs
  // (F110_NO_TMEM=1 selects the shared-memory kernel, for A/B measurements)
  static const bool no_tmem = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  if (!no_tmem) return (p.N == 63) ? launch_tmw<6, 2, true, true>(p, stream) : launch_tmw<6, 2, false, true>(p, stream);
  return (p.N == 63) ? launch_one<6, 2, true, true>(p, stream) : launch_one<6, 2, false, true>(p, stream);
}
}  // namespace f110
