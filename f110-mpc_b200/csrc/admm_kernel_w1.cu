// admm_kernel instantiations for horizons 1..31 (one warp per QP; 2 / 4 QPs per warp when N + 1 <= 16 / 8).
#include "admm_kernel_impl.cuh"

namespace f110 {
static bool no_tmem() {
  static const bool v = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  return v;
}
cudaError_t launch_admm_w1(const KParams& p, cudaStream_t stream, int nlev) {
  switch (nlev) {
    case 1: return launch_one<1, 1, false, false, 4>(p, stream);
    case 2: return launch_one<2, 1, false, false, 4>(p, stream);
    case 3: return (p.N == 7) ? launch_one<3, 1, true, false, 4>(p, stream) : launch_one<3, 1, false, false, 4>(p, stream);
    case 4: return (p.N == 15) ? launch_one<4, 1, true, false, 2>(p, stream) : launch_one<4, 1, false, false, 2>(p, stream);
    case 5:
      // horizons 16..31: the tensor-memory variant (F110_NO_TMEM=1 selects the shared-memory kernel, for A/B measurements)
      if (p.work && !no_tmem()) return (p.N == 31) ? launch_tm<5, true>(p, stream) : launch_tm<5, false>(p, stream);
      return (p.N == 31) ? launch_one<5, 1, true>(p, stream) : launch_one<5, 1, false>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace f110
