// admm_kernel instantiations for horizons 1..31 (one warp per QP; 2 / 4 QPs per warp when N + 1 <= 16 / 8).
#include "admm_kernel_impl.cuh"

namespace f110 {
static bool no_tmem() {
  static const bool v = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  return v;
}
cudaError_t launch_admm_w1(const KParams& p, cudaStream_t stream, int nlev) {
  // every horizon of the base row set runs the persistent tensor-memory kernel (F110_NO_TMEM=1 selects the shared-memory kernels,
  // for A/B measurements); N + 1 <= 16 / 8: 2 / 4 QPs side by side in one warp
  const bool tm = p.work && !no_tmem();
  switch (nlev) {
    case 1: return tm ? launch_tm<1, false, false, 4>(p, stream) : launch_one<1, 1, false, false, 4>(p, stream);
    case 2: return tm ? launch_tm<2, false, false, 4>(p, stream) : launch_one<2, 1, false, false, 4>(p, stream);
    case 3:
      if (tm) return (p.N == 7) ? launch_tm<3, true, false, 4>(p, stream) : launch_tm<3, false, false, 4>(p, stream);
      return (p.N == 7) ? launch_one<3, 1, true, false, 4>(p, stream) : launch_one<3, 1, false, false, 4>(p, stream);
    case 4:
      if (tm) return (p.N == 15) ? launch_tm<4, true, false, 2>(p, stream) : launch_tm<4, false, false, 2>(p, stream);
      return (p.N == 15) ? launch_one<4, 1, true, false, 2>(p, stream) : launch_one<4, 1, false, false, 2>(p, stream);
    case 5:
      if (tm) return (p.N == 31) ? launch_tm<5, true>(p, stream) : launch_tm<5, false>(p, stream);
      return (p.N == 31) ? launch_one<5, 1, true>(p, stream) : launch_one<5, 1, false>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace f110
