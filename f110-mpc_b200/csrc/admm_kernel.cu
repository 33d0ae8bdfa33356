// Launch dispatch of the batched ADMM solve.  The kernel template lives in admm_kernel_impl.cuh and is instantiated in
// one translation unit per warps-per-QP class (and per row set) so the instantiations compile in parallel.
#include <cstdlib>

#include "admm_kernel.cuh"

namespace f110 {

cudaError_t launch_admm_w1(const KParams& p, cudaStream_t stream, int nlev);  // horizons 1..31,  one warp per QP
cudaError_t launch_admm_w2(const KParams& p, cudaStream_t stream);            // horizons 32..63, two warps per QP
cudaError_t launch_admm_w4(const KParams& p, cudaStream_t stream);            // horizons 64..127, four warps per QP
cudaError_t launch_admm_w1r(const KParams& p, cudaStream_t stream, int nlev); // + steering-rate rows, horizons 1..31
cudaError_t launch_admm_w2r(const KParams& p, cudaStream_t stream);           // + steering-rate rows, horizons 32..63
cudaError_t launch_admm_w1s(const KParams& p, cudaStream_t stream, int nlev); // + state-box rows, horizons 1..31

bool admm_state_on_chip(int N, int rate_rows, int state_rows) {
  static const bool no_tmem = [] { const char* e = std::getenv("F110_NO_TMEM"); return e && e[0] == '1'; }();
  return !no_tmem && !rate_rows && !state_rows && N >= 16 && N <= 127;
}

cudaError_t launch_admm(const KParams& p, cudaStream_t stream, int* launches) {
  int nlev = 0;
  while ((1 << nlev) <= p.N) ++nlev;
  if (launches) *launches = 1;
  if (p.state_rows) {   // (f110_mpc_create refuses them with steering-rate rows or above horizon 31)
    if (p.rate_rows || nlev < 1 || nlev > 5) return cudaErrorInvalidValue;
    return launch_admm_w1s(p, stream, nlev);
  }
  if (p.rate_rows) {
    // (horizons above 63 are refused at create: the 4x4 multipliers of a four-warp QP fit neither the strip nor the shared memory)
    if (nlev >= 1 && nlev <= 5) return launch_admm_w1r(p, stream, nlev);
    if (nlev == 6) return launch_admm_w2r(p, stream);
    return cudaErrorInvalidValue;
  }
  if (nlev >= 1 && nlev <= 5) return launch_admm_w1(p, stream, nlev);
  if (nlev == 6) return launch_admm_w2(p, stream);
  if (nlev == 7) return launch_admm_w4(p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace f110
