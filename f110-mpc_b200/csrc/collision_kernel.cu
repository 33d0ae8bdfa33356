// Mini-path collision check — bit-exact restatement of the loop in project.cpp:76-113 for a batch of
// scenes.  One warp per (scene, path); lanes stride over the path's samples; the free-sample count is a
// ballot + popcount.  HBM-bound integer/float work: every table point is read once (coalesced 16-byte
// pairs), every grid cell touched is a 4-byte gather from the scene's 40 KB grid (L2 resident).
//
// The float/double operation order follows the reference literally and every operation is an explicit
// round-to-nearest intrinsic so the compiler cannot contract a*b+c into an FMA (the reference is built
// for baseline x86-64, no FMA):
//   transforms.cpp:13-19   world = basis * (x, y, 0) in double, + float(pose) , narrowed to float
//   occupancy_grid.cpp:30  col = int((x - off.x) / discrete + blocks / 2)      float math, truncation
//   occupancy_grid.cpp:90  InGrid ; occupancy_grid.cpp:165 IsOccupied = grid(row, col) != 0
#include "admm_kernel.cuh"

namespace f110 {

namespace {
// x86 cvttss2si semantics: out-of-range / NaN -> INT_MIN ("integer indefinite")
__device__ __forceinline__ int trunc_x86(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000;
  return __float2int_rz(v);
}

__global__ void __launch_bounds__(128) collision_kernel(int scenes, int paths, int samples, int blocks, float discrete,
                                                        const float* __restrict__ grid, const float* __restrict__ offset,
                                                        const double* __restrict__ rot, const double* __restrict__ pose_xy,
                                                        const double* __restrict__ table_xy, uint8_t* __restrict__ valid,
                                                        int32_t* __restrict__ free_count, float* __restrict__ end_world,
                                                        const int32_t* __restrict__ scene_gate, int gate_value) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)scenes * paths) return;
  const int sc = (int)(wid / paths), pa = (int)(wid % paths);
  // fleet loop: only the cars that plan on this tick need the check (project.cpp:73: OdomCallback plans when no path is held)
  if (scene_gate && scene_gate[sc] != gate_value) return;
  const float* g = grid + (size_t)sc * blocks * blocks;
  const float offx = offset[2 * sc], offy = offset[2 * sc + 1];
  const double r00 = rot[4 * sc], r01 = rot[4 * sc + 1], r10 = rot[4 * sc + 2], r11 = rot[4 * sc + 3];
  const float posex = (float)pose_xy[2 * sc], posey = (float)pose_xy[2 * sc + 1];  // transforms.cpp:17-18
  const float half = (float)(blocks / 2);
  const double2* tp = reinterpret_cast<const double2*>(table_xy) + (size_t)pa * samples;
  int free_pts = 0;
  float ex = 0.f, ey = 0.f;
  for (int j0 = 0; j0 < samples; j0 += 32) {
    const int j = j0 + lane;
    bool is_free = false;
    if (j < samples) {
      const double2 pt = tp[j];
      const double cx = (double)(float)pt.x, cy = (double)(float)pt.y;  // project.cpp:86 narrows to float
      const double zterm = __dmul_rn(0.0, 0.0);
      const double wx = __dadd_rn(__dadd_rn(__dmul_rn(r00, cx), __dmul_rn(r01, cy)), zterm);
      const double wy = __dadd_rn(__dadd_rn(__dmul_rn(r10, cx), __dmul_rn(r11, cy)), zterm);
      const float fx = (float)__dadd_rn(wx, (double)posex);  // transforms.cpp:19
      const float fy = (float)__dadd_rn(wy, (double)posey);
      const int col = trunc_x86(__fadd_rn(__fdiv_rn(__fsub_rn(fx, offx), discrete), half));  // occupancy_grid.cpp:30
      const int row = trunc_x86(__fadd_rn(__fdiv_rn(__fsub_rn(fy, offy), discrete), half));  // occupancy_grid.cpp:31
      const bool in_grid = !(col >= blocks || col < 0 || row >= blocks || row < 0);          // occupancy_grid.cpp:90-101
      if (in_grid) is_free = (g[(size_t)row + (size_t)col * blocks] == 0.f);                 // project.cpp:92 -> grid_(row, col)
      if (j == samples - 1) { ex = fx; ey = fy; }
    }
    free_pts += __popc(__ballot_sync(0xffffffffu, is_free));
  }
  const int last_lane = (samples - 1) & 31;
  ex = __shfl_sync(0xffffffffu, ex, last_lane);
  ey = __shfl_sync(0xffffffffu, ey, last_lane);
  if (lane == 0) {
    const bool ok = (free_pts == samples);  // project.cpp:103
    valid[wid] = ok ? 1 : 0;
    free_count[wid] = free_pts;
    end_world[2 * wid] = ok ? ex : 0.f;     // project.cpp:108-111
    end_world[2 * wid + 1] = ok ? ey : 0.f;
  }
}
}  // namespace

cudaError_t launch_collision(int scenes, int paths, int samples, int blocks, float discrete, const float* grid,
                             const float* offset, const double* rot, const double* pose_xy, const double* table_xy,
                             uint8_t* valid, int32_t* free_count, float* end_world, cudaStream_t stream,
                             const int32_t* scene_gate, int gate_value) {
  const long long warps = (long long)scenes * paths;
  if (warps == 0) return cudaSuccess;
  const int wpb = 4;
  const int grid_dim = (int)((warps + wpb - 1) / wpb);
  collision_kernel<<<grid_dim, 32 * wpb, 0, stream>>>(scenes, paths, samples, blocks, discrete, grid, offset, rot, pose_xy,
                                                      table_xy, valid, free_count, end_world, scene_gate, gate_value);
  return cudaGetLastError();
}

}  // namespace f110
