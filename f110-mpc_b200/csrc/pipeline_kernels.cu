// Perception / planning kernels that feed the two hot kernels without a host round trip:
//
//   fill_grid_kernel        OccGrid::FillOccGrid                      reference src/occupancy_grid.cpp:55-88
//   half_spaces_kernel      Constraints::FindHalfSpaces               reference src/constraints.cpp:116-265
//   select_build_kernel     Trajectory::get_best_global_idx           reference src/trajectory.cpp:81-108
//                           + best-path argmin + mini-path -> world   reference src/project.cpp:121-149
//                           + the parameter record MPC::Update gets   reference src/mpc.cpp:69-80
//
// This translation unit is compiled with -fmad=false: every float/double expression below is evaluated in the
// reference's order with IEEE round-to-nearest and no FMA contraction (the reference targets baseline x86-64).
// The only operations that are NOT bit-reproducible against glibc are the transcendental calls; cosf/sinf/atan2f
// on float arguments are computed here in double precision and rounded once to float, which is the correctly
// rounded float result up to a ~1e-8 double-rounding probability (glibc's own float routines are within 0.56 ulp).
// Parity for these kernels is therefore stated as a cell / index mismatch COUNT against the oracle, not bit-exact.
#include "admm_kernel.cuh"

namespace f110 {

namespace {

__device__ __forceinline__ float cosf_cr(float a) { return (float)cos((double)a); }
__device__ __forceinline__ float sinf_cr(float a) { return (float)sin((double)a); }
__device__ __forceinline__ int trunc_x86(float v) {  // cvttss2si
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000;
  return __float2int_rz(v);
}

// ---- tf2 restated (see host/transforms.cpp) ---------------------------------------------------------------------------
struct Basis { double m[3][3]; };
__device__ Basis basis_of(double x, double y, double z, double w) {
  const double n2 = x * x + y * y + z * z + w * w;
  const double s = 2.0 / n2;
  const double xs = x * s, ys = y * s, zs = z * s;
  const double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
  Basis b;
  b.m[0][0] = 1.0 - (yy + zz); b.m[0][1] = xy - wz;         b.m[0][2] = xz + wy;
  b.m[1][0] = xy + wz;         b.m[1][1] = 1.0 - (xx + zz); b.m[1][2] = yz - wx;
  b.m[2][0] = xz - wy;         b.m[2][1] = yz + wx;         b.m[2][2] = 1.0 - (xx + yy);
  return b;
}
__device__ void quaternion_of(const Basis& b, double* q) {
  const double trace = b.m[0][0] + b.m[1][1] + b.m[2][2];
  if (trace > 0.0) {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (b.m[2][1] - b.m[1][2]) * s; q[1] = (b.m[0][2] - b.m[2][0]) * s; q[2] = (b.m[1][0] - b.m[0][1]) * s;
  } else {
    const int i = b.m[0][0] < b.m[1][1] ? (b.m[1][1] < b.m[2][2] ? 2 : 1) : (b.m[0][0] < b.m[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(b.m[i][i] - b.m[j][j] - b.m[k][k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (b.m[k][j] - b.m[j][k]) * s; q[j] = (b.m[j][i] + b.m[i][j]) * s; q[k] = (b.m[k][i] + b.m[i][k]) * s;
  }
}

// ---- Constraints::FindHalfSpaces for one scene, executed by one warp.  The per-beam predicates do not depend on the
// visiting order: they are evaluated 32 at a time into ballot masks (coalesced reads); the run-length scan itself IS order
// dependent (SURVEY a13'), so lane 0 runs it sequentially and literally over the masks.
__device__ void find_half_spaces_warp(int lane, unsigned* fov, unsigned* far, int n_beams, int num_scans, float angle_min, float angle_inc,
                                      float ftg_thresh, float divider, float buffer, double px, double py, float heading,
                                      const float* __restrict__ rs, double* __restrict__ out, int32_t* __restrict__ gap) {
  const int nb = num_scans < n_beams ? num_scans : n_beams;
  const int words = (nb + 31) / 32;
  const float half_fov = 1.571f / divider;
  for (int w = 0; w < words; ++w) {
    const int i = 32 * w + lane;
    bool in_fov = false, is_far = false;
    if (i < nb) {
      const float bearing = angle_min + i * angle_inc;                          // constraints.cpp:133
      in_fov = bearing > -half_fov && bearing < half_fov;                       // :135
      is_far = rs[i] > ftg_thresh;                                              // :138
    }
    const unsigned m_fov = __ballot_sync(0xffffffffu, in_fov), m_far = __ballot_sync(0xffffffffu, in_fov && is_far);
    if (lane == 0) { fov[w] = m_fov; far[w] = m_far; }
  }
  __syncwarp();
  if (lane != 0) return;
  // The reference visits the in-view beams one by one (constraints.cpp:130-168).  Its state only changes where the "far"
  // predicate changes, and inside a run of far beams hi - lo grows by one per beam while the update test is a strict >,
  // so visiting the maximal stretches of equal predicate gives the same (widest, best_lo, best_hi): per stretch, the test
  // at its first beam (with the stale hi for a far stretch — `hi` is never reset, SURVEY a13') and the test at its last.
  auto next_bit = [words](const unsigned* m, int pos, bool want) -> int {   // first index >= pos whose bit == want
    int w = pos >> 5;
    if (w >= words) return words * 32;
    unsigned x = (want ? m[w] : ~m[w]) & (0xffffffffu << (pos & 31));
    while (!x) {
      if (++w >= words) return words * 32;
      x = want ? m[w] : ~m[w];
    }
    return w * 32 + __ffs(x) - 1;
  };
  int widest = -1, lo = -1, hi = -1, best_lo = 0, best_hi = 0;
  const int f0 = next_bit(fov, 0, true);
  if (f0 < nb) {
    int f1 = next_bit(fov, f0, false) - 1;
    if (f1 >= nb) f1 = nb - 1;
    const bool contiguous = next_bit(fov, f1 + 1, true) >= nb;
    if (contiguous) {
      int pos = f0;
      while (pos <= f1) {
        if ((far[pos >> 5] >> (pos & 31)) & 1u) {
          int e = next_bit(far, pos, false) - 1;
          if (e > f1) e = f1;
          lo = pos;                                                             // :147 (a stretch starts outside a gap)
          if (hi - lo > widest) { widest = hi - lo; best_hi = hi; best_lo = lo; }   // first beam of the run, stale hi
          if (e > pos) {
            hi = e;                                                             // :143 at the last beam of the run
            if (hi - lo > widest) { widest = hi - lo; best_hi = hi; best_lo = lo; }
          }
          pos = e + 1;
        } else {
          int e = next_bit(far, pos, true) - 1;
          if (e > f1) e = f1;
          if (hi - lo > widest) { widest = hi - lo; best_hi = hi; best_lo = lo; }   // :154-160
          pos = e + 1;
        }
      }
    } else {
      // non-monotone bearings: literal beam-by-beam visit
      bool inside = false;
      for (int w = 0; w < words; ++w) {
        unsigned mf = fov[w];
        const unsigned mr = far[w];
        while (mf) {
          const int bit = __ffs(mf) - 1;
          mf &= mf - 1;
          const int i = 32 * w + bit;
          if ((mr >> bit) & 1u) {
            if (inside) hi = i; else { lo = i; inside = true; }
          } else {
            inside = false;
            if (hi - lo > widest) { widest = hi - lo; best_hi = hi; best_lo = lo; }
          }
          if (hi - lo > widest) { widest = hi - lo; best_hi = hi; best_lo = lo; }
        }
      }
    }
  }
  if ((float)(best_hi - best_lo) > 2 * buffer) {                                // :173
    best_hi = (int)((float)best_hi - buffer);
    best_lo = (int)((float)best_lo + buffer);
  }
  gap[0] = best_lo; gap[1] = best_hi;
  if (best_lo < 0 || best_hi < 0 || best_lo >= n_beams || best_hi >= n_beams) {  // the reference reads ranges[-1] here
    for (int j = 0; j < 6; ++j) out[j] = 0.0;
    gap[0] = -1; gap[1] = -1;
    return;
  }
  const float a_lo = angle_min + best_lo * angle_inc + heading;                 // :179
  const float a_hi = angle_min + best_hi * angle_inc + heading;                 // :180
  const float p1x = (float)(rs[best_lo] * cosf_cr(a_lo) + px), p1y = (float)(rs[best_lo] * sinf_cr(a_lo) + py);  // :182-183
  const float p2x = (float)(rs[best_hi] * cosf_cr(a_hi) + px), p2y = (float)(rs[best_hi] * sinf_cr(a_hi) + py);  // :185-186
  const float qx = (float)px, qy = (float)py;                                   // :188-189
  float a1 = qy - p1y, b1 = p1x - qx, c1 = qx * p1y - qy * p1x;                 // :233-235
  if (a1 * p2x + b1 * p2y + c1 < 0) { a1 = -a1; b1 = -b1; c1 = -c1; }           // :237
  float a2 = qy - p2y, b2 = p2x - qx, c2 = qx * p2y - qy * p2x;                 // :244-246
  if (a2 * p1x + b2 * p1y + c2 < 0) { a2 = -a2; b2 = -b2; c2 = -c2; }           // :248
  out[0] = a1; out[1] = b1; out[2] = c1 + 0.5;                                  // :258-260
  out[3] = a2; out[4] = b2; out[5] = c2 + 0.5;                                  // :262-264
}

// ---- per-scene preparation, one CTA (8 warps) per scene, everything that only needs the pose and the scan:
//   warps 0..6  OccGrid::FillOccGrid: the grid is stamped as bytes in shared memory and written out once, coalesced
//   warp  7     Constraints::FindHalfSpaces on the scene's scan (skipped when l1l2 == nullptr; then it stamps too)
//   thread 0    the car->world rotation CarPointToWorldPoint applies after its tf2 round trip, pose xy
constexpr int PREP_THREADS = 256;
__global__ void __launch_bounds__(PREP_THREADS) scene_prep_kernel(int scenes, int blocks, float discrete, float dilation, int n_beams,
                                                                  int num_scans, float angle_min, float angle_inc, float ftg_thresh,
                                                                  float divider, float buffer, const double* __restrict__ pose7,
                                                                  const float* __restrict__ ranges, float* __restrict__ grid,
                                                                  float* __restrict__ offset, double* __restrict__ rot,
                                                                  double* __restrict__ pose_xy, double* __restrict__ l1l2,
                                                                  int32_t* __restrict__ gap, int mode) {
  // mode 0: everything;  1: no grid fill (rotation, pose, half-planes only: the grid of an earlier scan stays, as between two
  // ScanCallbacks of the reference, project.cpp:41-59);  2: the grid fill only
  extern __shared__ __align__(16) unsigned char prep_sm[];   // blocks * blocks occupancy bytes, then 2 x words mask words
  const int sc = blockIdx.x;
  if (sc >= scenes) return;
  const int ncell = blocks * blocks;
  unsigned char* cells = prep_sm;
  const int nthr = (int)blockDim.x;   // PREP_THREADS, or 32 when there is no grid to fill (mode 1: one warp per scene, all scenes resident at once)
  const bool do_fill = mode != 1, do_rest = mode != 2;
  unsigned* masks = reinterpret_cast<unsigned*>(prep_sm + (do_fill ? (ncell + 15) / 16 * 16 : 0));   // (no byte image without a fill)
  // grid_ = Zero (occupancy_grid.cpp:57): the byte image is cleared a 32-bit word at a time (its allocation is rounded up to 16 bytes)
  if (do_fill)
    for (int i = threadIdx.x; i < (ncell + 3) / 4; i += nthr) reinterpret_cast<unsigned*>(cells)[i] = 0u;
  const double* p = pose7 + 7 * (size_t)sc;
  const double qz = p[5], qw = p[6];
  const float yaw = (float)atan2(2 * qw * qz, 1 - 2 * qz * qz);                 // :60, also Transforms::GetCarOrientation
  const float offx = (float)(p[0] + 0.275 * cosf_cr(yaw));                      // :63
  const float offy = (float)(p[1] + 0.275 * sinf_cr(yaw));                      // :64
  const float* r = ranges + (size_t)sc * n_beams;
  if (threadIdx.x == 0 && do_fill) { offset[2 * sc] = offx; offset[2 * sc + 1] = offy; }
  if (threadIdx.x == 0 && do_rest) {
    double q[4];
    quaternion_of(basis_of(p[3], p[4], p[5], p[6]), q);
    const Basis b = basis_of(q[0], q[1], q[2], q[3]);
    rot[4 * sc] = b.m[0][0]; rot[4 * sc + 1] = b.m[0][1]; rot[4 * sc + 2] = b.m[1][0]; rot[4 * sc + 3] = b.m[1][1];
    pose_xy[2 * sc] = p[0]; pose_xy[2 * sc + 1] = p[1];
  }
  __syncthreads();
  const bool gap_warp = (l1l2 != nullptr) && ((int)threadIdx.x >= nthr - 32);
  if (gap_warp) {
    if (do_rest)
    // State(pose.x, pose.y, float yaw) of project.cpp:163-164
    find_half_spaces_warp(threadIdx.x & 31, masks, masks + (n_beams + 31) / 32, n_beams, num_scans, angle_min, angle_inc, ftg_thresh,
                          divider, buffer, p[0], p[1], yaw, r, l1l2 + 6 * (size_t)sc, gap + 2 * (size_t)sc);
  } else if (do_fill) {
    const int stampers = (l1l2 != nullptr) ? nthr - 32 : nthr;
    const float half = (float)(blocks / 2);
    const int nb = num_scans < n_beams ? num_scans : n_beams;
    for (int ii = threadIdx.x; ii < nb; ii += stampers) {
      const float angle = angle_min + ii * angle_inc + yaw;                     // :71
      double sn, cs;
      sincos((double)angle, &sn, &cs);
      float cx = r[ii] * (float)cs;                                             // :50
      float cy = r[ii] * (float)sn;                                             // :51
      cx += offx;                                                               // :73
      cy += offy;                                                               // :74
      // the row index does not depend on x_off: evaluate the (identical) y_off sequence once per beam, not once per x_off
      // (at most 8 stamps per axis: the launcher rejects 2*dilation/discrete + 1 > 8)
      int rows[8];
      float y_off = -dilation;
#pragma unroll
      for (int t = 0; t < 8; ++t) {                                             // :78
        rows[t] = -1;
        if (y_off <= dilation) rows[t] = trunc_x86(((cy + y_off) - offy) / discrete + half);   // :31 (uniform branch: the offsets are launch constants)
        y_off += discrete;
      }
      for (float x_off = -dilation; x_off <= dilation; x_off += discrete) {     // :76
        const int col = trunc_x86(((cx + x_off) - offx) / discrete + half);     // :80 -> :30
        if (col < 0 || col >= blocks) continue;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (rows[t] >= 0 && rows[t] < blocks) cells[rows[t] + col * blocks] = 1;   // :83  grid_(row, col) = 1
      }
    }
  }
  __syncthreads();
  if (!do_fill) return;
  // write-out: HBM-bound (4 bytes per cell, 40 KB per scene).  Four cells per thread and store: one 32-bit read of the byte image,
  // one 16-byte store, coalesced to 512 B per warp instruction.
  float* g = grid + (size_t)sc * ncell;
  if ((ncell & 3) == 0) {
    float4* g4 = reinterpret_cast<float4*>(g);   // (cudaMalloc base, 4 * ncell bytes per scene: 16-byte aligned)
    const unsigned* c4 = reinterpret_cast<const unsigned*>(cells);
    for (int i = threadIdx.x; i < ncell / 4; i += nthr) {
      const unsigned w = c4[i];
      g4[i] = make_float4((w & 0xffu) ? 1.f : 0.f, (w & 0xff00u) ? 1.f : 0.f, (w & 0xff0000u) ? 1.f : 0.f, (w & 0xff000000u) ? 1.f : 0.f);
    }
  } else {
    for (int i = threadIdx.x; i < ncell; i += nthr) g[i] = cells[i] ? 1.f : 0.f;
  }
}

// ---- look-ahead point and best surviving path: one warp per scene.  The lanes evaluate the per-waypoint
// distances and the two order-dependent argmin scans (strict <, float-narrowed running minimum — trajectory.cpp:103-107,
// project.cpp:132-135) in their closed forms.
constexpr int SB_WARPS = 4;
__global__ void __launch_bounds__(32 * SB_WARPS) select_kernel(int scenes, int paths, int n_wp, float lookahead,
                                                              const double* __restrict__ pose7, const float* __restrict__ wp_xy,
                                                              const uint8_t* __restrict__ valid, const float* __restrict__ end_world,
                                                              int32_t* __restrict__ chosen, int32_t* __restrict__ best_global,
                                                              const int32_t* __restrict__ scene_gate, int gate_value) {
  extern __shared__ double off_sm[];  // SB_WARPS x n_wp look-ahead offsets (negative = waypoint behind the car)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sc = blockIdx.x * SB_WARPS + warp;
  if (sc >= scenes) return;
  if (scene_gate && scene_gate[sc] != gate_value) return;   // fleet loop: only the cars that plan on this tick
  double* offs = off_sm + (size_t)warp * n_wp;
  const double* p = pose7 + 7 * (size_t)sc;
  // any valid path? (project.cpp:115-119)
  bool mine = false;
  for (int i = lane; i < paths; i += 32) mine |= valid[(size_t)sc * paths + i] != 0;
  if (!__any_sync(0xffffffffu, mine)) {
    if (lane == 0) { chosen[sc] = -1; best_global[sc] = -1; }
    return;
  }
  // Transforms::WorldToCarTransform (transforms.cpp:22-31) then TransformPoint per waypoint (:33-44)
  const Basis b = basis_of(p[3], p[4], p[5], p[6]);
  Basis inv;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) inv.m[r][c] = b.m[c][r];
  const double ox = -p[0], oy = -p[1], oz = -p[2];
  const double tx = inv.m[0][0] * ox + inv.m[0][1] * oy + inv.m[0][2] * oz;
  const double ty = inv.m[1][0] * ox + inv.m[1][1] * oy + inv.m[1][2] * oz;
  double q[4];
  quaternion_of(inv, q);
  const Basis w2c = basis_of(q[0], q[1], q[2], q[3]);
#pragma unroll 4   // independent waypoints: keep several global loads and square roots in flight
  for (int i = lane; i < n_wp; i += 32) {                                       // trajectory.cpp:93, any order: values only
    const double wx = (double)wp_xy[2 * i], wy = (double)wp_xy[2 * i + 1];
    const float fx = (float)((w2c.m[0][0] * wx + w2c.m[0][1] * wy + w2c.m[0][2] * 0.0) + tx);
    const float fy = (float)((w2c.m[1][0] * wx + w2c.m[1][1] * wy + w2c.m[1][2] * 0.0) + ty);
    const double dist = sqrt((double)fx * (double)fx + (double)fy * (double)fy);  // pow(pow(x,2)+pow(y,2), 0.5)
    offs[i] = (fx < 0) ? -1.0 : fabs(dist - (double)lookahead);                 // :100, :102
  }
  __syncwarp();
  // The reference scans the waypoints in order with a FLOAT running minimum (trajectory.cpp:103-107):
  //     if (off_i < (double)best) { best = (float)off_i; idx = i; }
  // Rounding is monotone, so `best` never increases and ends at F = min_i (float)off_i; the first waypoint whose rounded
  // offset is F is always accepted (whatever `best` was before it is a float above F, hence above off_i), and after it
  // only waypoints with off_i < F (they round up to F) are accepted.  So the scan's answer has a closed form,
  //     idx = max( g, max{ i : (float)off_i == F and off_i < F } ),   g = min{ i : (float)off_i == F },
  // which every lane can evaluate on its own waypoints (tests/test_host_cpu.py checks it against the literal loop).
  float fmin_l = 3.402823466e+38f;   // numeric_limits<float>::max(): nothing ahead leaves it there
  for (int i = lane; i < n_wp; i += 32) {
    const double off = offs[i];
    if (off >= 0.0) { const float f = (float)off; fmin_l = f < fmin_l ? f : fmin_l; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, fmin_l, o); fmin_l = v < fmin_l ? v : fmin_l; }
  const float F = fmin_l;
  int g_l = 0x7fffffff, late_l = -1;
  for (int i = lane; i < n_wp; i += 32) {
    const double off = offs[i];
    if (off >= 0.0 && (float)off == F) {
      g_l = i < g_l ? i : g_l;
      if (off < (double)F) late_l = i;   // ascending i per lane: the last one stays
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const int a = __shfl_xor_sync(0xffffffffu, g_l, o), b = __shfl_xor_sync(0xffffffffu, late_l, o);
    g_l = a < g_l ? a : g_l;
    late_l = b > late_l ? b : late_l;
  }
  const int best_idx = (g_l == 0x7fffffff) ? -1 : (late_l > g_l ? late_l : g_l);
  // best path: smallest end-point distance to that waypoint, strict <, first wins (project.cpp:125-136) = the lowest index
  // among the paths at the minimum distance
  int pick = -1;
  if (best_idx >= 0) {
    const double gx = (double)wp_xy[2 * best_idx], gy = (double)wp_xy[2 * best_idx + 1];
    double dmin_all = 1.7976931348623157e308;                                   // project.cpp:125
    for (int base = 0; base < paths; base += 32) {
      const int i = base + lane;
      double d = 1.7976931348623157e308;
      if (i < paths && valid[(size_t)sc * paths + i]) {
        const double ex = (double)end_world[2 * ((size_t)sc * paths + i)], ey = (double)end_world[2 * ((size_t)sc * paths + i) + 1];
        d = sqrt((ex - gx) * (ex - gx) + (ey - gy) * (ey - gy));
      }
      double m = d;
#pragma unroll
      for (int o = 16; o; o >>= 1) { const double v = __shfl_xor_sync(0xffffffffu, m, o); m = v < m ? v : m; }
      if (m < dmin_all) {                                                       // a later chunk wins only with a strictly smaller distance
        dmin_all = m;
        const unsigned hit = __ballot_sync(0xffffffffu, d == m && i < paths && valid[(size_t)sc * paths + i]);
        pick = base + __ffs(hit) - 1;
      }
    }
  }
  if (lane == 0) {
    best_global[sc] = best_idx;
    chosen[sc] = pick;
  }
}

// ---- parameter records (include/f110_mpc_b200.h: x0 | (v, steer) | l1 | l2 | ref[0..N-1]), one warp per record ------------
// qp_mode 0: one record per scene, for the selected path (the reference's behaviour, project.cpp:141-149 + mpc.cpp:69-80)
// qp_mode 1: one record per (scene, path), empty where the path collides  (BASELINE config 2: a QP per surviving path)
// qp_mode 2: one record per (scene, path), colliding paths included
__global__ void __launch_bounds__(32 * SB_WARPS) build_records_kernel(int scenes, int paths, int samples, int N, int stride, int qp_mode,
                                                                     double v_lin, const double* __restrict__ pose7,
                                                                     const double* __restrict__ rot, const uint8_t* __restrict__ valid,
                                                                     const int32_t* __restrict__ chosen, const double* __restrict__ table_xy,
                                                                     const double* __restrict__ prev_steer, const double* __restrict__ l1l2,
                                                                     double* __restrict__ recs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long slot = (long long)blockIdx.x * SB_WARPS + warp;
  const long long nslots = qp_mode == 0 ? scenes : (long long)scenes * paths;
  if (slot >= nslots) return;
  const int sc = qp_mode == 0 ? (int)slot : (int)(slot / paths);
  int pick = qp_mode == 0 ? chosen[sc] : (int)(slot % paths);
  if (qp_mode == 1 && !valid[(size_t)sc * paths + pick]) pick = -1;
  double* rec = recs + (size_t)slot * stride;
  if (pick < 0) {
    if (lane == 0) rec[3] = __longlong_as_double(0x7ff8000000000000LL);  // empty slot: the solve kernel reports F110_UNSOLVED
    return;
  }
  const double* p = pose7 + 7 * (size_t)sc;
  if (lane == 0) {
    const float yaw = (float)atan2(2 * p[6] * p[5], 1 - 2 * p[5] * p[5]);       // Transforms::GetCarOrientation (float)
    rec[0] = p[0]; rec[1] = p[1]; rec[2] = (double)yaw;                         // project.cpp:163-164
    rec[3] = v_lin;                                                             // project.cpp:170
    rec[4] = prev_steer ? prev_steer[sc] : 0.0;
  }
  if (lane < 6) rec[5 + lane] = l1l2 ? l1l2[6 * (size_t)sc + lane] : 0.0;
  const double r00 = rot[4 * sc], r01 = rot[4 * sc + 1], r10 = rot[4 * sc + 2], r11 = rot[4 * sc + 3];
  const float posex = (float)p[0], posey = (float)p[1];
  const double* tp = table_xy + (size_t)pick * samples * 2;
  for (int k = lane; k < N; k += 32) {                                          // project.cpp:145-149, clamped to the path length
    const int kk = k < samples ? k : samples - 1;
    const double cx = (double)(float)tp[2 * kk], cy = (double)(float)tp[2 * kk + 1];
    const float fx = (float)(((r00 * cx + r01 * cy) + 0.0 * 0.0) + (double)posex);
    const float fy = (float)(((r10 * cx + r11 * cy) + 0.0 * 0.0) + (double)posey);
    rec[11 + 3 * k] = (double)fx; rec[12 + 3 * k] = (double)fy; rec[13 + 3 * k] = 0.0;
  }
}

}  // namespace

cudaError_t launch_scene_prep(int scenes, int blocks, float discrete, float dilation, int n_beams, int num_scans, float angle_min,
                              float angle_inc, float thresh, float divider, float buffer, const double* pose7, const float* ranges,
                              float* grid, float* offset, double* rot, double* pose_xy, double* l1l2, int32_t* gap, cudaStream_t st, int mode) {
  if (scenes == 0) return cudaSuccess;
  const size_t smem = (mode == 1 ? 0 : (size_t)(blocks * blocks + 15) / 16 * 16) + 2 * (size_t)((n_beams + 31) / 32) * sizeof(unsigned);
  // without a grid to fill the only parallel work is the gap finder's warp: one-warp CTAs keep every scene resident at once (the
  // run-length scan of a scene is one lane's serial work) and need no shared memory for the byte image
  scene_prep_kernel<<<scenes, (mode == 1 && l1l2) ? 32 : PREP_THREADS, smem, st>>>(scenes, blocks, discrete, dilation, n_beams, num_scans, angle_min, angle_inc, thresh,
                                                        divider, buffer, pose7, ranges, grid, offset, rot, pose_xy, l1l2, gap, mode);
  return cudaGetLastError();
}
cudaError_t launch_select(int scenes, int paths, int n_wp, float lookahead, const double* pose7, const float* wp_xy, const uint8_t* valid,
                          const float* end_world, int32_t* chosen, int32_t* best_global, cudaStream_t st, const int32_t* scene_gate,
                          int gate_value) {
  if (scenes == 0) return cudaSuccess;
  select_kernel<<<(scenes + SB_WARPS - 1) / SB_WARPS, 32 * SB_WARPS, (size_t)SB_WARPS * n_wp * sizeof(double), st>>>(
      scenes, paths, n_wp, lookahead, pose7, wp_xy, valid, end_world, chosen, best_global, scene_gate, gate_value);
  return cudaGetLastError();
}
cudaError_t launch_build_records(int scenes, int paths, int samples, int N, int stride, int qp_mode, double v_lin, const double* pose7,
                                 const double* rot, const uint8_t* valid, const int32_t* chosen, const double* table_xy,
                                 const double* prev_steer, const double* l1l2, double* recs, cudaStream_t st) {
  const long long nslots = qp_mode == 0 ? scenes : (long long)scenes * paths;
  if (nslots == 0) return cudaSuccess;
  build_records_kernel<<<(unsigned)((nslots + SB_WARPS - 1) / SB_WARPS), 32 * SB_WARPS, 0, st>>>(
      scenes, paths, samples, N, stride, qp_mode, v_lin, pose7, rot, valid, chosen, table_xy, prev_steer, l1l2, recs);
  return cudaGetLastError();
}

}  // namespace f110
