// Batched closed loop: S simulated cars, each running the reference's control loop
// (project.cpp:62-238) against the kinematic plant (Model::simulate_dynamics, model.cpp:61-76), T ticks on the device without a
// host round trip.  Per tick and car, in the order of the ROS-free driver of the single-car loop (host/host_capi.cpp,
// f110h_closed_loop):
//
//   OdomCallback   no mini-path held  -> planning cycle (project.cpp:73-157): collision check of the path table against the car's
//                                        occupancy grid (as last filled), look-ahead point, best surviving path -> held path
//                  mini-path held     -> (once a scan has been seen, :167) input_to_pass = the held input at the drive index or
//                                        (0.5, 0), speed overwritten with 4.5 (:169-170); within 1.98 m of the path's end the path
//                                        is dropped and the MPC cycle skipped (:180-186), otherwise one warm-started MPC cycle
//                                        (mpc.cpp:69-143) on the held path with the half-planes of the car's scan at the current
//                                        state; a solve that is not "solved" keeps the previous input trajectory (:133-142);
//                                        either way the drive index returns to 0 (:190-191)
//   ScanCallback   every scan_every ticks: the occupancy grid is refilled at the current pose (project.cpp:41-59)
//   DriveLoop      every drive_every ticks: publish the held input at the drive index, or (0.5, 0) when it ran out (:210-236)
//   plant          x <- x + dt (v cos th, v sin th, v tan(delta) / 0.35)
//
// Each car keeps ONE scan for the whole run (the reference's MPC keeps the first scan it sees for good, project.cpp:45-49; there
// is no world map here to ray-cast new ones), so the scan is a per-car constant in the car frame.
// Everything per-car lives on the device: pose, grid, held path, held inputs, drive index, warm-start slot b = car b.
#include <cstring>
#include <vector>

#include "api_internal.h"

using f110api::cuda_fail;
using f110api::fail;

namespace {

enum : int { PH_PLAN = 0, PH_IDLE = 1, PH_DROP = 2, PH_CONTROL = 3 };
constexpr int LOG_I = 4, LOG_D = 13;   // per (tick, car): phase, chosen, status, iters | x, y, yaw, applied v, steer, l1l2[6], u0[2]

struct FleetState {
  int cars, N, samples, paths, rec_stride, nvar;
  double* pose3;        // [cars][3] x, y, yaw
  double* pose7;        // [cars][7]
  double* applied;      // [cars][2] last published input
  int32_t* has_path;    // [cars]
  int32_t* first_scan;  // [cars]
  int32_t* phase;       // [cars]
  double* path;         // [cars][samples][2] held mini-path, world frame (float-narrowed like project.cpp:145-149)
  double* held;         // [cars][N][2] held input trajectory
  int32_t* held_n;      // [cars]
  uint32_t* idx;        // [cars] drive index
  double* pass;         // [cars][2] input_to_pass of this tick
};

// pose7 of the tick + phase decision (one thread per car)
__global__ void fleet_begin_kernel(FleetState f) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.cars) return;
  const double x = f.pose3[3 * c], y = f.pose3[3 * c + 1], yaw = f.pose3[3 * c + 2];
  double* p = f.pose7 + 7 * (size_t)c;
  p[0] = x; p[1] = y; p[2] = 0.0; p[3] = 0.0; p[4] = 0.0; p[5] = sin(yaw / 2.0); p[6] = cos(yaw / 2.0);
  int ph;
  if (!f.has_path[c]) ph = PH_PLAN;
  else if (!f.first_scan[c]) ph = PH_IDLE;                                      // project.cpp:167
  else {
    // GetNextInput (project.cpp:210-218), speed overwritten (project.cpp:170)
    const uint32_t i = f.idx[c];
    const double steer = (i < (uint32_t)f.held_n[c]) ? f.held[((size_t)c * f.N + i) * 2 + 1] : 0.0;
    f.pass[2 * c] = 4.5; f.pass[2 * c + 1] = steer;
    // Transforms::CalcDist on float pairs (project.cpp:172-182, transforms.cpp:46-49)
    const double* e = f.path + ((size_t)c * f.samples + (f.samples - 1)) * 2;
    const float ex = (float)e[0], ey = (float)e[1], cx = (float)x, cy = (float)y;
    const double dx = (double)__fsub_rn(cx, ex), dy = (double)__fsub_rn(cy, ey);
    const float dist = (float)sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // pow(dx,2) + pow(dy,2): no FMA contraction
    if (dist < 1.98) { ph = PH_DROP; f.has_path[c] = 0; }                        // project.cpp:182-186
    else ph = PH_CONTROL;
  }
  f.phase[c] = ph;
}

// planning result -> held path (one warp per car)
__global__ void fleet_apply_plan_kernel(FleetState f, const int32_t* __restrict__ chosen, const double* __restrict__ rot,
                                        const double* __restrict__ table_xy) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= f.cars || f.phase[c] != PH_PLAN) return;
  const int pick = chosen[c];
  if (pick < 0) return;                                                         // "NO VALID TRAJS": stay without a path
  const double r00 = rot[4 * c], r01 = rot[4 * c + 1], r10 = rot[4 * c + 2], r11 = rot[4 * c + 3];
  const float posex = (float)f.pose3[3 * c], posey = (float)f.pose3[3 * c + 1];
  const double* tp = table_xy + (size_t)pick * f.samples * 2;
  for (int k = lane; k < f.samples; k += 32) {                                  // project.cpp:141-149 via transforms.cpp:3-20
    const double cx = (double)(float)tp[2 * k], cy = (double)(float)tp[2 * k + 1];
    const float fx = (float)(((r00 * cx + r01 * cy) + 0.0 * 0.0) + (double)posex);
    const float fy = (float)(((r10 * cx + r11 * cy) + 0.0 * 0.0) + (double)posey);
    f.path[((size_t)c * f.samples + k) * 2] = (double)fx;
    f.path[((size_t)c * f.samples + k) * 2 + 1] = (double)fy;
  }
  __syncwarp();
  if (lane == 0) f.has_path[c] = 1;
}

// parameter records of the cars in the control phase; an empty slot (NaN linearisation speed) for everybody else
__global__ void fleet_records_kernel(FleetState f, const double* __restrict__ l1l2, double* __restrict__ recs) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= f.cars) return;
  double* rec = recs + (size_t)c * f.rec_stride;
  if (f.phase[c] != PH_CONTROL) {
    if (lane == 0) rec[3] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  const double* p = f.pose7 + 7 * (size_t)c;
  if (lane == 0) {
    const float yaw = (float)atan2(2 * p[6] * p[5], 1 - 2 * p[5] * p[5]);       // Transforms::GetCarOrientation (project.cpp:163)
    rec[0] = p[0]; rec[1] = p[1]; rec[2] = (double)yaw;
    rec[3] = f.pass[2 * c]; rec[4] = f.pass[2 * c + 1];
  }
  if (lane < 6) rec[5 + lane] = l1l2[6 * (size_t)c + lane];
  for (int k = lane; k < f.N; k += 32) {                                        // mpc.cpp:221-229 reads desired[0..N-1]
    const int kk = k < f.samples ? k : f.samples - 1;
    rec[11 + 3 * k] = f.path[((size_t)c * f.samples + kk) * 2];
    rec[12 + 3 * k] = f.path[((size_t)c * f.samples + kk) * 2 + 1];
    rec[13 + 3 * k] = 0.0;
  }
}

// solve result -> held inputs; scan / drive bookkeeping; log; plant (one thread per car)
__global__ void fleet_end_kernel(FleetState f, const double* __restrict__ x_sol, const int32_t* __restrict__ status,
                                 const int32_t* __restrict__ iters, const int32_t* __restrict__ chosen, const double* __restrict__ l1l2,
                                 int scan_tick, int drive_tick, double dt, int32_t* __restrict__ log_i, double* __restrict__ log_d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.cars) return;
  const int ph = f.phase[c];
  double u0v = __longlong_as_double(0x7ff8000000000000LL), u0s = u0v;
  if (ph == PH_CONTROL) {
    const double* xs = x_sol + (size_t)c * f.nvar + 3 * (f.N + 1);
    u0v = xs[0]; u0s = xs[1];
    if (status[c] == 1) {                                                       // OsqpEigen solve() == true only for "solved" (mpc.cpp:133)
      int n = 0;
      for (int k = 0; k < f.N; ++k) {                                           // UpdateSolvedTrajectory (mpc.cpp:145-159): stops at a NaN
        const double v = xs[2 * k], a = xs[2 * k + 1];
        if (v != v || a != a) break;
        f.held[((size_t)c * f.N + k) * 2] = v; f.held[((size_t)c * f.N + k) * 2 + 1] = a;
        n = k + 1;
      }
      f.held_n[c] = n;
    }
    f.idx[c] = 0;                                                               // project.cpp:190-191
  } else if (ph == PH_DROP) {
    f.idx[c] = 0;                                                               // the skipped Update is still followed by :190-191
  }
  if (scan_tick) f.first_scan[c] = 1;                                           // project.cpp:45-49 (the grid fill is a separate launch)
  if (drive_tick && f.first_scan[c]) {                                          // DriveLoop body (project.cpp:224-236)
    const uint32_t i = f.idx[c];
    if (i < (uint32_t)f.held_n[c]) { f.applied[2 * c] = f.held[((size_t)c * f.N + i) * 2]; f.applied[2 * c + 1] = f.held[((size_t)c * f.N + i) * 2 + 1]; }
    else { f.applied[2 * c] = 0.5; f.applied[2 * c + 1] = 0.0; }
    f.idx[c] = i + 1;
  }
  const double x = f.pose3[3 * c], y = f.pose3[3 * c + 1], th = f.pose3[3 * c + 2], v = f.applied[2 * c], de = f.applied[2 * c + 1];
  if (log_i) {
    int32_t* li = log_i + (size_t)c * LOG_I;
    li[0] = ph; li[1] = (ph == PH_PLAN) ? chosen[c] : -2; li[2] = (ph == PH_CONTROL) ? status[c] : 0; li[3] = (ph == PH_CONTROL) ? iters[c] : 0;
    double* ld = log_d + (size_t)c * LOG_D;
    ld[0] = x; ld[1] = y; ld[2] = th; ld[3] = v; ld[4] = de;
    for (int j = 0; j < 6; ++j) ld[5 + j] = l1l2[6 * (size_t)c + j];
    ld[11] = u0v; ld[12] = u0s;
  }
  // Model::simulate_dynamics (model.cpp:61-76)
  const double r0 = v * cos(th), r1 = v * sin(th), r2 = tan(de) * v / 0.35;
  f.pose3[3 * c] = x + r0 * dt; f.pose3[3 * c + 1] = y + r1 * dt; f.pose3[3 * c + 2] = th + r2 * dt;
}

}  // namespace

struct f110_fleet {
  f110_mpc_solver* s = nullptr;
  f110_cycle_config cc;
  int cars = 0, paths = 0, samples = 0, n_wp = 0, drive_every = 2, scan_every = 4;
  long long tick = 0;
  double dt = 0.01;
  FleetState st{};
  f110_cycle_scratch cyc;
  double* d_table = nullptr;
  float* d_wp = nullptr;
  float* d_ranges = nullptr;
  double* d_x = nullptr;
  int32_t *d_status = nullptr, *d_iters = nullptr, *d_chosen = nullptr;
  uint8_t* d_valid = nullptr;
  int32_t* d_log_i = nullptr;
  double* d_log_d = nullptr;
  size_t log_ticks = 0;
};

extern "C" {

void f110_fleet_destroy(f110_fleet* f) {
  if (!f) return;
  cudaSetDevice(f->s->device);
  cudaDeviceSynchronize();
  FleetState& t = f->st;
  cudaFree(t.pose3); cudaFree(t.pose7); cudaFree(t.applied); cudaFree(t.has_path); cudaFree(t.first_scan); cudaFree(t.phase);
  cudaFree(t.path); cudaFree(t.held); cudaFree(t.held_n); cudaFree(t.idx); cudaFree(t.pass);
  f->cyc.release();
  cudaFree(f->d_table); cudaFree(f->d_wp); cudaFree(f->d_ranges); cudaFree(f->d_x); cudaFree(f->d_status); cudaFree(f->d_iters);
  cudaFree(f->d_chosen); cudaFree(f->d_valid); cudaFree(f->d_log_i); cudaFree(f->d_log_d);
  delete f;
}

int f110_fleet_create(f110_mpc_solver* s, const f110_cycle_config* cc, int cars, const double* table_xy, int paths, int samples,
                      const float* wp_xy, int n_wp, int drive_every, int scan_every, double dt_tick, f110_fleet** out) {
  if (!s || !cc || !table_xy || !wp_xy || !out) return fail(F110_ERR_ARG, "f110_fleet_create: null argument");
  if (cars < 1 || cars > s->max_batch) return fail(F110_ERR_ARG, "f110_fleet_create: car count exceeds the solver's max_batch");
  if (drive_every < 1 || scan_every < 1 || !(dt_tick > 0.0)) return fail(F110_ERR_ARG, "f110_fleet_create: bad tick settings");
  if (!s->st.warm_start) return fail(F110_ERR_ARG, "f110_fleet_create: the solver must be created with warm_start = 1 (the reference's setting, mpc.cpp:98)");
  if (cc->qp_mode != 0) return fail(F110_ERR_ARG, "f110_fleet_create: qp_mode must be 0 (one QP per car)");
  CUDA_TRY(cudaSetDevice(s->device));
  f110_fleet* f = new f110_fleet();
  f->s = s; f->cc = *cc; f->cars = cars; f->paths = paths; f->samples = samples; f->n_wp = n_wp;
  f->drive_every = drive_every; f->scan_every = scan_every; f->dt = dt_tick;
  const int N = s->cfg.horizon;
  FleetState& t = f->st;
  t.cars = cars; t.N = N; t.samples = samples; t.paths = paths; t.rec_stride = (f110_mpc_record_doubles(N) + 1) & ~1; t.nvar = 5 * N + 3;
  cudaError_t e = cudaSuccess;
  auto alloc = [&e](auto** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  const size_t C = cars;
  alloc(&t.pose3, C * 3 * sizeof(double)); alloc(&t.pose7, C * 7 * sizeof(double)); alloc(&t.applied, C * 2 * sizeof(double));
  alloc(&t.has_path, C * sizeof(int32_t)); alloc(&t.first_scan, C * sizeof(int32_t)); alloc(&t.phase, C * sizeof(int32_t));
  alloc(&t.path, C * samples * 2 * sizeof(double)); alloc(&t.held, C * N * 2 * sizeof(double)); alloc(&t.held_n, C * sizeof(int32_t));
  alloc(&t.idx, C * sizeof(uint32_t)); alloc(&t.pass, C * 2 * sizeof(double));
  alloc(&f->d_table, (size_t)paths * samples * 2 * sizeof(double)); alloc(&f->d_wp, (size_t)n_wp * 2 * sizeof(float));
  alloc(&f->d_ranges, C * cc->n_beams * sizeof(float)); alloc(&f->d_x, C * t.nvar * sizeof(double));
  alloc(&f->d_status, C * sizeof(int32_t)); alloc(&f->d_iters, C * sizeof(int32_t)); alloc(&f->d_chosen, C * sizeof(int32_t));
  alloc(&f->d_valid, C * paths);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_table, table_xy, (size_t)paths * samples * 2 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_wp, wp_xy, (size_t)n_wp * 2 * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { f110_fleet_destroy(f); return cuda_fail(e, "f110_fleet_create"); }
  const int rc = f110api::cycle_prepare(s, f->cyc, cc, cars, paths, samples, n_wp, f->d_table);
  if (rc != F110_OK) { f110_fleet_destroy(f); return rc; }
  *out = f;
  return F110_OK;
}

int f110_fleet_reset(f110_fleet* f, const double* pose3, const float* ranges) {
  if (!f || !pose3 || !ranges) return fail(F110_ERR_ARG, "f110_fleet_reset: null argument");
  f110_mpc_solver* s = f->s;
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  FleetState& t = f->st;
  const size_t C = f->cars;
  const int blocks = (int)(f->cc.occ_size / f->cc.occ_discrete);
  CUDA_TRY(cudaMemcpyAsync(t.pose3, pose3, C * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(f->d_ranges, ranges, C * f->cc.n_beams * sizeof(float), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemsetAsync(t.has_path, 0, C * sizeof(int32_t), s->stream));
  CUDA_TRY(cudaMemsetAsync(t.first_scan, 0, C * sizeof(int32_t), s->stream));
  CUDA_TRY(cudaMemsetAsync(t.held_n, 0, C * sizeof(int32_t), s->stream));
  CUDA_TRY(cudaMemsetAsync(t.idx, 0, C * sizeof(uint32_t), s->stream));
  CUDA_TRY(cudaMemsetAsync(f->cyc.grid, 0, C * blocks * blocks * sizeof(float), s->stream));    // OccGrid starts empty (occupancy_grid.cpp:3-11)
  CUDA_TRY(cudaMemsetAsync(f->cyc.offset, 0, C * 2 * sizeof(float), s->stream));
  std::vector<double> app(2 * C);
  for (size_t c = 0; c < C; ++c) { app[2 * c] = 0.5; app[2 * c + 1] = 0.0; }                      // nothing published yet
  CUDA_TRY(cudaMemcpyAsync(t.applied, app.data(), 2 * C * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  f->tick = 0;
  return f110_mpc_reset(s);
}

int f110_fleet_run(f110_fleet* f, int ticks, int32_t* log_i, double* log_d) {
  if (!f || ticks < 0) return fail(F110_ERR_ARG, "f110_fleet_run: bad argument");
  if ((log_i == nullptr) != (log_d == nullptr)) return fail(F110_ERR_ARG, "f110_fleet_run: pass both logs or neither");
  if (ticks == 0) return F110_OK;
  f110_mpc_solver* s = f->s;
  CUDA_TRY(cudaSetDevice(s->device));
  const f110_cycle_config* cc = &f->cc;
  FleetState& t = f->st;
  const int C = f->cars;
  if (log_i && (size_t)ticks > f->log_ticks) {
    cudaFree(f->d_log_i); cudaFree(f->d_log_d); f->d_log_i = nullptr; f->d_log_d = nullptr; f->log_ticks = 0;
    CUDA_TRY(cudaMalloc(&f->d_log_i, (size_t)ticks * C * LOG_I * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&f->d_log_d, (size_t)ticks * C * LOG_D * sizeof(double)));
    f->log_ticks = ticks;
  }
  cudaStream_t st = s->stream;
  auto& c = f->cyc;
  const int blocks = (int)(cc->occ_size / cc->occ_discrete);
  const int num_scans = (int)((cc->angle_max - cc->angle_min) / cc->angle_increment + 1);
  const int tpb = 128, wpb = 4;
  s->last_launches = 0;
  for (int k = 0; k < ticks; ++k, ++f->tick) {
    fleet_begin_kernel<<<(C + tpb - 1) / tpb, tpb, 0, st>>>(t);
    // rotation / pose / half-planes at the current pose; the grid stays as the last ScanCallback left it
    cudaError_t e = f110::launch_scene_prep(C, blocks, cc->occ_discrete, cc->occ_dilation, cc->n_beams, num_scans, cc->angle_min,
                                            cc->angle_increment, cc->follow_gap_thresh, cc->fov_divider, cc->buffer, t.pose7, f->d_ranges,
                                            c.grid, c.offset, c.rot, c.pose_xy, c.l1l2, c.gap, st, 1);
    if (e == cudaSuccess) e = f110::launch_collision(C, f->paths, f->samples, blocks, cc->occ_discrete, c.grid, c.offset, c.rot, c.pose_xy,
                                                     f->d_table, f->d_valid, c.free_cnt, c.endw, st, t.phase, PH_PLAN);
    if (e == cudaSuccess) e = f110::launch_select(C, f->paths, f->n_wp, cc->lookahead, t.pose7, f->d_wp, f->d_valid, c.endw, f->d_chosen,
                                                  c.best_global, st, t.phase, PH_PLAN);
    if (e != cudaSuccess) return cuda_fail(e, "f110_fleet_run: planning kernels");
    fleet_apply_plan_kernel<<<(C + wpb - 1) / wpb, 32 * wpb, 0, st>>>(t, f->d_chosen, c.rot, f->d_table);
    fleet_records_kernel<<<(C + wpb - 1) / wpb, 32 * wpb, 0, st>>>(t, c.l1l2, c.recs);
    const int rc = f110api::solve_device_range(s, 0, C, c.recs, t.rec_stride, f->d_x, nullptr, nullptr, f->d_status, f->d_iters, nullptr, nullptr, st);
    if (rc != F110_OK) return rc;
    const int scan_tick = (f->tick % f->scan_every) == 0, drive_tick = (f->tick % f->drive_every) == 0;
    fleet_end_kernel<<<(C + tpb - 1) / tpb, tpb, 0, st>>>(t, f->d_x, f->d_status, f->d_iters, f->d_chosen, c.l1l2, scan_tick, drive_tick, f->dt,
                                                         log_i ? f->d_log_i + (size_t)k * C * LOG_I : nullptr,
                                                         log_i ? f->d_log_d + (size_t)k * C * LOG_D : nullptr);
    if (scan_tick) {   // FillOccGrid at the pose of this tick's odometry (the plant step above has not been seen by a callback yet)
      e = f110::launch_scene_prep(C, blocks, cc->occ_discrete, cc->occ_dilation, cc->n_beams, num_scans, cc->angle_min, cc->angle_increment,
                                  cc->follow_gap_thresh, cc->fov_divider, cc->buffer, t.pose7, f->d_ranges, c.grid, c.offset, c.rot, c.pose_xy,
                                  nullptr, c.gap, st, 2);
      if (e != cudaSuccess) return cuda_fail(e, "f110_fleet_run: grid fill");
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "f110_fleet_run: fleet kernels");
    s->last_launches += 7 + scan_tick;   // begin, scene prep, collision, select, apply plan, records, end (+ the solve, counted above) + grid fill
  }
  if (log_i) {
    CUDA_TRY(cudaMemcpyAsync(log_i, f->d_log_i, (size_t)ticks * C * LOG_I * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(log_d, f->d_log_d, (size_t)ticks * C * LOG_D * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return F110_OK;
}

int f110_fleet_get_pose(f110_fleet* f, double* pose3) {
  if (!f || !pose3) return fail(F110_ERR_ARG, "f110_fleet_get_pose: null argument");
  CUDA_TRY(cudaSetDevice(f->s->device));
  CUDA_TRY(cudaMemcpy(pose3, f->st.pose3, (size_t)f->cars * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  return F110_OK;
}

}  // extern "C"
