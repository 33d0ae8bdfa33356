// Kernel parameter block + launch entry of the batched ADMM solve (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef ADMM_W2_GLOBAL_LEVELS
#define ADMM_W2_GLOBAL_LEVELS 1   // two-warp QPs: top PCR levels whose multipliers live in global memory (0 or 1; 1 = 4 QPs per SM, 2-4 % faster)
#endif
#ifndef ADMM_MIN_BLOCKS
#define ADMM_MIN_BLOCKS 8   // resident CTAs per SM the register allocation must allow
#endif

namespace f110 {

// per-QP scratch line in global memory: scaling vectors, previous iterate, factor-step inputs (row indices SCR_* in
// admm_kernel_impl.cuh), one column per stage
constexpr int SCR_ROWS_ALLOC = 56;                       // rows of the per-QP scratch line (44 used, 56 with state-box rows)
constexpr int SCRATCH_DOUBLES = SCR_ROWS_ALLOC * 128;    // line length at 4 warps per QP (horizon 64..127); 32 / 64 columns below
constexpr int TM_COLS = 256;                             // tensor-memory columns one persistent CTA of the TMEM variant allocates (two CTAs per SM own all 512)
constexpr int WORK_SLOTS = 64;                           // work-counter pairs a handle cycles through (one per launch in flight) + one for the captured B = 1 graph

struct KParams {
  // problem family (f110_mpc_config)
  int N, B, stride, gap_mode;
  double dt, wheelbase;
  double Q[3], R[2], u_des[2], u_min[2], u_max[2];
  double qu[2];            // -R u_des
  double one_minus_alpha;
  // steering-rate rows (f110_mpc_config.rate_rows): N extra rows  delta_k - delta_{k-1} in [-rate_delta, rate_delta]
  int rate_rows;
  double rate_delta;
  // state-box rows (f110_mpc_config.state_rows): 3(N+1) extra rows  x_k, y_k in [x_cur -+ state_lim], ori_k free
  int state_rows;
  double state_lim;
  // OSQP settings (f110_solver_settings)
  double rho0, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval, warm_start;
  // TMA staging of the parameter record (set by the launcher): bytes of one bulk copy (0 = plain loads) and where the
  // record lands in dynamic shared memory (in doubles)
  int rec_bulk_bytes, rec_smem_offset;
  int tm_unit_doubles;   // multi-warp tensor-memory variant: shared-memory doubles per QP of a CTA (set by the launcher)
  // buffers (device)
  const double* recs;
  double* x_out;      // [B][5N+3] or null
  double* y_out;      // [B][7N+5 (+N with rate rows)] or null
  double* u0_out;     // [B][2] or null
  int32_t* status;    // [B] or null
  int32_t* iters;     // [B] or null
  int32_t* rho_updates;  // [B] or null
  double* info;       // [B][4] or null
  double* packed;     // [B][4] = (u0_v, u0_steer, status, iters) as doubles, the row that is gathered across GPUs; or null
  double* state;      // [B][state_doubles(N)] warm-start slots (scaled iterates x, z, y + rho + flag) or null
  double* scratch;    // [B][SCRATCH_DOUBLES]
  double* mult_global;  // [B][28 * 128] top-level PCR multipliers of multi-warp QPs (horizon >= 32; used from 64 up), else null
  double* scratch_dummy;  // 4 more lines: lane groups without a QP (several short-horizon QPs per warp, odd batch) scribble here
  int* work;              // tensor-memory variant: {next QP, warps run dry}, both 0 at launch; the kernel re-arms them itself
  // single-QP latency path (f110_mpc_solve_host, one QP): results go straight to mapped pinned host memory; the last warp of the
  // persistent kernel then raises this flag (mapped host memory too) to done_seq, after a system-wide fence.  Null otherwise.
  int32_t* done_flag;
  int32_t done_seq;
};

// constraint rows: dynamics 3(N+1) | gap pairs 2(N+1) | input box 2N | steering rate N (optional) | state box 3(N+1) (optional)
__host__ __device__ inline int num_rows(int N, int rate_rows, int state_rows = 0) { return 7 * N + 5 + (rate_rows ? N : 0) + (state_rows ? 3 * (N + 1) : 0); }
// doubles per warm-start slot: x(n) + z(m) + y(m) + rho + valid flag
__host__ __device__ inline int state_doubles(int N, int rate_rows, int state_rows = 0) { return (5 * N + 3) + 2 * num_rows(N, rate_rows, state_rows) + 2; }

// Launch the solve for p.B QPs on `stream`. Returns the cudaError of the launch.
cudaError_t launch_admm(const KParams& p, cudaStream_t stream, int* launches);
// True when the kernel launch_admm picks for this problem family keeps all per-QP working state on chip (tensor memory + shared
// memory: horizons 16..127 of the base row set): two launches of one handle may then overlap without sharing scratch lines.
bool admm_state_on_chip(int N, int rate_rows, int state_rows);

cudaError_t launch_collision(int scenes, int paths, int samples, int blocks, float discrete, const float* grid,
                             const float* offset, const double* rot, const double* pose_xy,
                             const double* table_xy, uint8_t* valid, int32_t* free_count, float* end_world,
                             cudaStream_t stream, const int32_t* scene_gate = nullptr, int gate_value = 0);   // gate: scenes with scene_gate[s] != gate_value are skipped

// perception / planning kernels (pipeline_kernels.cu, compiled with -fmad=false)
cudaError_t launch_scene_prep(int scenes, int blocks, float discrete, float dilation, int n_beams, int num_scans, float angle_min,
                              float angle_inc, float thresh, float divider, float buffer, const double* pose7, const float* ranges,
                              float* grid, float* offset, double* rot, double* pose_xy, double* l1l2, int32_t* gap, cudaStream_t st,
                              int mode = 0);   // 0: everything, 1: no grid fill, 2: the grid fill only
cudaError_t launch_select(int scenes, int paths, int n_wp, float lookahead, const double* pose7, const float* wp_xy, const uint8_t* valid,
                          const float* end_world, int32_t* chosen, int32_t* best_global, cudaStream_t st,
                          const int32_t* scene_gate = nullptr, int gate_value = 0);
cudaError_t launch_build_records(int scenes, int paths, int samples, int N, int stride, int qp_mode, double v_lin, const double* pose7,
                                 const double* rot, const uint8_t* valid, const int32_t* chosen, const double* table_xy,
                                 const double* prev_steer, const double* l1l2, double* recs, cudaStream_t st);

}  // namespace f110
