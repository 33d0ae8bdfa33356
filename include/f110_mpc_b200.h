/* f110_mpc_b200.h — C ABI of the B200-native batched per-cycle MPC solve.
 *
 * Drop-in boundary for the OSQP call inside the reference's MPC::Update
 * (reference src/mpc.cpp:81-142, the only user of the OsqpEigen::Solver member declared at
 * include/f110-mpc/mpc.h:63) and for the mini-path collision check of OdomCallback
 * (reference src/project.cpp:76-113).  Plain pointers and sizes only; no torch / Eigen / ROS
 * types.  All entry points return 0 on success, non-zero on an API error
 * (f110_last_error() gives the text).  There is NO CPU fallback behind this ABI: without a
 * CUDA device every compute entry fails with F110_ERR_CUDA.
 *
 * The QP is never shipped as CSC matrices.  One QP = one parameter record of
 * f110_mpc_record_doubles(N) doubles — exactly what MPC::Update receives per cycle
 * (mpc.cpp:69-80):
 *     x0[3]      current state (x, y, ori)                      mpc.cpp:71
 *     u_lin[2]   linearisation input (v, steer)                 mpc.cpp:73  -> model.cpp:30-59
 *     l1[3]      half-plane line 1 (a, b, c+0.5)                mpc.cpp:75  -> constraints.cpp:255-260
 *     l2[3]      half-plane line 2                              constraints.cpp:262-264
 *     ref[3*N]   desired states 0..N-1 (x, y, ori)              mpc.cpp:72, 221-229
 * The kernel linearises the kinematic bicycle (model.cpp:30-59), stacks the horizon
 * (mpc.cpp:208-306) and runs the OSQP ADMM iteration with a structure-exploiting KKT solve.
 *
 * Solution layout = the reference's (mpc.cpp:26-29):
 *     x[5N+3]  = [x_0(3) ... x_N(3) | u_0(2) ... u_{N-1}(2)]
 *     y[7N+5]  = [dynamics 3(N+1) | gap pairs 2(N+1) | input box 2N]
 * With f110_mpc_config.state_rows = 1, 3(N+1) state-box rows follow the input box (y has 10N+8 entries): row 3k+j bounds x_k[j].
 * With f110_mpc_config.rate_rows = 1, N steering-rate rows follow the input box (y has 8N+5 entries):
 *     row k:  delta_k - delta_{k-1} in [-rate_delta, +rate_delta]  (k >= 1);   row 0:  delta_0 - u_lin[1]  likewise.
 * The reference has no such rows (its relic of an extra input constraint is the commented slip block,
 * constraints.cpp:23-39, mpc.cpp:250); SURVEY.md section 8f rank 4 asks for them on top of the reference's row set.
 */
#ifndef F110_MPC_B200_H
#define F110_MPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F110_OK 0
#define F110_ERR_ARG 1
#define F110_ERR_CUDA 2
#define F110_ERR_UNSUPPORTED 3

/* Per-QP status codes: OSQP's status_val (what OsqpEigen::Solver::solve() tests, mpc.cpp:133). */
#define F110_SOLVED 1
#define F110_SOLVED_INACCURATE 2
#define F110_PRIMAL_INFEASIBLE_INACCURATE 3
#define F110_DUAL_INFEASIBLE_INACCURATE 4
#define F110_MAX_ITER_REACHED (-2)
#define F110_PRIMAL_INFEASIBLE (-3)
#define F110_DUAL_INFEASIBLE (-4)
#define F110_NON_CVX (-7)
#define F110_UNSOLVED (-10)

#define F110_MAX_HORIZON 127

/* What MPC::MPC reads from the parameter server (mpc.cpp:3-24) plus Constraints' input box
 * (constraints.cpp:18-21) and Model's wheelbase (model.cpp:32). */
typedef struct f110_mpc_config {
  int32_t horizon;   /* params.yaml:12 */
  int32_t gap_mode;  /* 0 = as shipped: gap rows bounded by (-INFTY, +INFTY) (mpc.cpp:297-298);
                        1 = lower bound -l(2) restored (the commented code on those lines), every stage;
                        2 = same, but the stage-0 pair — all-ones rows that are never overwritten
                            (mpc.cpp:237-241, 267) and are not half-planes — stays loose */
  double dt;         /* (double)0.01f — MPC::dt_ is a float (mpc.h:48) */
  double wheelbase;  /* (double)0.3302f (model.cpp:32) */
  double q[3];       /* state weights  q0 q1 q2 (params.yaml:1-3) */
  double r[2];       /* input weights  r0 r1    (params.yaml:5-6) */
  double u_des[2];   /* des_vel, des_steer      (params.yaml:42-43) */
  double u_min[2];   /* umin, -0.43f            (constraints.cpp:20-21) */
  double u_max[2];   /* umax, +0.43f            (constraints.cpp:18-19) */
  int32_t rate_rows; /* 0 = the reference's row set; 1 = append N steering-rate rows (horizon <= 63) */
  int32_t state_rows;/* 0 = the reference's row set; 1 = append the state box the reference stores but never stacks
                        (constraints.cpp:14-17 + Constraints::SetXLims, :108-114): 3(N+1) rows, x_k and y_k within
                        +-state_lim of the current state, ori_k free.  Horizon <= 31, not together with rate_rows */
  double rate_delta; /* max steering change per step (rad) = steering-rate limit (rad/s) * dt */
  double state_lim;  /* d of SetXLims (params.yaml:50 state_lims = 1) */
} f110_mpc_config;

/* The OSQP settings the reference leaves at their defaults (it only sets warm start + verbosity,
 * mpc.cpp:98-99), exported as knobs. */
typedef struct f110_solver_settings {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int32_t max_iter, check_termination, scaling;
  int32_t adaptive_rho, adaptive_rho_interval; /* interval 0 (OSQP's timing-based choice) is mapped to 25 */
  int32_t warm_start;                          /* 1: iterates and rho persist per QP slot between solves */
  int32_t scaled_termination;                  /* must be 0 (OSQP default) */
  int32_t reserved;
} f110_solver_settings;

typedef struct f110_mpc_solver f110_mpc_solver; /* opaque; owns device scratch only */

void f110_mpc_default_config(f110_mpc_config* cfg);
void f110_solver_default_settings(f110_solver_settings* s);
int f110_mpc_record_doubles(int horizon); /* 11 + 3N */
int f110_mpc_num_variables(int horizon);  /* 5N + 3  (mpc.cpp:26-28) */
int f110_mpc_num_constraints(int horizon);/* 7N + 5  (mpc.cpp:29) */
int f110_mpc_num_rows(const f110_mpc_config* cfg); /* 7N + 5, + N with rate_rows, + 3(N+1) with state_rows: length of one dual vector */
const char* f110_last_error(void);
int f110_device_count(void);

/* replaces: OsqpEigen::Solver construction + settings()/data()/initSolver() (mpc.cpp:98-129). */
int f110_mpc_create(const f110_mpc_config* cfg, const f110_solver_settings* settings, int max_batch,
                    int device, f110_mpc_solver** out);
void f110_mpc_destroy(f110_mpc_solver* s);

/* replaces: updateGradient / updateLinearConstraintsMatrix / updateBounds / solve / getSolution
 * (mpc.cpp:83-94, 133, 140) for `count` independent QPs.  HOST buffers; the call copies the
 * records to the device, solves, copies results back and synchronises.  Any output may be NULL.
 *   recs        count x f110_mpc_record_doubles(N), row stride `rec_stride` doubles
 *   x, y        primal / dual (reference layout); NaN-filled for infeasible QPs like OSQP
 *   u0          count x 2, first applied control (mpc.cpp:145-159 -> project.cpp:190-191)
 *   status      per-QP OSQP status_val;  iters  per-QP ADMM iterations */
int f110_mpc_solve_host(f110_mpc_solver* s, int count, const double* recs, int rec_stride, double* x,
                        double* y, double* u0, int32_t* status, int32_t* iters);

/* Same, DEVICE buffers, stream-ordered on `cuda_stream` (a cudaStream_t; NULL = default stream),
 * no synchronisation.  info: count x 4 = objective, primal residual, dual residual, rho at exit.
 * When d_recs is 16-byte aligned and rec_stride is even (e.g. record_doubles + 1), the kernel stages each record into
 * shared memory with one bulk asynchronous copy (TMA); otherwise it reads it with plain loads. */
int f110_mpc_solve_device(f110_mpc_solver* s, int count, const double* d_recs, int rec_stride, double* d_x,
                          double* d_y, double* d_u0, int32_t* d_status, int32_t* d_iters,
                          int32_t* d_rho_updates, double* d_info, void* cuda_stream);

/* Forget the warm-start state of every slot (next solve starts from x = z = y = 0, rho = settings.rho).
 * Ordering: the clear is queued on the handle's internal stream — i.e. after every f110_mpc_solve_host / f110_cycle_host call
 * made so far — and has completed when the call returns.  Solves the caller queued on its OWN streams through
 * f110_mpc_solve_device / f110_cycle_device are not waited for: synchronise those streams before calling reset. */
int f110_mpc_reset(f110_mpc_solver* s);
/* Multi-GPU helper: the NEXT f110_mpc_solve_device call also writes count x 4 doubles
 * (u0_v, u0_steer, status, iters) to d_packed — the row each rank contributes to the final gather of the
 * chosen controls, produced by the solve kernel itself so no packing kernel is needed.  One-shot. */
int f110_mpc_set_packed_output(f110_mpc_solver* s, double* d_packed);
/* Number of kernels the last solve call launched (for launch accounting). */
int f110_mpc_last_launches(const f110_mpc_solver* s);

/* ---- mini-path collision check (project.cpp:76-113 with occupancy_grid.cpp:27-33, 90-101, 165-168
 * and transforms.cpp:13-19).  Bit-exact integer/float contract; the tf2 rotation is an explicit input.
 *   grid        scenes x blocks*blocks floats, Eigen column-major: cell(row, col) at row + col*blocks
 *   offset      scenes x 2 floats   (OccGrid::occ_offset_)
 *   rot         scenes x 4 doubles  (R00 R01 R10 R11 of the car->world basis)
 *   pose_xy     scenes x 2 doubles  (pose.position.x, .y)
 *   table_xy    paths x samples x 2 doubles (mini-path table, base_link; shared by all scenes)
 * outputs (scenes x paths): valid (1 = every sample in-grid and free), free_count,
 *   end_world (x, y floats; the world end point of valid paths, project.cpp:108-111). */
int f110_collision_check_device(int scenes, int paths, int samples, int blocks, float discrete,
                                const float* d_grid, const float* d_offset, const double* d_rot,
                                const double* d_pose_xy, const double* d_table_xy, uint8_t* d_valid,
                                int32_t* d_free_count, float* d_end_world, void* cuda_stream);
int f110_collision_check_host(int scenes, int paths, int samples, int blocks, float discrete, const float* grid,
                              const float* offset, const double* rot, const double* pose_xy,
                              const double* table_xy, uint8_t* valid, int32_t* free_count, float* end_world,
                              int device);

/* ---- whole planning + control cycle on the device, no host round trip.
 * For each of `scenes` independent cars:  FillOccGrid (occupancy_grid.cpp:55-88)  ->  mini-path collision check
 * (project.cpp:76-113)  ->  look-ahead point + best surviving path (trajectory.cpp:81-108, project.cpp:121-149)
 * ->  FindHalfSpaces on the scene's scan (constraints.cpp:116-265)  ->  one tracking QP  ->  first control.
 * The transcendental calls of the fill / gap stages are not bit-identical to glibc's (see pipeline_kernels.cu);
 * the integer/float paths of the check and the selection are. */
typedef struct f110_cycle_config {
  int32_t n_beams;                                 /* ranges per scan */
  float angle_min, angle_max, angle_increment;     /* sensor_msgs/LaserScan, same for every scene */
  int32_t occ_size; float occ_discrete, occ_dilation; /* occupancy_grid.cpp:6-8 */
  float follow_gap_thresh, fov_divider, buffer;    /* constraints.cpp:9-12 */
  float lookahead;                                 /* trajectory.cpp:10 */
  int32_t use_half_spaces;                         /* 1: l1, l2 from each scene's scan; 0: zero rows */
  int32_t qp_mode;                                 /* 0: one QP per scene, for the selected path (the reference's behaviour);
                                                      1: one QP per (scene, path), skipped (F110_UNSOLVED) where the path collides
                                                         — BASELINE config 2, "a QP per surviving path";
                                                      2: one QP per (scene, path), colliding paths included.
                                                      Modes 1, 2: outputs u0/status/iters have scenes*paths rows (slot = scene*paths + path)
                                                      and scenes*paths must not exceed max_batch. */
  int32_t reserved;
  double v_lin;                                    /* linearisation speed, 4.5 (project.cpp:170) */
} f110_cycle_config;
void f110_cycle_default_config(f110_cycle_config* c);
/*   d_pose7      scenes x 7 doubles (px py pz qx qy qz qw)      d_ranges   scenes x n_beams floats
 *   d_prev_steer scenes doubles or NULL (previous steering, the linearisation point)
 *   d_table_xy   paths x samples x 2 doubles                     d_wp_xy    n_wp x 2 floats (raceline, trajectory.cpp:28-32)
 * outputs (scenes): u0 x2, status (F110_UNSOLVED where no path was valid), iters, chosen path index (-1 = none),
 *   d_valid scenes x paths or NULL.  Needs scenes <= max_batch of the handle.  Stream-ordered, no sync. */
int f110_cycle_device(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* d_pose7, const float* d_ranges,
                      const double* d_prev_steer, const double* d_table_xy, int paths, int samples, const float* d_wp_xy,
                      int n_wp, double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_chosen, uint8_t* d_valid,
                      void* cuda_stream);
/* Same cycle, HOST buffers (copies in, runs, copies back, synchronises): the reference-facing call for a whole
 * planning + control cycle — laser scans and poses in, controls out.  valid may be NULL.
 * Batches of 64 scenes or more run as two halves on two internal streams (the copies and small kernels of the second half
 * under the solve of the first); results do not depend on the split.  The mini-path table and the raceline are uploaded only
 * when their bytes change.  Environment: F110_CYCLE_CHUNKS=1..16 overrides the number of pipelined chunks (tuning only). */
int f110_cycle_host(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* pose7, const float* ranges,
                    const double* prev_steer, const double* table_xy, int paths, int samples, const float* wp_xy, int n_wp,
                    double* u0, int32_t* status, int32_t* iters, int32_t* chosen, uint8_t* valid);
/* Device buffers the last f110_cycle_device call filled (for inspection / tests): grids (scenes x blocks^2 floats),
 * offsets (x2 floats), l1l2 (x6 doubles), records (row stride = record_doubles rounded up to even), best_global (int32).
 * Any pointer may be NULL. */
int f110_cycle_buffers(f110_mpc_solver* s, float** d_grid, float** d_offset, double** d_l1l2, double** d_recs,
                       int32_t** d_best_global);

/* ---- asynchronous form of f110_cycle_host (the shape of the reference's own loop: OdomCallback computes cycle k+1 while
 * DriveLoop applies cycle k's result, project.cpp:160-191, 220-238).  f110_cycle_submit queues the copies and kernels of one
 * cycle and returns a ticket; f110_cycle_wait blocks until that cycle's results are in the caller's arrays.  At most TWO cycles
 * may be in flight per handle (f110_cycle_set_depth: 1..4): cycle k+1's host-to-device copies and perception kernels run under
 * cycle k's solve.  The solves themselves run in submission order when they share state — warm_start = 1 (the handle's warm-start
 * slots) or a problem family whose kernel keeps scratch lines in global memory (steering-rate / state-box rows, horizons below
 * 16); cold-started solves of the base row set overlap, which changes no result.  Inputs are read straight from the caller's
 * buffers when those are pinned (cudaHostAlloc / cudaHostRegister) and must then stay untouched until the matching wait; pageable
 * inputs are copied into the handle's own pinned staging before the call returns.  Results are identical to f110_cycle_host.
 * `gathered` (f110_cycle_wait): NULL, or — on the root of an attached gather ring — world x rows x 4 doubles. */
int f110_cycle_submit(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* pose7, const float* ranges,
                      const double* prev_steer, const double* table_xy, int paths, int samples, const float* wp_xy, int n_wp,
                      int* ticket);
int f110_cycle_wait(f110_mpc_solver* s, int ticket, double* u0, int32_t* status, int32_t* iters, int32_t* chosen, uint8_t* valid,
                    double* gathered);
/* Zero-copy alternative to `gathered` on the gather root: after f110_cycle_wait(ticket) the gathered rows of that cycle stay in
 * the handle's pinned host buffer until the `depth`-th next f110_cycle_submit; *rows points at world x rows x 4 doubles there. */
int f110_cycle_gathered_view(f110_mpc_solver* s, int ticket, const double** rows, size_t* doubles);
/* Number of cycles that may be in flight per handle (1..4, default 2).  Only while no cycle is in flight.  Deeper pipelines pay when
 * one QP of a batch runs far longer than the rest (a max_iter straggler): later cycles stream past it. */
int f110_cycle_set_depth(f110_mpc_solver* s, int depth);

/* ---- multi-GPU: QPs are independent, shards are contiguous, the only exchange is a final gather of the
 * packed rows (u0_v, u0_steer, status, iters) to one GPU — and that gather is the solve kernel's own store over NVLink.
 *
 * (1) one process, several GPUs: the reference-facing batched call (mirrors what MPC::Update does for one QP, mpc.cpp:69-143).
 *     The batch is cut into contiguous shards of whole `unit`s (QPs that belong together, e.g. the 140 lane x path QPs of one
 *     scenario); every GPU's kernel stores its rows into one buffer on devices[0] (peer access; one peer copy per GPU without it). */
typedef struct f110_mpc_multi f110_mpc_multi;
int f110_mpc_create_multi(const f110_mpc_config* cfg, const f110_solver_settings* settings, int max_batch, const int* devices,
                          int n_devices, f110_mpc_multi** out);
void f110_mpc_destroy_multi(f110_mpc_multi* m);
int f110_mpc_solve_multi_host(f110_mpc_multi* m, int count, int unit, const double* recs, int rec_stride, double* u0,
                              int32_t* status, int32_t* iters);
int f110_mpc_multi_devices(const f110_mpc_multi* m);
int f110_mpc_multi_uses_peer_stores(const f110_mpc_multi* m, int index);  /* 1: device `index` stores into devices[0] directly */
int f110_mpc_multi_last_shard(const f110_mpc_multi* m, int index, int* first_qp, int* count);

/* (2) one process per GPU (torchrun): a gather ring on the root rank's GPU, mapped into the other ranks through CUDA IPC.
 *     Layout: slots x world flags (int32, padded to 256 bytes), then `slots` slots of world x rows_per_rank x 4 doubles.  Rank r's
 *     cycle number c lands in slot c % slots, block r; afterwards the rank raises flags[c % slots][r] = c + 1 (f110_stream_signal)
 *     and the root's stream waits for the slot's flags (f110_stream_wait_flags: a stream memory operation, no kernel spins) before
 *     it reads the slot — cycles of one rank may therefore complete in any order.  No rank may be
 *     `slots` or more cycles ahead of the root's reads.  f110_gather_create zero-fills; the 64-byte handle goes to the peers by
 *     any host channel.  f110_cycle_set_gather attaches the ring to the asynchronous cycle entry (NULL detaches);
 *     f110_gather_slot + f110_mpc_set_packed_output do the same by hand for f110_mpc_solve_device. */
int f110_gather_bytes(int world, int rows_per_rank, int slots, size_t* bytes);
int f110_gather_create(int device, int world, int rows_per_rank, int slots, void** d_ring, unsigned char* ipc_handle64);
int f110_gather_open(int device, const unsigned char* ipc_handle64, void** d_ring);
int f110_gather_close(int device, void* d_ring, int opened);
int f110_gather_slot(void* d_ring, int world, int rank, int rows_per_rank, int slots, long long cycle, double** d_rows,
                     int32_t** d_flag);
int f110_stream_signal(void* cuda_stream, int32_t* d_flag, int32_t value);
int f110_stream_wait_flags(void* cuda_stream, const int32_t* d_flags, int n, int skip, int32_t value);
int f110_cycle_set_gather(f110_mpc_solver* s, void* d_ring, int world, int rank, int rows_per_rank, int slots);

/* ---- batched closed loop: `cars` simulated cars, each running the reference's control loop
 * (OdomCallback / ScanCallback / DriveLoop, project.cpp:41-238: plan a mini-path when none is held, otherwise one warm-started MPC
 * cycle per odometry tick against the held path, path dropped within 1.98 m of its end, previous inputs kept when a solve fails)
 * against the kinematic plant (Model::simulate_dynamics, model.cpp:61-76) — `ticks` ticks on the device with no host round trip.
 * Car b uses warm-start slot b of `s`, which must have been created with warm_start = 1 and max_batch >= cars; cc->qp_mode must be 0.
 * Each car keeps one scan (car frame) for the whole run.  Tick order: odometry, scan (every scan_every ticks), drive (every
 * drive_every ticks), plant step of dt_tick seconds.
 *   f110_fleet_reset   pose3: cars x (x, y, yaw);  ranges: cars x n_beams;  clears every car's state and the solver's warm starts
 *   f110_fleet_run     optional logs (both or neither), one row per (tick, car):
 *                        log_i x4: phase (0 plan, 1 idle before the first scan, 2 path dropped, 3 MPC cycle), chosen path of a planning
 *                                  tick (-1 none valid, -2 not a planning tick), status and iterations of an MPC cycle
 *                        log_d x13: x, y, yaw at the tick's odometry; input published (v, steer); l1, l2 of the tick; u0 of an MPC cycle */
typedef struct f110_fleet f110_fleet;
int f110_fleet_create(f110_mpc_solver* s, const f110_cycle_config* cc, int cars, const double* table_xy, int paths, int samples,
                      const float* wp_xy, int n_wp, int drive_every, int scan_every, double dt_tick, f110_fleet** out);
void f110_fleet_destroy(f110_fleet* f);
int f110_fleet_reset(f110_fleet* f, const double* pose3, const float* ranges);
int f110_fleet_run(f110_fleet* f, int ticks, int32_t* log_i, double* log_d);
int f110_fleet_get_pose(f110_fleet* f, double* pose3);

/* ---- measurement utility (not on the solve path): FP64 FMA issue rate of `device` in TFLOP/s, the
 * roofline denominator for the ADMM kernel. */
int f110_bench_fp64_fma(int device, int iters, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* F110_MPC_B200_H */
