// Internal declarations shared by the translation units behind the C ABI (c_api.cu, async_api.cu, multi_api.cu).
#pragma once
#include <cstdint>
#include <string>

#include "../../include/f110_mpc_b200.h"
#include "admm_kernel.cuh"

namespace f110api {
int fail(int code, const std::string& msg);            // records the text f110_last_error() returns, hands back `code`
int cuda_fail(cudaError_t e, const char* what);
}  // namespace f110api
#define CUDA_TRY(expr)                                            \
  do {                                                            \
    cudaError_t e__ = (expr);                                     \
    if (e__ != cudaSuccess) return f110api::cuda_fail(e__, #expr); \
  } while (0)

// per-scene device scratch of one planning + control cycle (grown on demand)
struct f110_cycle_scratch {
  int blocks = 0, paths = 0;
  size_t cap_scenes = 0, cap_qps = 0;
  float *grid = nullptr, *offset = nullptr, *endw = nullptr;
  double *rot = nullptr, *pose_xy = nullptr, *state3 = nullptr, *l1l2 = nullptr, *recs = nullptr;
  uint8_t* valid = nullptr;
  int32_t *free_cnt = nullptr, *gap = nullptr, *best_global = nullptr;
  void release() {
    cudaFree(grid); cudaFree(offset); cudaFree(endw); cudaFree(rot); cudaFree(pose_xy); cudaFree(state3); cudaFree(l1l2);
    cudaFree(recs); cudaFree(valid); cudaFree(free_cnt); cudaFree(gap); cudaFree(best_global);
    *this = f110_cycle_scratch();
  }
};

// one pipeline of the asynchronous cycle entry (f110_cycle_submit / f110_cycle_wait): its own stream, staging and scratch
struct f110_cycle_lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_done = nullptr;
  cudaEvent_t ev_own = nullptr, ev_gather = nullptr;   // gather root: this rank's rows are in the slot / the slot is on the host
  bool gather_pending = false;
  unsigned char* stage = nullptr;      // device: [table | waypoints | inputs | outputs]
  size_t stage_bytes = 0;
  unsigned char* pin_in = nullptr;     // pinned staging of pageable caller inputs
  size_t pin_in_bytes = 0;
  unsigned char* pin_out = nullptr;    // pinned mirror of the output block (+ the gathered rows of all ranks on the gather root)
  size_t pin_out_bytes = 0;
  unsigned long long tab_hash = 0;
  f110_cycle_scratch cyc;
  bool busy = false;
  int ticket = -1;
  // layout of the cycle in flight
  int scenes = 0, paths = 0;
  long long nqp = 0;
  size_t o_u0 = 0, o_st = 0, o_it = 0, o_ch = 0, o_val = 0, b_out = 0, gather_bytes = 0;
};

struct f110_mpc_solver {
  f110_mpc_config cfg;
  f110_solver_settings st;
  int max_batch = 0;
  int device = 0;
  int last_launches = 0;
  double* d_state = nullptr;    // warm-start slots
  double* d_scratch = nullptr;  // per-QP scratch lines (scaling vectors, previous iterate)
  double* d_mult = nullptr;     // per-QP top-level multipliers of four-warp QPs (horizon >= 64)
  int* d_work = nullptr;        // work-counter pairs of the persistent tensor-memory kernel: WORK_SLOTS round-robin + 1 for the B = 1 graph
  unsigned work_seq = 0;
  // staging for the host-buffer entry: one device block [u0 | status | iters | x | y] so results come back in
  // one copy, plus a small pinned mirror used for latency-critical small batches
  double* d_recs = nullptr;
  unsigned char* d_out = nullptr;
  unsigned char* h_pin = nullptr;   // pinned: records of <= kSmallBatch QPs, then their outputs
  size_t out_bytes = 0;
  double* d_packed_next = nullptr;     // optional packed result rows for the NEXT solve (f110_mpc_set_packed_output)
  unsigned char* cyc_stage = nullptr;  // device staging of f110_cycle_host
  size_t cyc_stage_bytes = 0;
  unsigned char* cyc_pin = nullptr;    // pinned mirror of the output block
  size_t cyc_pin_bytes = 0;
  unsigned long long cyc_tab_hash = 0; // content hash of the uploaded mini-path table + raceline
  f110_cycle_scratch cyc;              // f110_cycle_device / f110_cycle_host scratch
  cudaStream_t stream = nullptr;
  // f110_cycle_host pipelines its scenes in chunks over two streams (copies of chunk c+1 under the kernels of chunk c)
  // single-QP latency path: the copy-in / solve / copy-out triple captured once as a CUDA graph per output shape (u0 only, +x, +x+y)
  cudaGraphExec_t lat_graph[3] = {nullptr, nullptr, nullptr};
  int32_t* done_flag_next = nullptr;   // single-QP latency path: completion flag (device view of mapped host memory) for the NEXT solve
  int32_t done_seq = 0;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_tab = nullptr, ev_join = nullptr;
  // asynchronous cycles: `depth` lanes (2 by default, f110_cycle_set_depth).  Everything before a solve overlaps the previous
  // cycle's solve.  The solves themselves are serialised through ev_solve when consecutive cycles share state (warm-start slots,
  // per-QP scratch lines in global memory); cold-started solves of a kernel that keeps its
  // working state on chip share nothing and overlap too — the next cycle's CTAs fill the SMs the previous solve's tail leaves idle.
  static constexpr int kMaxLanes = 4;
  f110_cycle_lane lane[kMaxLanes];
  int depth = 2;
  cudaEvent_t ev_solve = nullptr;
  cudaStream_t gather_stream = nullptr;   // gather root: waits for the other ranks' flags and copies the slot out
  int next_ticket = 0;
  // gather of the packed rows across GPUs (f110_cycle_set_gather): where this rank's rows go, the flag it raises, and — on the
  // root — the flags it waits for and the rows it copies out
  struct Gather {
    double* ring = nullptr;      // [slots][world][rows][4] doubles on the root GPU (peer-mapped on the others)
    int32_t* flags = nullptr;    // [slots][world] on the root GPU: flags[c % slots][r] = c + 1 once rank r has delivered cycle c
    int slots = 0, world = 0, rank = 0, rows = 0;
    long long seq = 0;           // cycles submitted with the gather on
  } gather;
};

namespace f110api {
int solve_device_range(f110_mpc_solver* s, int slot0, int count, const double* d_recs, int rec_stride, double* d_x, double* d_y,
                       double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_rho_updates, double* d_info, void* cuda_stream);
int cycle_prepare(f110_mpc_solver* s, f110_cycle_scratch& c, const f110_cycle_config* cc, int scenes, int paths, int samples, int n_wp,
                  const double* d_table_xy);
int cycle_device_range(f110_mpc_solver* s, f110_cycle_scratch& c, const f110_cycle_config* cc, int scene0, int scenes, const double* d_pose7,
                       const float* d_ranges, const double* d_prev_steer, const double* d_table_xy, int paths, int samples,
                       const float* d_wp_xy, int n_wp, double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_chosen,
                       uint8_t* d_valid, cudaStream_t st, cudaEvent_t wait_before_solve);
unsigned long long table_hash(const double* table_xy, size_t n_tab, const float* wp_xy, size_t n_wpb, int paths, int samples, int n_wp);
cudaError_t launch_signal(cudaStream_t st, int32_t* flag, int32_t value);
}  // namespace f110api
