// C ABI (include/f110_mpc_b200.h) over the CUDA kernels.  No CPU fallback: every compute entry needs a
// CUDA device and fails with F110_ERR_CUDA otherwise.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <string>

#include "api_internal.h"

namespace {
thread_local std::string g_err;
}
namespace f110api {
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) { return fail(F110_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
}  // namespace f110api
using f110api::cuda_fail;
using f110api::fail;

extern "C" {

const char* f110_last_error(void) { return g_err.c_str(); }

int f110_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void f110_mpc_default_config(f110_mpc_config* c) {
  // params.yaml values with the reference's C++ destination types
  c->horizon = 30;
  c->gap_mode = 0;
  c->dt = (double)0.01f;            // float dt_ (mpc.h:48) widened at mpc.cpp:73
  c->wheelbase = (double)0.3302f;   // model.cpp:32
  c->q[0] = 10.0; c->q[1] = 10.0; c->q[2] = 0.0;
  c->r[0] = 0.10; c->r[1] = 5.0;
  c->u_des[0] = 4.5; c->u_des[1] = 0.0;
  c->u_min[0] = (double)3.0f; c->u_min[1] = (double)-0.43f;  // constraints.cpp:20-21
  c->u_max[0] = (double)4.5f; c->u_max[1] = (double)0.43f;   // constraints.cpp:18-19
  c->rate_rows = 0; c->rate_delta = 0.0;    // the reference has no steering-rate rows
  c->state_rows = 0; c->state_lim = 1.0;    // the state box is stored (params.yaml:50 state_lims = 1) but never stacked
}

void f110_solver_default_settings(f110_solver_settings* s) {
  // OSQP v0.6.x defaults; the reference overrides only warm_start (mpc.cpp:98)
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->check_termination = 25; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 25;
  s->warm_start = 1; s->scaled_termination = 0; s->reserved = 0;
}

int f110_mpc_record_doubles(int N) { return 11 + 3 * N; }
int f110_mpc_num_variables(int N) { return 5 * N + 3; }
int f110_mpc_num_constraints(int N) { return 7 * N + 5; }
int f110_mpc_num_rows(const f110_mpc_config* c) { return c ? f110::num_rows(c->horizon, c->rate_rows, c->state_rows) : 0; }

int f110_mpc_create(const f110_mpc_config* cfg, const f110_solver_settings* st, int max_batch, int device,
                    f110_mpc_solver** out) {
  if (!cfg || !st || !out || max_batch <= 0) return fail(F110_ERR_ARG, "f110_mpc_create: null argument or max_batch <= 0");
  if (cfg->horizon < 1 || cfg->horizon > F110_MAX_HORIZON) return fail(F110_ERR_ARG, "f110_mpc_create: horizon out of range");
  if (cfg->gap_mode < 0 || cfg->gap_mode > 2) return fail(F110_ERR_ARG, "f110_mpc_create: gap_mode must be 0, 1 or 2");
  if (cfg->rate_rows && cfg->horizon > 63) return fail(F110_ERR_UNSUPPORTED, "f110_mpc_create: steering-rate rows need horizon <= 63");
  if (cfg->rate_rows && !(cfg->rate_delta >= 0.0)) return fail(F110_ERR_ARG, "f110_mpc_create: rate_delta must be >= 0");
  if (cfg->state_rows && (cfg->rate_rows || cfg->horizon > 31)) return fail(F110_ERR_UNSUPPORTED, "f110_mpc_create: state-box rows need horizon <= 31 and no steering-rate rows");
  if (cfg->state_rows && !(cfg->state_lim >= 0.0)) return fail(F110_ERR_ARG, "f110_mpc_create: state_lim must be >= 0");
  if (st->scaled_termination) return fail(F110_ERR_UNSUPPORTED, "f110_mpc_create: scaled_termination = 1 is not supported");
  if (st->max_iter < 1 || st->check_termination < 0 || st->scaling < 0) return fail(F110_ERR_ARG, "f110_mpc_create: bad settings");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(F110_ERR_CUDA, "f110_mpc_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(F110_ERR_ARG, "f110_mpc_create: bad device index");
  CUDA_TRY(cudaSetDevice(device));
  f110_mpc_solver* s = new f110_mpc_solver();
  s->cfg = *cfg;
  s->st = *st;
  if (s->st.adaptive_rho_interval == 0) s->st.adaptive_rho_interval = 25;  // OSQP's timing-based choice, fixed
  s->max_batch = max_batch;
  s->device = device;
  const int N = cfg->horizon;
  const size_t ssz = (size_t)max_batch * f110::state_doubles(N, cfg->rate_rows, cfg->state_rows) * sizeof(double);
  e = cudaMalloc(&s->d_state, ssz);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_scratch, (size_t)(max_batch + 4) * f110::SCR_ROWS_ALLOC * (cfg->horizon < 32 ? 32 : (cfg->horizon < 64 ? 64 : 128)) * sizeof(double));
  if (e == cudaSuccess && cfg->horizon >= (ADMM_W2_GLOBAL_LEVELS > 0 ? 32 : 64) && !cfg->rate_rows)
    e = cudaMalloc(&s->d_mult, (size_t)max_batch * 28 * 128 * sizeof(double));
  const size_t wsz = (size_t)(f110::WORK_SLOTS + 1) * 2 * sizeof(int);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_work, wsz);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_work, 0, wsz, s->stream);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_solve, cudaEventDisableTiming);
  // the slots are cleared on the handle's own (non-blocking) stream and the clear is waited for: a memset on the legacy default
  // stream would not be ordered against solves on s->stream or on a caller's stream
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_state, 0, ssz, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) {
    f110_mpc_destroy(s);
    return cuda_fail(e, "f110_mpc_create: device allocation");
  }
  *out = s;
  return F110_OK;
}

void f110_mpc_destroy(f110_mpc_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->d_state);
  cudaFree(s->d_scratch);
  cudaFree(s->d_mult);
  cudaFree(s->d_work);
  cudaFree(s->d_recs); cudaFree(s->d_out);
  s->cyc.release();
  cudaFree(s->cyc_stage);
  if (s->cyc_pin) cudaFreeHost(s->cyc_pin);
  if (s->h_pin) cudaFreeHost(s->h_pin);
  if (s->stream) cudaStreamDestroy(s->stream);
  for (cudaGraphExec_t g : s->lat_graph) if (g) cudaGraphExecDestroy(g);
  if (s->stream2) cudaStreamDestroy(s->stream2);
  if (s->ev_tab) cudaEventDestroy(s->ev_tab);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  for (f110_cycle_lane& l : s->lane) {
    l.cyc.release();
    cudaFree(l.stage);
    if (l.pin_in) cudaFreeHost(l.pin_in);
    if (l.pin_out) cudaFreeHost(l.pin_out);
    if (l.ev_done) cudaEventDestroy(l.ev_done);
    if (l.ev_own) cudaEventDestroy(l.ev_own);
    if (l.ev_gather) cudaEventDestroy(l.ev_gather);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  if (s->ev_solve) cudaEventDestroy(s->ev_solve);
  if (s->gather_stream) cudaStreamDestroy(s->gather_stream);
  delete s;
}

int f110_mpc_reset(f110_mpc_solver* s) {
  if (!s) return fail(F110_ERR_ARG, "f110_mpc_reset: null solver");
  CUDA_TRY(cudaSetDevice(s->device));
  // Ordered after everything already queued on the handle's stream, and complete when this returns.  Work the caller queued on
  // its OWN streams (solve_device / cycle_device) is not waited for: synchronise those before calling reset.
  CUDA_TRY(cudaMemsetAsync(s->d_state, 0, (size_t)s->max_batch * f110::state_doubles(s->cfg.horizon, s->cfg.rate_rows, s->cfg.state_rows) * sizeof(double), s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return F110_OK;
}

int f110_mpc_last_launches(const f110_mpc_solver* s) { return s ? s->last_launches : 0; }

int f110_mpc_set_packed_output(f110_mpc_solver* s, double* d_packed) {
  if (!s) return fail(F110_ERR_ARG, "f110_mpc_set_packed_output: null solver");
  s->d_packed_next = d_packed;
  return F110_OK;
}

}  // extern "C"

// Solve `count` QPs that occupy warm-start / scratch slots slot0 .. slot0 + count - 1 of the handle.
int f110api::solve_device_range(f110_mpc_solver* s, int slot0, int count, const double* d_recs, int rec_stride, double* d_x, double* d_y,
                                double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_rho_updates, double* d_info,
                                void* cuda_stream) {
  if (!s || !d_recs) return fail(F110_ERR_ARG, "f110_mpc_solve_device: null solver or records");
  if (count < 0 || slot0 < 0 || slot0 + count > s->max_batch) return fail(F110_ERR_ARG, "f110_mpc_solve_device: count exceeds max_batch");
  if (rec_stride < f110_mpc_record_doubles(s->cfg.horizon)) return fail(F110_ERR_ARG, "f110_mpc_solve_device: record stride too small");
  if (count == 0) return F110_OK;
  f110::KParams p;
  p.N = s->cfg.horizon; p.B = count; p.stride = rec_stride; p.gap_mode = s->cfg.gap_mode;
  p.dt = s->cfg.dt; p.wheelbase = s->cfg.wheelbase;
  for (int i = 0; i < 3; ++i) p.Q[i] = s->cfg.q[i];
  for (int i = 0; i < 2; ++i) { p.R[i] = s->cfg.r[i]; p.u_des[i] = s->cfg.u_des[i]; p.u_min[i] = s->cfg.u_min[i]; p.u_max[i] = s->cfg.u_max[i]; }
  for (int i = 0; i < 2; ++i) p.qu[i] = -1.0 * s->cfg.r[i] * s->cfg.u_des[i];
  p.one_minus_alpha = 1.0 - s->st.alpha;
  p.rate_rows = s->cfg.rate_rows ? 1 : 0; p.rate_delta = s->cfg.rate_delta;
  p.state_rows = s->cfg.state_rows ? 1 : 0; p.state_lim = s->cfg.state_lim;
  p.rho0 = s->st.rho; p.sigma = s->st.sigma; p.alpha = s->st.alpha; p.eps_abs = s->st.eps_abs; p.eps_rel = s->st.eps_rel;
  p.eps_prim_inf = s->st.eps_prim_inf; p.eps_dual_inf = s->st.eps_dual_inf; p.adaptive_rho_tolerance = s->st.adaptive_rho_tolerance;
  p.max_iter = s->st.max_iter; p.check_termination = s->st.check_termination; p.scaling = s->st.scaling;
  p.adaptive_rho = s->st.adaptive_rho; p.adaptive_rho_interval = s->st.adaptive_rho_interval; p.warm_start = s->st.warm_start;
  p.recs = d_recs; p.x_out = d_x; p.y_out = d_y; p.u0_out = d_u0; p.status = d_status; p.iters = d_iters;
  p.rho_updates = d_rho_updates; p.info = d_info; p.packed = s->d_packed_next;
  s->d_packed_next = nullptr;
  p.done_flag = s->done_flag_next; p.done_seq = s->done_seq;
  s->done_flag_next = nullptr;
  const int T = s->cfg.horizon < 32 ? 32 : (s->cfg.horizon < 64 ? 64 : 128);   // threads (stage slots) per QP
  p.state = s->st.warm_start ? s->d_state + (size_t)slot0 * f110::state_doubles(s->cfg.horizon, s->cfg.rate_rows, s->cfg.state_rows) : nullptr;
  p.scratch = s->d_scratch + (size_t)slot0 * f110::SCR_ROWS_ALLOC * T;
  p.scratch_dummy = s->d_scratch + (size_t)s->max_batch * f110::SCR_ROWS_ALLOC * T;
  p.mult_global = s->d_mult ? s->d_mult + (size_t)slot0 * 28 * T : nullptr;
  // each launch gets its own counter pair (launches of one handle may overlap on different streams); the kernel leaves it zeroed.
  // The captured single-QP graph replays with a fixed pointer, so it owns the extra slot.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing((cudaStream_t)cuda_stream, &cap);
  p.work = s->d_work + 2 * (cap == cudaStreamCaptureStatusActive ? f110::WORK_SLOTS : (int)(s->work_seq++ % f110::WORK_SLOTS));
  CUDA_TRY(cudaSetDevice(s->device));
  int launched = 0;
  cudaError_t e = f110::launch_admm(p, (cudaStream_t)cuda_stream, &launched);
  if (e != cudaSuccess) return cuda_fail(e, "admm kernel launch");
  s->last_launches += launched;
  return F110_OK;
}

extern "C" {

int f110_mpc_solve_device(f110_mpc_solver* s, int count, const double* d_recs, int rec_stride, double* d_x, double* d_y,
                          double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_rho_updates, double* d_info,
                          void* cuda_stream) {
  if (s) s->last_launches = 0;
  return f110api::solve_device_range(s, 0, count, d_recs, rec_stride, d_x, d_y, d_u0, d_status, d_iters, d_rho_updates, d_info, cuda_stream);
}

int f110_mpc_solve_host(f110_mpc_solver* s, int count, const double* recs, int rec_stride, double* x, double* y, double* u0,
                        int32_t* status, int32_t* iters) {
  if (!s || !recs) return fail(F110_ERR_ARG, "f110_mpc_solve_host: null solver or records");
  if (count < 0 || count > s->max_batch) return fail(F110_ERR_ARG, "f110_mpc_solve_host: count exceeds max_batch");
  if (count == 0) return F110_OK;
  const int N = s->cfg.horizon, n = 5 * N + 3, m = f110::num_rows(N, s->cfg.rate_rows, s->cfg.state_rows);
  const int rd = f110_mpc_record_doubles(N);
  const int rdp = (rd + 1) & ~1;  // device stride: even, so every record is 16-byte aligned for the kernel's TMA staging
  if (rec_stride < rd) return fail(F110_ERR_ARG, "f110_mpc_solve_host: record stride too small");
  constexpr int kSmallBatch = 16;
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->d_recs) {
    const size_t B = s->max_batch > kSmallBatch ? s->max_batch : kSmallBatch;  // the compact small-batch layout needs kSmallBatch slots
    s->out_bytes = B * (2 * sizeof(double) + 2 * sizeof(int32_t)) + 16 + B * (size_t)(n + m) * sizeof(double);
    CUDA_TRY(cudaMalloc(&s->d_recs, B * rdp * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s->d_out, s->out_bytes));
    CUDA_TRY(cudaHostAlloc(&s->h_pin, kSmallBatch * (size_t)(rdp + 2 + 1 + n + m) * sizeof(double) + 64, cudaHostAllocMapped));   // (+64: completion flag of the single-QP path)
    std::memset(s->h_pin + kSmallBatch * (size_t)(rdp + 2 + 1 + n + m) * sizeof(double), 0, 64);
  }
  const bool small = count <= kSmallBatch;
  // small batches use a compact layout sized for kSmallBatch so the whole result is ONE device-to-host copy
  const size_t cap = small ? (size_t)kSmallBatch : (size_t)s->max_batch;
  const size_t o_status = cap * 2 * sizeof(double), o_iters = o_status + cap * sizeof(int32_t);
  const size_t o_x = (o_iters + cap * sizeof(int32_t) + 15) / 16 * 16, o_y = o_x + cap * n * sizeof(double);
  double* d_u0 = reinterpret_cast<double*>(s->d_out);
  int32_t* d_status = reinterpret_cast<int32_t*>(s->d_out + o_status);
  int32_t* d_iters = reinterpret_cast<int32_t*>(s->d_out + o_iters);
  double* d_x = reinterpret_cast<double*>(s->d_out + o_x);
  double* d_y = reinterpret_cast<double*>(s->d_out + o_y);
  if (small) {
    // latency path: records staged through pinned memory (true async DMA)
    double* hp = reinterpret_cast<double*>(s->h_pin);
    for (int b = 0; b < count; ++b) std::memcpy(hp + (size_t)b * rdp, recs + (size_t)b * rec_stride, rd * sizeof(double));
    // One QP at horizons 16..31 (the reference's own call pattern and horizon, mpc.cpp:69-143): no copy at all.  The kernel stages
    // the record from the pinned buffer itself (its TMA bulk copy reads mapped host memory), writes the results into the pinned
    // buffer and raises a flag there; the host polls the flag instead of waiting on the stream.  One kernel launch is the whole
    // call.  F110_LATENCY_PATH=graph selects the captured copy-in / solve / copy-out graph below instead (A/B measurements).
    static const bool direct_ok = [] { const char* e = std::getenv("F110_LATENCY_PATH"); return !(e && e[0] == 'g'); }();
    if (direct_ok && count == 1 && !s->cfg.rate_rows && !s->cfg.state_rows && N >= 16 && N <= 31 && !s->d_packed_next) {
      unsigned char* ho = s->h_pin + (size_t)kSmallBatch * rdp * sizeof(double);
      volatile int32_t* flag = reinterpret_cast<volatile int32_t*>(s->h_pin + kSmallBatch * (size_t)(rdp + 2 + 1 + n + m) * sizeof(double));
      unsigned char* pin_d = nullptr;   // device view of the pinned block
      CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&pin_d), s->h_pin, 0));
      unsigned char* ho_d = pin_d + (ho - s->h_pin);
      const int32_t seq = ++s->done_seq;
      s->done_flag_next = reinterpret_cast<int32_t*>(pin_d + (reinterpret_cast<volatile unsigned char*>(flag) - s->h_pin));
      s->last_launches = 0;
      const int rc1 = f110api::solve_device_range(s, 0, 1, reinterpret_cast<const double*>(pin_d), rdp, x ? reinterpret_cast<double*>(ho_d + o_x) : nullptr,
                                                  y ? reinterpret_cast<double*>(ho_d + o_y) : nullptr, reinterpret_cast<double*>(ho_d),
                                                  reinterpret_cast<int32_t*>(ho_d + o_status), reinterpret_cast<int32_t*>(ho_d + o_iters), nullptr,
                                                  nullptr, s->stream);
      if (rc1) { s->done_flag_next = nullptr; return rc1; }
      // poll the flag; every so often ask the stream as well, so that a failed launch or a kernel without the flag (F110_NO_TMEM)
      // still ends the wait
      for (unsigned spins = 0; *flag != seq; ++spins) {
        if ((spins & 1023u) == 1023u) {
          const cudaError_t q = cudaStreamQuery(s->stream);
          if (q == cudaSuccess) break;
          if (q != cudaErrorNotReady) return cuda_fail(q, "f110_mpc_solve_host: single-QP solve");
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      if (u0) std::memcpy(u0, ho, 2 * sizeof(double));
      if (status) std::memcpy(status, ho + o_status, sizeof(int32_t));
      if (iters) std::memcpy(iters, ho + o_iters, sizeof(int32_t));
      if (x) std::memcpy(x, ho + o_x, (size_t)n * sizeof(double));
      if (y) std::memcpy(y, ho + o_y, (size_t)m * sizeof(double));
      return F110_OK;
    }
    if (count == 1 && !s->cfg.rate_rows && !s->cfg.state_rows && N <= 31 && !s->d_packed_next) {   // (longer horizons set a function attribute at launch: not captured;
                                                                          //  a packed-output request is per call and must not be frozen into the graph)
      // one QP (the reference's own call pattern, mpc.cpp:69-143): copy-in, solve and copy-out are replayed as one captured graph,
      // one driver call instead of three.  Every address and size in it is fixed for the handle's lifetime.
      unsigned char* ho = s->h_pin + (size_t)kSmallBatch * rdp * sizeof(double);
      const int shape = y ? 2 : (x ? 1 : 0);
      if (!s->lat_graph[shape]) {
        const size_t bytes = y ? o_y + (size_t)m * sizeof(double) : (x ? o_x + (size_t)n * sizeof(double) : o_x);
        cudaGraph_t g = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t ce = cudaMemcpyAsync(s->d_recs, hp, (size_t)rdp * sizeof(double), cudaMemcpyHostToDevice, s->stream);
        int rcg = F110_OK;
        if (ce == cudaSuccess)
          rcg = f110_mpc_solve_device(s, 1, s->d_recs, rdp, shape >= 1 ? d_x : nullptr, shape == 2 ? d_y : nullptr, d_u0, d_status, d_iters, nullptr,
                                      nullptr, s->stream);
        if (ce == cudaSuccess && rcg == F110_OK) ce = cudaMemcpyAsync(ho, s->d_out, bytes, cudaMemcpyDeviceToHost, s->stream);
        const cudaError_t ee = cudaStreamEndCapture(s->stream, &g);
        if (rcg != F110_OK) { if (g) cudaGraphDestroy(g); return rcg; }
        if (ce != cudaSuccess || ee != cudaSuccess) { if (g) cudaGraphDestroy(g); return cuda_fail(ce != cudaSuccess ? ce : ee, "f110_mpc_solve_host: graph capture"); }
        ce = cudaGraphInstantiate(&s->lat_graph[shape], g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return cuda_fail(ce, "f110_mpc_solve_host: graph instantiate");
      }
      s->last_launches = 1;
      CUDA_TRY(cudaGraphLaunch(s->lat_graph[shape], s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      if (u0) std::memcpy(u0, ho, 2 * sizeof(double));
      if (status) std::memcpy(status, ho + o_status, sizeof(int32_t));
      if (iters) std::memcpy(iters, ho + o_iters, sizeof(int32_t));
      if (x) std::memcpy(x, ho + o_x, (size_t)n * sizeof(double));
      if (y) std::memcpy(y, ho + o_y, (size_t)m * sizeof(double));
      return F110_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_recs, hp, (size_t)count * rdp * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  } else {
    CUDA_TRY(cudaMemcpy2DAsync(s->d_recs, rdp * sizeof(double), recs, (size_t)rec_stride * sizeof(double), rd * sizeof(double), count,
                               cudaMemcpyHostToDevice, s->stream));
  }
  int rc = f110_mpc_solve_device(s, count, s->d_recs, rdp, x ? d_x : nullptr, y ? d_y : nullptr, d_u0, d_status, d_iters, nullptr, nullptr,
                                 s->stream);
  if (rc) return rc;
  if (small) {
    unsigned char* ho = s->h_pin + (size_t)kSmallBatch * rdp * sizeof(double);
    const size_t bytes = y ? o_y + (size_t)count * m * sizeof(double) : (x ? o_x + (size_t)count * n * sizeof(double) : o_x);
    CUDA_TRY(cudaMemcpyAsync(ho, s->d_out, bytes, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (u0) std::memcpy(u0, ho, (size_t)count * 2 * sizeof(double));
    if (status) std::memcpy(status, ho + o_status, (size_t)count * sizeof(int32_t));
    if (iters) std::memcpy(iters, ho + o_iters, (size_t)count * sizeof(int32_t));
    if (x) std::memcpy(x, ho + o_x, (size_t)count * n * sizeof(double));
    if (y) std::memcpy(y, ho + o_y, (size_t)count * m * sizeof(double));
    return F110_OK;
  }
  if (x) CUDA_TRY(cudaMemcpyAsync(x, d_x, (size_t)count * n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (y) CUDA_TRY(cudaMemcpyAsync(y, d_y, (size_t)count * m * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (u0) CUDA_TRY(cudaMemcpyAsync(u0, d_u0, (size_t)count * 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, d_status, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  if (iters) CUDA_TRY(cudaMemcpyAsync(iters, d_iters, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return F110_OK;
}

void f110_cycle_default_config(f110_cycle_config* c) {
  c->n_beams = 1080;                       // f1tenth simulator convention
  c->angle_min = -2.35f; c->angle_increment = 4.7f / 1079; c->angle_max = c->angle_min + 1079 * c->angle_increment;
  c->occ_size = 10; c->occ_discrete = 0.1f; c->occ_dilation = 0.15f;
  c->follow_gap_thresh = 3.0f; c->fov_divider = 1.5f; c->buffer = 3.0f;
  c->lookahead = 2.5f;
  c->use_half_spaces = 1;
  c->qp_mode = 0;
  c->reserved = 0;
  c->v_lin = 4.5;
}

}  // extern "C"

// Argument checks + (re)allocation of the per-scene device scratch `c` (grown on demand; the cycle entries each own one).
int f110api::cycle_prepare(f110_mpc_solver* s, f110_cycle_scratch& c, const f110_cycle_config* cc, int scenes, int paths, int samples, int n_wp,
                           const double* d_table_xy) {
  if (cc->qp_mode < 0 || cc->qp_mode > 2) return fail(F110_ERR_ARG, "f110_cycle_device: bad qp_mode");
  const long long nqp = cc->qp_mode == 0 ? scenes : (long long)scenes * paths;
  if (scenes < 0 || nqp > s->max_batch) return fail(F110_ERR_ARG, "f110_cycle_device: QP count exceeds max_batch");
  if (paths <= 0 || samples <= 0 || n_wp <= 0 || cc->n_beams <= 0) return fail(F110_ERR_ARG, "f110_cycle_device: bad sizes");
  if (!(cc->occ_discrete > 0.f) || 2.f * cc->occ_dilation / cc->occ_discrete + 1.f > 8.f)
    return fail(F110_ERR_UNSUPPORTED, "f110_cycle_device: more than 8 dilation stamps per axis");
  if (n_wp > 1500) return fail(F110_ERR_UNSUPPORTED, "f110_cycle_device: more than 1500 raceline waypoints (shared-memory staging)");
  if (reinterpret_cast<uintptr_t>(d_table_xy) % 16) return fail(F110_ERR_ARG, "f110_cycle_device: table_xy must be 16-byte aligned");
  CUDA_TRY(cudaSetDevice(s->device));
  const int blocks = (int)(cc->occ_size / cc->occ_discrete);           // occupancy_grid.cpp:9
  const int N = s->cfg.horizon, rd = (f110_mpc_record_doubles(N) + 1) & ~1;  // even stride: records stay 16-byte aligned (TMA staging)
  if (c.blocks != blocks || c.paths < paths || c.cap_scenes < (size_t)scenes || c.cap_qps < (size_t)nqp) {
    // (the buffers may still be read by queued work of an earlier call: cudaFree waits for the device)
    const size_t S = (size_t)scenes > c.cap_scenes ? (size_t)scenes : c.cap_scenes;
    const size_t Q = (size_t)nqp > c.cap_qps ? (size_t)nqp : c.cap_qps;
    const int P = paths > c.paths ? paths : c.paths;
    c.release();
    CUDA_TRY(cudaMalloc(&c.grid, S * blocks * blocks * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.offset, S * 2 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.endw, S * P * 2 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.rot, S * 4 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.pose_xy, S * 2 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.state3, S * 3 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.l1l2, S * 6 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.recs, Q * rd * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.valid, S * P));
    CUDA_TRY(cudaMalloc(&c.free_cnt, S * P * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&c.gap, S * 2 * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&c.best_global, S * sizeof(int32_t)));
    c.blocks = blocks; c.paths = P; c.cap_scenes = S; c.cap_qps = Q;
  }
  return F110_OK;
}

// The 5 kernels of one cycle for scenes scene0 .. scene0 + scenes - 1.  The caller's pointers are already offset to scene0; the
// handle's per-scene scratch and the solver's warm-start slots are offset here, so ranges of one batch can run on different streams.
int f110api::cycle_device_range(f110_mpc_solver* s, f110_cycle_scratch& c, const f110_cycle_config* cc, int scene0, int scenes,
                                const double* d_pose7, const float* d_ranges, const double* d_prev_steer, const double* d_table_xy, int paths,
                                int samples, const float* d_wp_xy, int n_wp, double* d_u0, int32_t* d_status, int32_t* d_iters,
                                int32_t* d_chosen, uint8_t* d_valid, cudaStream_t st, cudaEvent_t wait_before_solve) {
  if (scenes == 0) return F110_OK;
  const int blocks = (int)(cc->occ_size / cc->occ_discrete);
  const int N = s->cfg.horizon, rd = (f110_mpc_record_doubles(N) + 1) & ~1;
  const size_t s0 = (size_t)scene0;
  const size_t q0 = cc->qp_mode == 0 ? s0 : s0 * paths;                 // first QP slot of the range
  const long long nqp = cc->qp_mode == 0 ? scenes : (long long)scenes * paths;
  float* grid = c.grid + s0 * blocks * blocks;
  float* offset = c.offset + s0 * 2;
  float* endw = c.endw + s0 * paths * 2;
  double* rot = c.rot + s0 * 4;
  double* pose_xy = c.pose_xy + s0 * 2;
  double* l1l2 = cc->use_half_spaces ? c.l1l2 + s0 * 6 : nullptr;
  double* recs = c.recs + q0 * rd;
  // beam count from the float expression of occupancy_grid.cpp:66 / constraints.cpp:118
  const int num_scans = (int)((cc->angle_max - cc->angle_min) / cc->angle_increment + 1);
  uint8_t* valid = d_valid ? d_valid : c.valid + s0 * paths;
  cudaError_t e = f110::launch_scene_prep(scenes, blocks, cc->occ_discrete, cc->occ_dilation, cc->n_beams, num_scans, cc->angle_min,
                                          cc->angle_increment, cc->follow_gap_thresh, cc->fov_divider, cc->buffer, d_pose7, d_ranges, grid,
                                          offset, rot, pose_xy, l1l2, c.gap + s0 * 2, st);
  if (e == cudaSuccess) e = f110::launch_collision(scenes, paths, samples, blocks, cc->occ_discrete, grid, offset, rot, pose_xy,
                                                   d_table_xy, valid, c.free_cnt + s0 * paths, endw, st);
  if (e == cudaSuccess) e = f110::launch_select(scenes, paths, n_wp, cc->lookahead, d_pose7, d_wp_xy, valid, endw, d_chosen, c.best_global + s0, st);
  if (e == cudaSuccess) e = f110::launch_build_records(scenes, paths, samples, N, rd, cc->qp_mode, cc->v_lin, d_pose7, rot, valid, d_chosen,
                                                       d_table_xy, d_prev_steer, l1l2, recs, st);
  if (e != cudaSuccess) return cuda_fail(e, "f110_cycle_device: kernel launch");
  if (wait_before_solve) CUDA_TRY(cudaStreamWaitEvent(st, wait_before_solve, 0));
  const int rc = solve_device_range(s, (int)q0, (int)nqp, recs, rd, nullptr, nullptr, d_u0, d_status, d_iters, nullptr, nullptr, st);
  if (rc == F110_OK && s->ev_solve) cudaEventRecord(s->ev_solve, st);   // the next cycle's solve (any stream of this handle) waits for this one
  s->last_launches += 4;
  return rc;
}

// FNV-1a over the constant tables (20 KB): cheaper than two more copies per cycle
unsigned long long f110api::table_hash(const double* table_xy, size_t n_tab, const float* wp_xy, size_t n_wpb, int paths, int samples, int n_wp) {
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&h](const void* ptr, size_t n) {
    const unsigned long long* w = static_cast<const unsigned long long*>(ptr);
    for (size_t i = 0; i < n / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
  };
  mix(table_xy, n_tab); mix(wp_xy, n_wpb);
  h ^= (unsigned long long)paths * 1315423911ull + (unsigned long long)samples * 2654435761ull + (unsigned long long)n_wp;
  return h;
}

extern "C" {

int f110_cycle_device(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* d_pose7, const float* d_ranges,
                      const double* d_prev_steer, const double* d_table_xy, int paths, int samples, const float* d_wp_xy, int n_wp,
                      double* d_u0, int32_t* d_status, int32_t* d_iters, int32_t* d_chosen, uint8_t* d_valid, void* cuda_stream) {
  if (!s || !cc || !d_pose7 || !d_ranges || !d_table_xy || !d_wp_xy || !d_chosen) return fail(F110_ERR_ARG, "f110_cycle_device: null argument");
  const int rc = f110api::cycle_prepare(s, s->cyc, cc, scenes, paths, samples, n_wp, d_table_xy);
  if (rc != F110_OK) return rc;
  s->last_launches = 0;
  return f110api::cycle_device_range(s, s->cyc, cc, 0, scenes, d_pose7, d_ranges, d_prev_steer, d_table_xy, paths, samples, d_wp_xy, n_wp, d_u0, d_status,
                            d_iters, d_chosen, d_valid, (cudaStream_t)cuda_stream, nullptr);
}

int f110_cycle_host(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* pose7, const float* ranges,
                    const double* prev_steer, const double* table_xy, int paths, int samples, const float* wp_xy, int n_wp, double* u0,
                    int32_t* status, int32_t* iters, int32_t* chosen, uint8_t* valid) {
  if (!s || !cc || !pose7 || !ranges || !table_xy || !wp_xy) return fail(F110_ERR_ARG, "f110_cycle_host: null argument");
  if (scenes <= 0) return scenes == 0 ? F110_OK : fail(F110_ERR_ARG, "f110_cycle_host: negative scene count");
  CUDA_TRY(cudaSetDevice(s->device));
  const long long nqp = cc->qp_mode == 0 ? scenes : (long long)scenes * paths;
  if (nqp > s->max_batch) return fail(F110_ERR_ARG, "f110_cycle_host: QP count exceeds max_batch");
  // One staging block on the device (grown on demand): [table | waypoints | inputs | outputs].  The outputs are contiguous so
  // they come back in ONE copy (into a pinned mirror, then scattered to the caller's arrays).  The mini-path table and the
  // raceline are start-up constants in the reference (project.cpp:34-37): they are uploaded only when their bytes change.
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t n_tab = (size_t)paths * samples * 2 * sizeof(double), n_wpb = (size_t)n_wp * 2 * sizeof(float);
  const size_t b_pose = up((size_t)scenes * 7 * sizeof(double)), b_rng = up((size_t)scenes * cc->n_beams * sizeof(float));
  const size_t b_prev = up((size_t)scenes * sizeof(double)), b_tab = up(n_tab), b_wp = up(n_wpb);
  const size_t o_u0 = 0, o_st = o_u0 + up((size_t)nqp * 2 * sizeof(double)), o_it = o_st + up((size_t)nqp * sizeof(int32_t));
  const size_t o_ch = o_it + up((size_t)nqp * sizeof(int32_t)), o_val = o_ch + up((size_t)scenes * sizeof(int32_t));
  const size_t b_out = o_val + up((size_t)scenes * paths);
  const size_t total = b_pose + b_rng + b_prev + b_tab + b_wp + b_out;
  if (total > s->cyc_stage_bytes || b_out > s->cyc_pin_bytes) {
    cudaFree(s->cyc_stage); s->cyc_stage = nullptr; s->cyc_stage_bytes = 0;
    if (s->cyc_pin) cudaFreeHost(s->cyc_pin);
    s->cyc_pin = nullptr; s->cyc_pin_bytes = 0;
    s->cyc_tab_hash = 0;
    CUDA_TRY(cudaMalloc(&s->cyc_stage, total));
    s->cyc_stage_bytes = total;
    CUDA_TRY(cudaHostAlloc(&s->cyc_pin, b_out, cudaHostAllocDefault));
    s->cyc_pin_bytes = b_out;
  }
  // The constant tables come FIRST: their device addresses depend only on (paths, samples, n_wp), which the content hash covers,
  // so a later call with fewer scenes or beams on the same handle finds them where they were uploaded.
  unsigned char* q = s->cyc_stage;
  double* d_tab = (double*)q; q += b_tab;
  float* d_wp = (float*)q; q += b_wp;
  double* d_pose = (double*)q; q += b_pose;
  float* d_rng = (float*)q; q += b_rng;
  double* d_prev = (double*)q; q += b_prev;
  unsigned char* d_out = q;
  cudaStream_t st = s->stream;
  if (!s->stream2) {
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_tab, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  }
  const unsigned long long h = f110api::table_hash(table_xy, n_tab, wp_xy, n_wpb, paths, samples, n_wp);
  if (h != s->cyc_tab_hash) {
    CUDA_TRY(cudaMemcpyAsync(d_tab, table_xy, n_tab, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_wp, wp_xy, n_wpb, cudaMemcpyHostToDevice, st));
    s->cyc_tab_hash = h;
  }
  int rc = f110api::cycle_prepare(s, s->cyc, cc, scenes, paths, samples, n_wp, d_tab);
  if (rc != F110_OK) return rc;
  // Scenes are independent, so a large batch is pipelined as two halves over two streams: the copies and the small kernels of the
  // second half run under the solve of the first, and the second solve fills the first one's tail.  Measured at 205 scenes x 20
  // QPs: 1 / 2 / 4 / 8 chunks = 525 / 503 / 555 / 731 us per call — the solve kernel owns every register file while it runs, so
  // finer chunks only add tails and launches.  F110_CYCLE_CHUNKS overrides (tuning).
  int chunks = scenes >= 64 ? 2 : 1;
  if (const char* ev = std::getenv("F110_CYCLE_CHUNKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 16) chunks = v < scenes ? v : scenes; }
  CUDA_TRY(cudaEventRecord(s->ev_tab, st));
  CUDA_TRY(cudaStreamWaitEvent(s->stream2, s->ev_tab, 0));   // stream2 needs the tables and must not overtake the previous call
  s->last_launches = 0;
  for (int ch = 0; ch < chunks; ++ch) {
    const int a = (int)((long long)scenes * ch / chunks), b = (int)((long long)scenes * (ch + 1) / chunks);
    const size_t q0 = cc->qp_mode == 0 ? (size_t)a : (size_t)a * paths;
    cudaStream_t cs = (ch & 1) ? s->stream2 : st;
    CUDA_TRY(cudaMemcpyAsync(d_pose + (size_t)a * 7, pose7 + (size_t)a * 7, (size_t)(b - a) * 7 * sizeof(double), cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaMemcpyAsync(d_rng + (size_t)a * cc->n_beams, ranges + (size_t)a * cc->n_beams, (size_t)(b - a) * cc->n_beams * sizeof(float),
                             cudaMemcpyHostToDevice, cs));
    if (prev_steer) CUDA_TRY(cudaMemcpyAsync(d_prev + a, prev_steer + a, (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, cs));
    rc = f110api::cycle_device_range(s, s->cyc, cc, a, b - a, d_pose + (size_t)a * 7, d_rng + (size_t)a * cc->n_beams, prev_steer ? d_prev + a : nullptr, d_tab,
                            paths, samples, d_wp, n_wp, (double*)(d_out + o_u0) + 2 * q0, (int32_t*)(d_out + o_st) + q0,
                            (int32_t*)(d_out + o_it) + q0, (int32_t*)(d_out + o_ch) + a, d_out + o_val + (size_t)a * paths, cs, nullptr);
    if (rc != F110_OK) { cudaStreamSynchronize(st); cudaStreamSynchronize(s->stream2); return rc; }
  }
  if (chunks > 1) {
    CUDA_TRY(cudaEventRecord(s->ev_join, s->stream2));
    CUDA_TRY(cudaStreamWaitEvent(st, s->ev_join, 0));
  }
  CUDA_TRY(cudaMemcpyAsync(s->cyc_pin, d_out, b_out, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const unsigned char* ho = s->cyc_pin;
  if (u0) std::memcpy(u0, ho + o_u0, (size_t)nqp * 2 * sizeof(double));
  if (status) std::memcpy(status, ho + o_st, (size_t)nqp * sizeof(int32_t));
  if (iters) std::memcpy(iters, ho + o_it, (size_t)nqp * sizeof(int32_t));
  if (chosen) std::memcpy(chosen, ho + o_ch, (size_t)scenes * sizeof(int32_t));
  if (valid) std::memcpy(valid, ho + o_val, (size_t)scenes * paths);
  return F110_OK;
}

int f110_cycle_buffers(f110_mpc_solver* s, float** d_grid, float** d_offset, double** d_l1l2, double** d_recs, int32_t** d_best_global) {
  if (!s) return fail(F110_ERR_ARG, "f110_cycle_buffers: null solver");
  if (d_grid) *d_grid = s->cyc.grid;
  if (d_offset) *d_offset = s->cyc.offset;
  if (d_l1l2) *d_l1l2 = s->cyc.l1l2;
  if (d_recs) *d_recs = s->cyc.recs;
  if (d_best_global) *d_best_global = s->cyc.best_global;
  return F110_OK;
}

int f110_collision_check_device(int scenes, int paths, int samples, int blocks, float discrete, const float* d_grid,
                                const float* d_offset, const double* d_rot, const double* d_pose_xy, const double* d_table_xy,
                                uint8_t* d_valid, int32_t* d_free_count, float* d_end_world, void* cuda_stream) {
  if (scenes < 0 || paths <= 0 || samples <= 0 || blocks <= 0) return fail(F110_ERR_ARG, "f110_collision_check: bad sizes");
  if (!d_grid || !d_offset || !d_rot || !d_pose_xy || !d_table_xy || !d_valid || !d_free_count || !d_end_world)
    return fail(F110_ERR_ARG, "f110_collision_check: null buffer");
  if (reinterpret_cast<uintptr_t>(d_table_xy) % 16) return fail(F110_ERR_ARG, "f110_collision_check: table_xy must be 16-byte aligned");
  cudaError_t e = f110::launch_collision(scenes, paths, samples, blocks, discrete, d_grid, d_offset, d_rot, d_pose_xy, d_table_xy,
                                         d_valid, d_free_count, d_end_world, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "collision kernel launch");
  return F110_OK;
}

// Device staging for the host-buffer collision entry: grown on demand, kept for the life of the process so a
// per-cycle caller pays no cudaMalloc.  One cache per host thread (and device).
namespace {
struct CollisionCache {
  int device = -1;
  size_t cap_grid = 0, cap_scene = 0, cap_tab = 0, cap_out = 0;
  float *grid = nullptr, *off = nullptr, *endw = nullptr;
  double *rot = nullptr, *pose = nullptr, *tab = nullptr;
  uint8_t* valid = nullptr;
  int32_t* free_cnt = nullptr;
  cudaStream_t stream = nullptr;
  void release() {
    cudaFree(grid); cudaFree(off); cudaFree(endw); cudaFree(rot); cudaFree(pose); cudaFree(tab); cudaFree(valid); cudaFree(free_cnt);
    grid = off = endw = nullptr; rot = pose = tab = nullptr; valid = nullptr; free_cnt = nullptr;
    cap_grid = cap_scene = cap_tab = cap_out = 0;
  }
};
thread_local CollisionCache g_cc;
}  // namespace

int f110_collision_check_host(int scenes, int paths, int samples, int blocks, float discrete, const float* grid, const float* offset,
                              const double* rot, const double* pose_xy, const double* table_xy, uint8_t* valid,
                              int32_t* free_count, float* end_world, int device) {
  if (scenes < 0 || paths <= 0 || samples <= 0 || blocks <= 0) return fail(F110_ERR_ARG, "f110_collision_check: bad sizes");
  if (!grid || !offset || !rot || !pose_xy || !table_xy || !valid || !free_count || !end_world)
    return fail(F110_ERR_ARG, "f110_collision_check: null buffer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(F110_ERR_CUDA, "f110_collision_check_host: no CUDA device (this library has no CPU fallback)");
  if (scenes == 0) return F110_OK;
  CUDA_TRY(cudaSetDevice(device));
  CollisionCache& c = g_cc;
  if (c.device != device) {
    if (c.device >= 0) { cudaSetDevice(c.device); c.release(); if (c.stream) cudaStreamDestroy(c.stream); c.stream = nullptr; cudaSetDevice(device); }
    c.device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  }
  const size_t gsz = (size_t)scenes * blocks * blocks, np = (size_t)scenes * paths, tsz = (size_t)paths * samples * 2;
  if (gsz > c.cap_grid) { cudaFree(c.grid); c.grid = nullptr; CUDA_TRY(cudaMalloc(&c.grid, gsz * sizeof(float))); c.cap_grid = gsz; }
  if ((size_t)scenes > c.cap_scene) {
    cudaFree(c.off); cudaFree(c.rot); cudaFree(c.pose); c.off = nullptr; c.rot = c.pose = nullptr;
    CUDA_TRY(cudaMalloc(&c.off, scenes * 2 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.rot, scenes * 4 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c.pose, scenes * 2 * sizeof(double)));
    c.cap_scene = scenes;
  }
  if (tsz > c.cap_tab) { cudaFree(c.tab); c.tab = nullptr; CUDA_TRY(cudaMalloc(&c.tab, tsz * sizeof(double))); c.cap_tab = tsz; }
  if (np > c.cap_out) {
    cudaFree(c.valid); cudaFree(c.free_cnt); cudaFree(c.endw); c.valid = nullptr; c.free_cnt = nullptr; c.endw = nullptr;
    CUDA_TRY(cudaMalloc(&c.valid, np)); CUDA_TRY(cudaMalloc(&c.free_cnt, np * sizeof(int32_t))); CUDA_TRY(cudaMalloc(&c.endw, np * 2 * sizeof(float)));
    c.cap_out = np;
  }
  cudaStream_t st = c.stream;
  CUDA_TRY(cudaMemcpyAsync(c.grid, grid, gsz * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c.off, offset, scenes * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c.rot, rot, scenes * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c.pose, pose_xy, scenes * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c.tab, table_xy, tsz * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = f110_collision_check_device(scenes, paths, samples, blocks, discrete, c.grid, c.off, c.rot, c.pose, c.tab, c.valid, c.free_cnt,
                                       c.endw, st);
  if (rc != F110_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(valid, c.valid, np, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(free_count, c.free_cnt, np * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(end_world, c.endw, np * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return F110_OK;
}

}  // extern "C"
