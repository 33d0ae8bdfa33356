// Asynchronous form of the host-buffer cycle entry (f110_cycle_submit / f110_cycle_wait) and the NVLink gather of the chosen
// controls across GPUs (f110_gather_*, f110_cycle_set_gather).
//
// Why: f110_cycle_host copies in, runs five kernels, copies out and synchronises — 96 us of every 410 us call was exposed copy
// and synchronisation (round-1 measurement).  The reference's own control loop has the same shape: OdomCallback runs a cycle
// while DriveLoop applies the previous cycle's result (project.cpp:160-191, 220-238).  Here two cycles (up to four,
// f110_cycle_set_depth) may be in flight per handle: cycle k+1's host-to-device copies and its small perception kernels run under
// cycle k's solve, cold-started solves of consecutive cycles overlap as well (the next cycle's CTAs take the SMs the previous
// solve's tail leaves idle), and the host only blocks in f110_cycle_wait.
//
// Gather: QPs are independent, so a multi-GPU batch needs nothing but a final gather of (u0, status, iters) to one GPU
// (SURVEY.md section 8e).  The solve kernel writes that packed row itself; with a gather ring attached, the row's destination is
// the root GPU's memory (mapped into the peers through CUDA IPC or peer access), so the transfer is the kernel's own store over
// NVLink — no collective launch on the step.  A rank raises a per-rank flag on the root after its solve; the root's stream waits for
// all flags with a stream memory operation before it copies the gathered rows out.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>

#include "api_internal.h"

using f110api::cuda_fail;
using f110api::fail;

namespace {

__global__ void signal_kernel(int32_t* flag, int32_t value) {
  // everything the stream did before this launch (the solve kernel's peer stores included) is complete; publish at system scope
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// spin fallback for the root when the driver has no stream wait-value entry point; gives up after ~2 s worth of polls
__global__ void wait_kernel(const int32_t* flags, int n, int skip, int32_t value) {
  for (int r = 0; r < n; ++r) {
    if (r == skip) continue;
    for (long long spin = 0; spin < (1ll << 26); ++spin) {
      int32_t v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if (v >= value) break;
      __nanosleep(64);
    }
  }
}

typedef CUresult (*wait_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
wait_value_fn wait_value_entry() {
  static wait_value_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<wait_value_fn>(p);
  }();
  return fn;
}

// cuStreamWriteValue32: the flag is raised by the stream itself once everything before it (the solve kernel's peer stores
// included) is complete.  Unlike a one-thread kernel it needs no SM slot — with overlapping solves every slot is held by a
// persistent CTA of the next cycle, and a signal kernel would queue behind them.
typedef CUresult (*write_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
write_value_fn write_value_entry() {
  static write_value_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    const char* e = std::getenv("F110_SIGNAL_KERNEL");   // A/B measurements: force the one-thread kernel
    if (e && e[0] == '1') p = nullptr;
    return reinterpret_cast<write_value_fn>(p);
  }();
  return fn;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

size_t up256(size_t v) { return (v + 255) / 256 * 256; }
// flags[slots][world] at the head of the ring allocation: one flag per (slot, rank), so cycles may complete in any order
size_t flag_bytes(int world, int slots) { return up256((size_t)slots * world * sizeof(int32_t)); }

}  // namespace

cudaError_t f110api::launch_signal(cudaStream_t st, int32_t* flag, int32_t value) {
  if (write_value_fn fn = write_value_entry()) {
    if (fn((CUstream)st, (CUdeviceptr)(uintptr_t)flag, (cuuint32_t)value, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS) return cudaSuccess;
  }
  signal_kernel<<<1, 1, 0, st>>>(flag, value);
  return cudaGetLastError();
}

extern "C" {

int f110_gather_bytes(int world, int rows_per_rank, int slots, size_t* bytes) {
  if (world < 1 || world > 64 || rows_per_rank < 1 || slots < 1 || !bytes) return fail(F110_ERR_ARG, "f110_gather_bytes: bad argument");
  *bytes = flag_bytes(world, slots) + (size_t)slots * world * rows_per_rank * 4 * sizeof(double);
  return F110_OK;
}

int f110_gather_create(int device, int world, int rows_per_rank, int slots, void** d_ring, unsigned char* ipc_handle64) {
  size_t bytes = 0;
  if (!d_ring) return fail(F110_ERR_ARG, "f110_gather_create: null argument");
  int rc = f110_gather_bytes(world, rows_per_rank, slots, &bytes);
  if (rc != F110_OK) return rc;
  CUDA_TRY(cudaSetDevice(device));
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && ipc_handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) std::memcpy(ipc_handle64, &h, 64);
  }
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "f110_gather_create"); }
  *d_ring = p;
  return F110_OK;
}

int f110_gather_open(int device, const unsigned char* ipc_handle64, void** d_ring) {
  if (!ipc_handle64 || !d_ring) return fail(F110_ERR_ARG, "f110_gather_open: null argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle64, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ring = p;
  return F110_OK;
}

int f110_gather_close(int device, void* d_ring, int opened) {
  if (!d_ring) return F110_OK;
  CUDA_TRY(cudaSetDevice(device));
  if (opened) CUDA_TRY(cudaIpcCloseMemHandle(d_ring));
  else CUDA_TRY(cudaFree(d_ring));
  return F110_OK;
}

int f110_gather_slot(void* d_ring, int world, int rank, int rows_per_rank, int slots, long long seq, double** d_rows, int32_t** d_flag) {
  if (!d_ring || rank < 0 || rank >= world || slots < 1) return fail(F110_ERR_ARG, "f110_gather_slot: bad argument");
  unsigned char* base = static_cast<unsigned char*>(d_ring);
  if (d_rows) *d_rows = reinterpret_cast<double*>(base + flag_bytes(world, slots)) + ((size_t)(seq % slots) * world + rank) * rows_per_rank * 4;
  if (d_flag) *d_flag = reinterpret_cast<int32_t*>(base) + (size_t)(seq % slots) * world + rank;
  return F110_OK;
}

int f110_stream_signal(void* cuda_stream, int32_t* d_flag, int32_t value) {
  if (!d_flag) return fail(F110_ERR_ARG, "f110_stream_signal: null flag");
  cudaError_t e = f110api::launch_signal((cudaStream_t)cuda_stream, d_flag, value);
  if (e != cudaSuccess) return cuda_fail(e, "f110_stream_signal");
  return F110_OK;
}

int f110_stream_wait_flags(void* cuda_stream, const int32_t* d_flags, int n, int skip, int32_t value) {
  if (!d_flags || n < 1) return fail(F110_ERR_ARG, "f110_stream_wait_flags: bad argument");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (wait_value_fn fn = wait_value_entry()) {
    for (int r = 0; r < n; ++r) {
      if (r == skip) continue;
      const CUresult cr = fn((CUstream)st, (CUdeviceptr)(uintptr_t)(d_flags + r), (cuuint32_t)value, CU_STREAM_WAIT_VALUE_GEQ);
      if (cr != CUDA_SUCCESS) return fail(F110_ERR_CUDA, "f110_stream_wait_flags: cuStreamWaitValue32 failed (" + std::to_string((int)cr) + ")");
    }
    return F110_OK;
  }
  wait_kernel<<<1, 1, 0, st>>>(d_flags, n, skip, value);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "f110_stream_wait_flags");
  return F110_OK;
}

int f110_cycle_set_gather(f110_mpc_solver* s, void* d_ring, int world, int rank, int rows_per_rank, int slots) {
  if (!s) return fail(F110_ERR_ARG, "f110_cycle_set_gather: null solver");
  if (!d_ring) { s->gather = f110_mpc_solver::Gather(); return F110_OK; }
  if (world < 1 || rank < 0 || rank >= world || rows_per_rank < 1 || slots < 1) return fail(F110_ERR_ARG, "f110_cycle_set_gather: bad argument");
  unsigned char* base = static_cast<unsigned char*>(d_ring);
  s->gather.flags = reinterpret_cast<int32_t*>(base);
  s->gather.ring = reinterpret_cast<double*>(base + flag_bytes(world, slots));
  s->gather.world = world; s->gather.rank = rank; s->gather.rows = rows_per_rank; s->gather.slots = slots; s->gather.seq = 0;
  return F110_OK;
}

int f110_cycle_set_depth(f110_mpc_solver* s, int depth) {
  if (!s || depth < 1 || depth > f110_mpc_solver::kMaxLanes) return fail(F110_ERR_ARG, "f110_cycle_set_depth: depth must be 1.." + std::to_string(f110_mpc_solver::kMaxLanes));
  for (const f110_cycle_lane& l : s->lane)
    if (l.busy) return fail(F110_ERR_ARG, "f110_cycle_set_depth: cycles are in flight (wait for them first)");
  // tickets map to lanes by ticket % depth: restart the numbering on a multiple of every depth so lane 0 comes next
  s->next_ticket = (s->next_ticket + 11) / 12 * 12;
  s->depth = depth;
  return F110_OK;
}

int f110_cycle_submit(f110_mpc_solver* s, const f110_cycle_config* cc, int scenes, const double* pose7, const float* ranges,
                      const double* prev_steer, const double* table_xy, int paths, int samples, const float* wp_xy, int n_wp, int* ticket) {
  if (!s || !cc || !pose7 || !ranges || !table_xy || !wp_xy || !ticket) return fail(F110_ERR_ARG, "f110_cycle_submit: null argument");
  if (scenes <= 0) return fail(F110_ERR_ARG, "f110_cycle_submit: scene count must be positive");
  const long long nqp = cc->qp_mode == 0 ? scenes : (long long)scenes * paths;
  if (nqp > s->max_batch) return fail(F110_ERR_ARG, "f110_cycle_submit: QP count exceeds max_batch");
  auto& g = s->gather;
  if (g.ring && nqp > g.rows) return fail(F110_ERR_ARG, "f110_cycle_submit: QP count exceeds the gather ring's rows per rank");
  f110_cycle_lane& L = s->lane[s->next_ticket % s->depth];
  if (L.busy)
    return fail(F110_ERR_ARG, "f110_cycle_submit: " + std::to_string(s->depth) + " cycles are already in flight (call f110_cycle_wait on the oldest ticket)");
  CUDA_TRY(cudaSetDevice(s->device));
  if (!L.stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
  }
  // device staging [table | waypoints | inputs | outputs]; pinned mirrors for the outputs (and for pageable inputs)
  const size_t n_tab = (size_t)paths * samples * 2 * sizeof(double), n_wpb = (size_t)n_wp * 2 * sizeof(float);
  const size_t n_pose = (size_t)scenes * 7 * sizeof(double), n_rng = (size_t)scenes * cc->n_beams * sizeof(float), n_prev = (size_t)scenes * sizeof(double);
  const size_t b_tab = up256(n_tab), b_wp = up256(n_wpb), b_pose = up256(n_pose), b_rng = up256(n_rng), b_prev = up256(n_prev);
  L.o_u0 = 0; L.o_st = L.o_u0 + up256((size_t)nqp * 2 * sizeof(double)); L.o_it = L.o_st + up256((size_t)nqp * sizeof(int32_t));
  L.o_ch = L.o_it + up256((size_t)nqp * sizeof(int32_t)); L.o_val = L.o_ch + up256((size_t)scenes * sizeof(int32_t));
  L.b_out = L.o_val + up256((size_t)scenes * paths);
  L.gather_bytes = (g.ring && g.rank == 0) ? (size_t)g.world * g.rows * 4 * sizeof(double) : 0;
  const size_t total = b_tab + b_wp + b_pose + b_rng + b_prev + L.b_out;
  if (total > L.stage_bytes) {
    cudaFree(L.stage); L.stage = nullptr; L.stage_bytes = 0; L.tab_hash = 0;
    CUDA_TRY(cudaMalloc(&L.stage, total));
    L.stage_bytes = total;
  }
  if (L.b_out + L.gather_bytes > L.pin_out_bytes) {
    if (L.pin_out) cudaFreeHost(L.pin_out);
    L.pin_out = nullptr; L.pin_out_bytes = 0;
    CUDA_TRY(cudaHostAlloc(&L.pin_out, L.b_out + L.gather_bytes, cudaHostAllocDefault));
    L.pin_out_bytes = L.b_out + L.gather_bytes;
  }
  unsigned char* q = L.stage;
  double* d_tab = (double*)q; q += b_tab;
  float* d_wp = (float*)q; q += b_wp;
  double* d_pose = (double*)q; q += b_pose;
  float* d_rng = (float*)q; q += b_rng;
  double* d_prev = (double*)q; q += b_prev;
  unsigned char* d_out = q;
  cudaStream_t st = L.stream;
  const unsigned long long h = f110api::table_hash(table_xy, n_tab, wp_xy, n_wpb, paths, samples, n_wp);
  if (h != L.tab_hash) {   // start-up constants in the reference (project.cpp:34-37): uploaded when their bytes change
    CUDA_TRY(cudaMemcpyAsync(d_tab, table_xy, n_tab, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_wp, wp_xy, n_wpb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // the caller may reuse pageable table memory right after this call
    L.tab_hash = h;
  }
  // per-cycle inputs: straight from the caller's memory when it is pinned, through the lane's pinned staging otherwise
  const bool direct = is_pinned(pose7) && is_pinned(ranges) && (!prev_steer || is_pinned(prev_steer));
  const void *src_pose = pose7, *src_rng = ranges, *src_prev = prev_steer;
  if (!direct) {
    const size_t need = b_pose + b_rng + b_prev;
    if (need > L.pin_in_bytes) {
      if (L.pin_in) cudaFreeHost(L.pin_in);
      L.pin_in = nullptr; L.pin_in_bytes = 0;
      CUDA_TRY(cudaHostAlloc(&L.pin_in, need, cudaHostAllocDefault));
      L.pin_in_bytes = need;
    }
    std::memcpy(L.pin_in, pose7, n_pose);
    std::memcpy(L.pin_in + b_pose, ranges, n_rng);
    if (prev_steer) std::memcpy(L.pin_in + b_pose + b_rng, prev_steer, n_prev);
    src_pose = L.pin_in; src_rng = L.pin_in + b_pose; src_prev = prev_steer ? L.pin_in + b_pose + b_rng : nullptr;
  }
  CUDA_TRY(cudaMemcpyAsync(d_pose, src_pose, n_pose, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_rng, src_rng, n_rng, cudaMemcpyHostToDevice, st));
  if (prev_steer) CUDA_TRY(cudaMemcpyAsync(d_prev, src_prev, n_prev, cudaMemcpyHostToDevice, st));
  int rc = f110api::cycle_prepare(s, L.cyc, cc, scenes, paths, samples, n_wp, d_tab);
  if (rc != F110_OK) return rc;
  double* d_rows = nullptr;
  int32_t* d_flag = nullptr;
  if (g.ring) {
    d_rows = g.ring + ((size_t)(g.seq % g.slots) * g.world + g.rank) * g.rows * 4;
    d_flag = g.flags + (size_t)(g.seq % g.slots) * g.world + g.rank;
    s->d_packed_next = d_rows;   // the solve kernel stores this rank's rows on the root GPU
  }
  s->last_launches = 0;
  // Consecutive solves must run in submission order when they share state: the warm-start slots or scratch lines in global memory.
  // Otherwise they may overlap (the gather ring has one flag per slot and rank, so cycles may complete in any order).
  static const bool force_ordered = [] { const char* e = std::getenv("F110_CYCLE_ORDERED"); return e && e[0] == '1'; }();   // A/B measurements
  const bool ordered = force_ordered || s->st.warm_start || !f110::admm_state_on_chip(s->cfg.horizon, s->cfg.rate_rows, s->cfg.state_rows);
  rc = f110api::cycle_device_range(s, L.cyc, cc, 0, scenes, d_pose, d_rng, prev_steer ? d_prev : nullptr, d_tab, paths, samples, d_wp, n_wp,
                                   (double*)(d_out + L.o_u0), (int32_t*)(d_out + L.o_st), (int32_t*)(d_out + L.o_it), (int32_t*)(d_out + L.o_ch),
                                   d_out + L.o_val, st, ordered ? s->ev_solve : nullptr);
  if (rc != F110_OK) { s->d_packed_next = nullptr; cudaStreamSynchronize(st); return rc; }
  if (g.ring) {
    const int32_t delivered = (int32_t)(g.seq + 1);
    CUDA_TRY(f110api::launch_signal(st, d_flag, delivered));
    s->last_launches += 1;
    if (g.rank == 0) {
      // the root waits for the other ranks and copies the slot out on its own gather stream: a rank that lags must not hold up the
      // lane (the next cycle on this lane queues behind everything on its stream)
      if (!s->gather_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s->gather_stream, cudaStreamNonBlocking));
      if (!L.ev_gather) {
        CUDA_TRY(cudaEventCreateWithFlags(&L.ev_gather, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&L.ev_own, cudaEventDisableTiming));
      }
      CUDA_TRY(cudaEventRecord(L.ev_own, st));                                  // this rank's rows are in the slot
      CUDA_TRY(cudaStreamWaitEvent(s->gather_stream, L.ev_own, 0));
      rc = f110_stream_wait_flags(s->gather_stream, g.flags + (size_t)(g.seq % g.slots) * g.world, g.world, 0, delivered);
      if (rc != F110_OK) return rc;
      const double* slot = g.ring + (size_t)(g.seq % g.slots) * g.world * g.rows * 4;
      CUDA_TRY(cudaMemcpyAsync(L.pin_out + L.b_out, slot, L.gather_bytes, cudaMemcpyDeviceToHost, s->gather_stream));
      CUDA_TRY(cudaEventRecord(L.ev_gather, s->gather_stream));
      L.gather_pending = true;
    }
    ++g.seq;
  }
  CUDA_TRY(cudaMemcpyAsync(L.pin_out, d_out, L.b_out, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(L.ev_done, st));
  L.busy = true; L.scenes = scenes; L.paths = paths; L.nqp = nqp;
  L.ticket = s->next_ticket;
  *ticket = s->next_ticket++;
  return F110_OK;
}

int f110_cycle_wait(f110_mpc_solver* s, int ticket, double* u0, int32_t* status, int32_t* iters, int32_t* chosen, uint8_t* valid,
                    double* gathered) {
  if (!s || ticket < 0) return fail(F110_ERR_ARG, "f110_cycle_wait: bad argument");
  f110_cycle_lane& L = s->lane[ticket % s->depth];
  if (!L.busy || L.ticket != ticket) return fail(F110_ERR_ARG, "f110_cycle_wait: no such cycle in flight");
  CUDA_TRY(cudaSetDevice(s->device));
  cudaError_t e = cudaEventSynchronize(L.ev_done);
  if (e == cudaSuccess && L.gather_pending) e = cudaEventSynchronize(L.ev_gather);
  L.gather_pending = false;
  L.busy = false;
  if (e != cudaSuccess) return cuda_fail(e, "f110_cycle_wait");
  const unsigned char* ho = L.pin_out;
  if (u0) std::memcpy(u0, ho + L.o_u0, (size_t)L.nqp * 2 * sizeof(double));
  if (status) std::memcpy(status, ho + L.o_st, (size_t)L.nqp * sizeof(int32_t));
  if (iters) std::memcpy(iters, ho + L.o_it, (size_t)L.nqp * sizeof(int32_t));
  if (chosen) std::memcpy(chosen, ho + L.o_ch, (size_t)L.scenes * sizeof(int32_t));
  if (valid) std::memcpy(valid, ho + L.o_val, (size_t)L.scenes * L.paths);
  if (gathered) {
    if (!L.gather_bytes) return fail(F110_ERR_ARG, "f110_cycle_wait: gathered rows exist on the gather root (rank 0) only");
    std::memcpy(gathered, ho + L.b_out, L.gather_bytes);
  }
  return F110_OK;
}

int f110_cycle_gathered_view(f110_mpc_solver* s, int ticket, const double** rows, size_t* doubles) {
  if (!s || ticket < 0 || !rows) return fail(F110_ERR_ARG, "f110_cycle_gathered_view: bad argument");
  f110_cycle_lane& L = s->lane[ticket % s->depth];
  if (L.busy || L.ticket != ticket) return fail(F110_ERR_ARG, "f110_cycle_gathered_view: call f110_cycle_wait on this ticket first (and before the lane's next submit)");
  if (!L.gather_bytes) return fail(F110_ERR_ARG, "f110_cycle_gathered_view: gathered rows exist on the gather root (rank 0) only");
  *rows = reinterpret_cast<const double*>(L.pin_out + L.b_out);
  if (doubles) *doubles = L.gather_bytes / sizeof(double);
  return F110_OK;
}

}  // extern "C"
