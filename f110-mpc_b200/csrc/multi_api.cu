// Several GPUs behind one handle, from one process (f110_mpc_create_multi / f110_mpc_solve_multi_host).
//
// QPs are independent (the reference only ever holds one, mpc.cpp:69-143), so a batch is cut into contiguous shards of whole
// `unit`s (a unit = the QPs that belong together, e.g. the 140 lane x path QPs of one config-4 scenario),
// one shard per GPU.  No collective in the solve.  The final gather is the solve kernel's own store: every GPU's packed rows
// (u0_v, u0_steer, status, iters) go straight into one buffer on the first GPU through peer access over NVLink, and come back to the
// host in one copy.  Without peer access between two devices the rows are staged locally and moved with one peer copy per GPU.
#include <cstring>
#include <vector>

#include "api_internal.h"

using f110api::cuda_fail;
using f110api::fail;

struct f110_mpc_multi {
  struct Dev {
    int device = 0;
    f110_mpc_solver* solver = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    double* d_recs = nullptr;      // this GPU's shard of the records (even stride)
    double* d_local = nullptr;     // packed rows when the root's memory is not reachable by peer access
    bool peer = false;
  };
  std::vector<Dev> devs;
  int horizon = 0, max_batch = 0, per_dev = 0;
  double* d_gather = nullptr;      // [max_batch][4] on devs[0]
  double* h_gather = nullptr;      // pinned mirror
  double* h_recs = nullptr;        // pinned staging of the caller's records (even stride)
  int last_shard[64][2];           // [device] -> (first QP, count) of the last call
};

namespace {
void shard_units(int units, int world, int rank, int* lo, int* hi) {   // contiguous, balanced (sharding.py::shard_range)
  const int base = units / world, rem = units % world;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}
}  // namespace

extern "C" {

void f110_mpc_destroy_multi(f110_mpc_multi* m) {
  if (!m) return;
  for (auto& d : m->devs) {
    cudaSetDevice(d.device);
    if (d.solver) f110_mpc_destroy(d.solver);
    cudaFree(d.d_recs); cudaFree(d.d_local);
    if (d.done) cudaEventDestroy(d.done);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  if (!m->devs.empty()) { cudaSetDevice(m->devs[0].device); cudaFree(m->d_gather); }
  if (m->h_gather) cudaFreeHost(m->h_gather);
  if (m->h_recs) cudaFreeHost(m->h_recs);
  delete m;
}

int f110_mpc_create_multi(const f110_mpc_config* cfg, const f110_solver_settings* st, int max_batch, const int* devices, int n_devices,
                          f110_mpc_multi** out) {
  if (!cfg || !st || !devices || !out || max_batch <= 0 || n_devices < 1 || n_devices > 64)
    return fail(F110_ERR_ARG, "f110_mpc_create_multi: bad argument");
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return fail(F110_ERR_ARG, "f110_mpc_create_multi: a device is listed twice");
  f110_mpc_multi* m = new f110_mpc_multi();
  m->horizon = cfg->horizon; m->max_batch = max_batch;
  m->per_dev = max_batch;   // any shard fits: a batch of one unit lands on a single GPU
  const int rdp = (f110_mpc_record_doubles(cfg->horizon) + 1) & ~1;
  m->devs.resize(n_devices);
  int rc = F110_OK;
  for (int i = 0; i < n_devices && rc == F110_OK; ++i) {
    auto& d = m->devs[i];
    d.device = devices[i];
    rc = f110_mpc_create(cfg, st, m->per_dev, d.device, &d.solver);
    if (rc != F110_OK) break;
    cudaError_t e = cudaSetDevice(d.device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&d.d_recs, (size_t)m->per_dev * rdp * sizeof(double));
    if (e == cudaSuccess && i == 0) e = cudaMalloc(&m->d_gather, (size_t)max_batch * 4 * sizeof(double));
    if (e == cudaSuccess && i > 0) {
      int can = 0;
      e = cudaDeviceCanAccessPeer(&can, d.device, m->devs[0].device);
      if (e == cudaSuccess && can) {
        const cudaError_t pe = cudaDeviceEnablePeerAccess(m->devs[0].device, 0);
        if (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled) { d.peer = true; cudaGetLastError(); }
      }
      if (e == cudaSuccess && !d.peer) e = cudaMalloc(&d.d_local, (size_t)m->per_dev * 4 * sizeof(double));
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "f110_mpc_create_multi: device set-up");
  }
  if (rc == F110_OK) {
    cudaError_t e = cudaHostAlloc(&m->h_gather, (size_t)max_batch * 4 * sizeof(double), cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc(&m->h_recs, (size_t)max_batch * rdp * sizeof(double), cudaHostAllocPortable);
    if (e != cudaSuccess) rc = cuda_fail(e, "f110_mpc_create_multi: pinned staging");
  }
  if (rc != F110_OK) { f110_mpc_destroy_multi(m); return rc; }
  *out = m;
  return F110_OK;
}

int f110_mpc_multi_devices(const f110_mpc_multi* m) { return m ? (int)m->devs.size() : 0; }

int f110_mpc_multi_uses_peer_stores(const f110_mpc_multi* m, int index) {
  return (m && index >= 0 && index < (int)m->devs.size()) ? (index == 0 || m->devs[index].peer ? 1 : 0) : 0;
}

int f110_mpc_multi_last_shard(const f110_mpc_multi* m, int index, int* first, int* count) {
  if (!m || index < 0 || index >= (int)m->devs.size()) return fail(F110_ERR_ARG, "f110_mpc_multi_last_shard: bad argument");
  if (first) *first = m->last_shard[index][0];
  if (count) *count = m->last_shard[index][1];
  return F110_OK;
}

int f110_mpc_solve_multi_host(f110_mpc_multi* m, int count, int unit, const double* recs, int rec_stride, double* u0, int32_t* status,
                              int32_t* iters) {
  if (!m || !recs) return fail(F110_ERR_ARG, "f110_mpc_solve_multi_host: null argument");
  if (count < 0 || count > m->max_batch) return fail(F110_ERR_ARG, "f110_mpc_solve_multi_host: count exceeds max_batch");
  if (unit < 1 || count % unit) return fail(F110_ERR_ARG, "f110_mpc_solve_multi_host: count must be a multiple of the shard unit");
  const int rd = f110_mpc_record_doubles(m->horizon), rdp = (rd + 1) & ~1;
  if (rec_stride < rd) return fail(F110_ERR_ARG, "f110_mpc_solve_multi_host: record stride too small");
  if (count == 0) return F110_OK;
  const int world = (int)m->devs.size(), units = count / unit;
  // records into pinned memory at the device stride: each shard is then ONE asynchronous copy
  for (int b = 0; b < count; ++b) std::memcpy(m->h_recs + (size_t)b * rdp, recs + (size_t)b * rec_stride, rd * sizeof(double));
  auto& root = m->devs[0];
  for (int r = 0; r < world; ++r) {
    auto& d = m->devs[r];
    int ulo, uhi;
    shard_units(units, world, r, &ulo, &uhi);
    const int q0 = ulo * unit, n = (uhi - ulo) * unit;
    m->last_shard[r][0] = q0; m->last_shard[r][1] = n;
    if (n == 0) continue;
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaMemcpyAsync(d.d_recs, m->h_recs + (size_t)q0 * rdp, (size_t)n * rdp * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    double* rows = (r == 0 || d.peer) ? m->d_gather + (size_t)q0 * 4 : d.d_local;
    int rc = f110_mpc_set_packed_output(d.solver, rows);
    if (rc == F110_OK) rc = f110_mpc_solve_device(d.solver, n, d.d_recs, rdp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d.stream);
    if (rc != F110_OK) return rc;
    if (r > 0 && !d.peer)
      CUDA_TRY(cudaMemcpyPeerAsync(m->d_gather + (size_t)q0 * 4, root.device, d.d_local, d.device, (size_t)n * 4 * sizeof(double), d.stream));
    CUDA_TRY(cudaEventRecord(d.done, d.stream));
  }
  CUDA_TRY(cudaSetDevice(root.device));
  for (int r = 1; r < world; ++r)
    if (m->last_shard[r][1] > 0) CUDA_TRY(cudaStreamWaitEvent(root.stream, m->devs[r].done, 0));
  CUDA_TRY(cudaMemcpyAsync(m->h_gather, m->d_gather, (size_t)count * 4 * sizeof(double), cudaMemcpyDeviceToHost, root.stream));
  CUDA_TRY(cudaStreamSynchronize(root.stream));
  for (int b = 0; b < count; ++b) {
    const double* row = m->h_gather + (size_t)b * 4;
    if (u0) { u0[2 * b] = row[0]; u0[2 * b + 1] = row[1]; }
    if (status) status[b] = (int32_t)row[2];
    if (iters) iters[b] = (int32_t)row[3];
  }
  return F110_OK;
}

}  // extern "C"
