// Micro-benchmark: is tensor memory (TMEM) usable as a per-lane store for the PCR multipliers of the ADMM kernel?
//
// The round-1 kernel is bound by the LSU data pipe (shared-memory loads of the multipliers + shuffles, 254 wavefronts per
// warp-iteration).  TMEM is read with tcgen05.ld, a different datapath.  This program measures, per SM:
//   T  tcgen05.ld.32x32b.x4 of 88 doubles per lane per "iteration" (44 loads), consumed by DFMAs
//   L  the same volume as conflict-free LDS.128
//   D  142 further DFMAs in 4 independent chains (the rest of an ADMM iteration)
//   S  78 SHFL.32
// in the combinations that model the current kernel (L+D+S) and the proposed one (T+D+S), at 4 and 8 warps per SM.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t addr, double& a, double& b) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
  a = __hiloint2double((int)r1, (int)r0);
  b = __hiloint2double((int)r3, (int)r2);
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, double a, double b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__double2loint(a)), "r"(__double2hiint(a)),
               "r"(__double2loint(b)), "r"(__double2hiint(b))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int PAIRS = 44;   // double2 per lane per iteration (88 doubles = the N=30 multiplier set)
constexpr int LEVELS = 4;   // loads are issued in batches of 11 pairs, like one PCR level, then waited for

template <bool T, bool L, bool D, bool S>
__global__ void __launch_bounds__(128, 2) kern(int iters, int cols, double* out, long long* cyc, int* ok) {
  extern __shared__ __align__(16) double sm[];
  __shared__ uint32_t tbase_s;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w == 0) tmem_alloc(&tbase_s, (uint32_t)cols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s + ((uint32_t)(32 * w) << 16);
  // fill: pair j of this lane = (1e-3*(j+1), 1e-3*(j+1) + lane*1e-6)
  double2* smp = reinterpret_cast<double2*>(sm) + threadIdx.x;   // pair-major, thread fastest: conflict-free LDS.128
  for (int j = 0; j < PAIRS; ++j) {
    const double a = 1e-3 * (j + 1), b = a + 1e-6 * lane;
    if (T) tmem_st2(tb + 4 * j, a, b);
    if (L) smp[j * 128] = make_double2(a, b);
  }
  if (T) tmem_wait_st();
  __syncthreads();
  // verify the round trip once
  if (T) {
    bool good = true;
    for (int j = 0; j < PAIRS; ++j) {
      double a, b;
      tmem_ld2(tb + 4 * j, a, b);
      tmem_wait_ld();
      good &= (a == 1e-3 * (j + 1)) && (b == 1e-3 * (j + 1) + 1e-6 * lane);
    }
    if (!good) atomicExch(ok, 0);
  }
  const uint32_t smp_u32 = smem_u32(smp);
  double r0 = 1.0 + lane, r1 = 2.0, r2 = 3.0;
  double c0 = 0.5 + 1e-3 * lane, c1 = 0.25 + 2e-3 * lane, c2 = 0.125 + 3e-3 * lane, c3 = 0.0625 + 4e-3 * lane;
  const double m = 1.0000001, q = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int lev = 0; lev < LEVELS; ++lev) {
      double2 v[PAIRS / LEVELS];
      if (T) {
#pragma unroll
        for (int j = 0; j < PAIRS / LEVELS; ++j) tmem_ld2(tb + 4 * (lev * (PAIRS / LEVELS) + j), v[j].x, v[j].y);
        tmem_wait_ld();
      } else if (L) {
#pragma unroll
        for (int j = 0; j < PAIRS / LEVELS; ++j)   // volatile: the compiler must not hoist the loads out of the iteration loop
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[j].x), "=d"(v[j].y) : "r"(smp_u32 + (uint32_t)((lev * (PAIRS / LEVELS) + j) * 128 * 16)));
      }
      double lo0 = r0, lo1 = r1, lo2 = r2;
      if (S) {   // 6 doubles exchanged per level (12 SHFL.32) + 30 more per iteration below
        lo0 = __shfl_up_sync(0xffffffffu, r0, 1); lo1 = __shfl_up_sync(0xffffffffu, r1, 1); lo2 = __shfl_up_sync(0xffffffffu, r2, 1);
        lo0 += __shfl_down_sync(0xffffffffu, r0, 1); lo1 += __shfl_down_sync(0xffffffffu, r1, 1); lo2 += __shfl_down_sync(0xffffffffu, r2, 1);
      }
      if (T || L) {
#pragma unroll
        for (int j = 0; j + 2 < PAIRS / LEVELS; j += 3) {
          r0 = fma(v[j].x, lo0, r0); r1 = fma(v[j].y, lo1, r1); r2 = fma(v[j + 1].x, lo2, r2);
          r0 = fma(v[j + 1].y, lo1, r0); r1 = fma(v[j + 2].x, lo2, r1); r2 = fma(v[j + 2].y, lo0, r2);
        }
        r0 = fma(v[9].x, lo0, r0); r1 = fma(v[9].y, lo1, r1); r2 = fma(v[10].x, lo2, r2); r0 = fma(v[10].y, lo0, r0);
      }
      r0 *= 0.001; r1 *= 0.001; r2 *= 0.001;
    }
    if (S) {   // the other exchanges of an iteration: 15 doubles = 30 SHFL.32 (12 + 24 above = 78 in all)
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        c0 += __shfl_up_sync(0xffffffffu, c1, 1) * q;
        c1 += __shfl_down_sync(0xffffffffu, c2, 1) * q;
        c2 += __shfl_up_sync(0xffffffffu, c3, 1) * q;
      }
    }
    if (D) {   // 140 DFMAs in 4 chains
#pragma unroll
      for (int e = 0; e < 35; ++e) { c0 = fma(c0, m, q); c1 = fma(c1, m, q); c2 = fma(c2, m, q); c3 = fma(c3, m, q); }
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && w == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 128 + threadIdx.x] = r0 + r1 + r2 + c0 + c1 + c2 + c3;
  __syncthreads();
  if (w == 0) tmem_free(tbase_s, (uint32_t)cols);
}

template <bool T, bool L, bool D, bool S>
static void run(const char* name, int ctas_per_sm, int iters) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * ctas_per_sm;
  double* out; long long* cyc; int* ok;
  CK(cudaMalloc(&out, grid * 128 * sizeof(double)));
  CK(cudaMalloc(&cyc, grid * sizeof(long long)));
  CK(cudaMalloc(&ok, sizeof(int)));
  int one = 1;
  CK(cudaMemcpy(ok, &one, sizeof(int), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)PAIRS * 128 * 16;   // 88 KB: at most 2 CTAs per SM
  CK(cudaFuncSetAttribute(kern<T, L, D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int cols = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<T, L, D, S><<<grid, 128, smem>>>(10, cols, out, cyc, ok);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kern<T, L, D, S><<<grid, 128, smem>>>(iters, cols, out, cyc, ok);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  int hok = 0;
  CK(cudaMemcpy(&hok, ok, sizeof(int), cudaMemcpyDeviceToHost));
  double avg = 0; long long mx = 0;
  for (int i = 0; i < grid; ++i) { avg += (double)h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= grid;
  const double per_iter = avg / iters;   // cycles per iteration of ONE warp (all warps of the SM advance together)
  const double bytes = (double)ctas_per_sm * 4 * 32 * PAIRS * 16;   // multiplier bytes read per SM per iteration
  printf("%-8s warps/SM %d  cycles/iteration %8.1f (max CTA %8.1f)  multiplier B/cyc/SM %7.1f  SM-cycles per warp-iteration %7.1f  ms %.3f  roundtrip %s\n", name,
         4 * ctas_per_sm, per_iter, (double)mx / iters, (T || L) ? bytes / per_iter : 0.0, per_iter / (4 * ctas_per_sm), ms, hok ? "ok" : "BAD");
  free(h);
  CK(cudaFree(out)); CK(cudaFree(cyc)); CK(cudaFree(ok));
}

int main() {
  const int iters = 2000;
  for (int c = 1; c <= 2; ++c) {
    run<true, false, false, false>("T", c, iters);
    run<false, true, false, false>("L", c, iters);
    run<false, false, true, false>("D", c, iters);
    run<false, false, false, true>("S", c, iters);
    run<true, false, true, false>("T+D", c, iters);
    run<false, true, true, false>("L+D", c, iters);
    run<true, false, true, true>("T+D+S", c, iters);
    run<false, true, true, true>("L+D+S", c, iters);
  }
  return 0;
}
