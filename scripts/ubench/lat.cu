// Micro-benchmarks (developer tool): dependent-issue latency of DFMA / DADD / SHFL(64-bit) / LDS.128 on one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, long long* cyc, int n, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_ilp(double* out, long long* cyc, int n, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int n) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_fma(double* out, long long* cyc, int n, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = fma(__shfl_up_sync(0xffffffffu, x, 1), a, b); x = fma(__shfl_up_sync(0xffffffffu, x, 1), a, b); x = fma(__shfl_up_sync(0xffffffffu, x, 1), a, b); x = fma(__shfl_up_sync(0xffffffffu, x, 1), a, b); }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ double2 sm[64];
  sm[threadIdx.x] = make_double2((double)((threadIdx.x + 1) & 31), 0.0); sm[threadIdx.x + 32] = sm[threadIdx.x];
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { idx = (int)sm[idx].x; idx = (int)sm[idx].x; idx = (int)sm[idx].x; idx = (int)sm[idx].x; }
  long long t1 = clock64();
  out[threadIdx.x] = idx; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dsetp(double* out, long long* cyc, int n, double a) {
  double x = threadIdx.x + 0.5;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x > a ? x * 0.999 : a; x = x > a ? x * 0.999 : a; x = x > a ? x * 0.999 : a; x = x > a ? x * 0.999 : a; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64);
  const int n = 4096; long long h;
  auto rep = [&](const char* name, int ops) { cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s %.2f cycles/op\n", name, (double)h / (4.0 * n) ); (void)ops; };
  for (int w = 0; w < 2; ++w) {
    k_dfma<<<1, 32>>>(d, c, n, 0.999999, 1e-9); cudaDeviceSynchronize(); if (w) rep("DFMA dependent", 4);
    k_dfma_ilp<<<1, 32>>>(d, c, n, 0.999999, 1e-9); cudaDeviceSynchronize(); if (w) rep("DFMA 4 independent (per op)", 4);
    k_shfl<<<1, 32>>>(d, c, n); cudaDeviceSynchronize(); if (w) rep("SHFL.64 dependent", 4);
    k_shfl_fma<<<1, 32>>>(d, c, n, 0.999999, 1e-9); cudaDeviceSynchronize(); if (w) rep("SHFL.64 + DFMA dependent", 4);
    k_lds<<<1, 32>>>(d, c, n); cudaDeviceSynchronize(); if (w) rep("LDS.128 + F2I dependent", 4);
    k_dsetp<<<1, 32>>>(d, c, n, 0.25); cudaDeviceSynchronize(); if (w) rep("DSETP+FSEL+DMUL dependent", 4);
  }
  // 8 warps on one SM partition mix: throughput of dependent DFMA with 2 warps per SMSP
  k_dfma<<<1, 256>>>(d, c, n, 0.999999, 1e-9); cudaDeviceSynchronize(); rep("DFMA dependent, 8 warps/SM", 4);
  k_dfma_ilp<<<1, 256>>>(d, c, n, 0.999999, 1e-9); cudaDeviceSynchronize(); rep("DFMA 4-ILP, 8 warps/SM", 4);
  return 0;
}
