// Measurement utility (not on the solve path): FP64 FMA issue-rate micro-benchmark used as the roofline
// denominator for the ADMM kernel ("FP64 FMA issue peak micro-benchmarked on the B200 box").
#include <cuda_runtime.h>
#include "../../include/f110_mpc_b200.h"

namespace {
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123456.789) out[0] = s;  // keep the chain alive
}
}  // namespace

extern "C" int f110_bench_fp64_fma(int device, int iters, double* tflops_out) {
  if (!tflops_out || iters <= 0) return F110_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return F110_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return F110_ERR_CUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return F110_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (e != cudaSuccess) return F110_ERR_CUDA;
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return F110_OK;
}
