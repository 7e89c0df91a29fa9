// tools/fp64_peak.cu -- measures the DFMA issue ceiling of the device (the co-limiter of the push+deposit
// kernel next to HBM; MEASURED_PEAKS.json only carries HBM and bf16 figures).  Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, iters = 4096;
  double* out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, 256>>>(out, 64, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<<<blocks, 256>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double fma = (double)blocks * 256 * iters * 64;
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm_at_1965MHz\": %.1f, \"ms\": %.3f}\n",
         p.name, p.multiProcessorCount, fma / (best * 1e-3), 2 * fma / (best * 1e-3) / 1e12,
         fma / (best * 1e-3) / p.multiProcessorCount / 1.965e9, best);
  return 0;
}
