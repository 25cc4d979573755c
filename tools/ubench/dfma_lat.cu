// DFMA dependent-issue latency and throughput on sm_100a (development micro-benchmark).
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, double a, double b) {
  double x[C];
#pragma unroll
  for (int i = 0; i < C; ++i) x[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < C; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int C>
void run(int warps_per_sm) {
  double* d; cudaMalloc(&d, 1 << 24);
  const int iters = 2000;
  int threads = warps_per_sm * 32;
  // one CTA per SM
  k<C><<<148, threads>>>(d, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<C><<<148, threads>>>(d, iters, 0.999999, 1e-7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  double n_inst_per_warp = (double)iters * 16 * C;
  printf("chains %d warps/SM %2d (per SMSP %.1f): cycles %.0f  cycles/DFMA/warp %.2f  DFMA warp-inst/clk/SMSP %.3f  TFLOP/s %.2f\n", C, warps_per_sm,
         warps_per_sm / 4.0, cyc, cyc / n_inst_per_warp, n_inst_per_warp * warps_per_sm / 4.0 / cyc,
         n_inst_per_warp * warps_per_sm * 32 * 2 * 148 / (ms * 1e-3) / 1e12);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
