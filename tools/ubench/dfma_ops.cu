// DFMA issue rate vs where the operands come from (register file / constant bank) -- dev micro-benchmark.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[4] = {0.999999, 1e-7, 1.000001, -1e-7};

// MODE 0: x = fma(x, a, b) with a, b registers shared by all chains (reuse-able)
// MODE 1: x_i = fma(x_i, y_i, w_i) with per-chain registers (3 distinct RF operands, no reuse)
// MODE 2: x_i = fma(x_i, cA, cB) with kernel-parameter constants (c[0][..] operands)
// MODE 3: x_i = fma(x_i, y_i, cB): two registers + one constant
template <int C, int MODE>
__global__ void k(double* out, const double* in, int iters, double pa, double pb) {
  double x[C], y[C], w[C];
#pragma unroll
  for (int i = 0; i < C; ++i) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[1024 + threadIdx.x + 32 * i]; w[i] = in[2048 + threadIdx.x + 32 * i]; }
  double a = in[4000], b = in[4001];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < C; ++i) {
        if (MODE == 0) x[i] = fma(x[i], a, b);
        if (MODE == 1) x[i] = fma(x[i], y[i], w[i]);
        if (MODE == 2) x[i] = fma(x[i], pa, pb);
        if (MODE == 3) x[i] = fma(x[i], y[i], pb);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += x[i] + y[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int C, int MODE>
void run(int warps_per_sm, double* d, double* in) {
  const int iters = 1000;
  k<C, MODE><<<148, warps_per_sm * 32>>>(d, in, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  k<C, MODE><<<148, warps_per_sm * 32>>>(d, in, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  double n = (double)iters * 16 * C;
  printf("mode %d chains %d warps/SMSP %.0f: DFMA warp-inst/clk/SMSP %.3f (%.0f%% of 0.5)\n", MODE, C, warps_per_sm / 4.0,
         n * warps_per_sm / 4.0 / cyc, n * warps_per_sm / 4.0 / cyc / 0.5 * 100);
}
int main() {
  double *d, *in; cudaMalloc(&d, 1 << 24); cudaMalloc(&in, 1 << 16);
  double h[8192]; for (int i = 0; i < 8192; ++i) h[i] = 0.9999 + 1e-6 * (i % 7);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {16, 32}) {
    run<1, 0>(w, d, in); run<2, 0>(w, d, in); run<4, 0>(w, d, in);
    run<1, 1>(w, d, in); run<2, 1>(w, d, in); run<4, 1>(w, d, in);
    run<1, 2>(w, d, in); run<2, 2>(w, d, in); run<4, 2>(w, d, in);
    run<1, 3>(w, d, in); run<2, 3>(w, d, in); run<4, 3>(w, d, in);
  }
  return 0;
}
