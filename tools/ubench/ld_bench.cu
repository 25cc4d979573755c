// Dense limb-darkening throughput of the solution-vector variants (development micro-benchmark).
#include <cstdio>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>
#include "../../jaxoplanet_b200/csrc/jxp_math.cuh"
using namespace jxp;

template <int PU, int MINB>
__global__ void __launch_bounds__(128, MINB) k_scalar(const double* b, const double* r, double* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double s0, s1, s2;
    ld_solution<double, 11, false, PU>(b[i], r[i], 2, nullptr, s0, s1, s2);
    out[i] = ld_combine(2, s0, s1, s2, 0.3, 0.2, 0.1);
  }
}
template <int N, int PU, int MINB>
__global__ void __launch_bounds__(128, MINB) k_multi(const double* b, const double* r, double* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + (N - 1) * stride < n; i += N * stride) {
    double bb[N], rr[N], s0[N], s1[N], s2[N];
#pragma unroll
    for (int q = 0; q < N; ++q) { bb[q] = b[i + q * stride]; rr[q] = r[i + q * stride]; }
    ld_solution_n<N, PU>(bb, rr, 2, s0, s1, s2);
#pragma unroll
    for (int q = 0; q < N; ++q) out[i + q * stride] = ld_combine(2, s0[q], s1[q], s2[q], 0.3, 0.2, 0.1);
  }
}
// planet fully inside the disk: the constant-node code, N points per thread
template <int N, int MINB>
__global__ void __launch_bounds__(128, MINB) k_inside(const double* b, const double* r, double* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + (N - 1) * stride < n; i += N * stride) {
    double bb[N], rr[N], s0[N], s1[N], s2[N];
#pragma unroll
    for (int q = 0; q < N; ++q) { bb[q] = b[i + q * stride]; rr[q] = r[i + q * stride]; }
    ld_solution_inside_n<N>(bb, rr, 2, s0, s1, s2);
#pragma unroll
    for (int q = 0; q < N; ++q) out[i + q * stride] = ld_combine(2, s0[q], s1[q], s2[q], 0.3, 0.2, 0.1);
  }
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int k = 0; k < 3; ++k) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  return best;
}
int main() {
  const int64_t n = 148LL * 128 * 16 * 64;  // 19.4 M points
  std::vector<double> hb(n), hr(n);
  std::mt19937_64 g(1); std::uniform_real_distribution<double> U(0, 1);
  for (int64_t i = 0; i < n; ++i) { hr[i] = 0.03 + 0.12 * U(g); hb[i] = U(g) * (1 + hr[i]); }
  double *b, *r, *o0, *o1; cudaMalloc(&b, n * 8); cudaMalloc(&r, n * 8); cudaMalloc(&o0, n * 8); cudaMalloc(&o1, n * 8);
  cudaMemcpy(b, hb.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(r, hr.data(), n * 8, cudaMemcpyHostToDevice);
  std::vector<double> h0(n), h1(n);
  auto report = [&](const char* name, float ms, double* o) {
    cudaMemcpy(h1.data(), o, n * 8, cudaMemcpyDeviceToHost);
    double md = 0; for (int64_t i = 0; i < n; ++i) md = fmax(md, fabs(h1[i] - h0[i]));
    printf("%-22s %.3f ms  %.3e evals/s  A-table %.1f%% of 36.6 TF  max|diff vs scalar| %.2e\n", name, ms, n / (ms * 1e-3), n * 746.0 / (ms * 1e-3) / 36.6e12 * 100, md);
  };
#define RUN_S(PU, MINB) { float ms = timeit([&] { k_scalar<PU, MINB><<<148 * MINB, 128>>>(b, r, o0, n); }); if (PU == 1 && MINB == 8) cudaMemcpy(h0.data(), o0, n * 8, cudaMemcpyDeviceToHost); report("scalar PU" #PU " MINB" #MINB, ms, o0); }
#define RUN_M(N, PU, MINB) { float ms = timeit([&] { k_multi<N, PU, MINB><<<148 * MINB, 128>>>(b, r, o1, n); }); report("multi N" #N " PU" #PU " MINB" #MINB, ms, o1); }
  RUN_S(1, 8) RUN_S(1, 5) RUN_S(5, 5) RUN_S(5, 4)
  RUN_M(1, 5, 5) RUN_M(2, 5, 4)
  // inside-the-disk points only (b < 1 - r)
  for (int64_t i = 0; i < n; ++i) hb[i] = U(g) * (1 - hr[i]) * 0.999;
  cudaMemcpy(b, hb.data(), n * 8, cudaMemcpyHostToDevice);
  { float ms = timeit([&] { k_scalar<5, 5><<<148 * 5, 128>>>(b, r, o0, n); }); cudaMemcpy(h0.data(), o0, n * 8, cudaMemcpyDeviceToHost); report("inside pts: general", ms, o0); }
#define RUN_I(N, MINB) { float ms = timeit([&] { k_inside<N, MINB><<<148 * MINB, 128>>>(b, r, o1, n); }); report("inside N" #N " MINB" #MINB, ms, o1); }
  RUN_I(1, 8) RUN_I(1, 5) RUN_I(2, 5) RUN_I(2, 4) RUN_I(3, 4) RUN_I(4, 3) RUN_I(4, 2)
  return 0;
}
