// Micro-benchmark (round 2): can the quotient's rounding leave the FP64 pipe?
// The butterfly of modarith.cuh spends 2 of its 8 FP64 instructions on qh = rint(h / q) via the magic-constant trick
// (DFMA + DADD).  FRND.F64 (rint) does it in ONE instruction on another pipe — if it is fast enough and co-issues.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 frnd_bench.cu -o frnd_bench
#include <cstdio>
#include <cuda_runtime.h>
#define MAGIC 6755399441055744.0
template <int V>
__device__ __forceinline__ void bf(double &x, double &y, double w, double q, double qinv) {
  double h = __dmul_rn(y, w);
  double l = __fma_rn(y, w, -h);
  double qh;
  if (V == 0) qh = __fma_rn(h, qinv, MAGIC) - MAGIC;   // DFMA + DADD
  else qh = rint(__dmul_rn(h, qinv));                    // DMUL + FRND.F64
  double r = __fma_rn(-qh, q, h);
  double t = __dadd_rn(r, l);
  double X = x;
  x = __dadd_rn(X, t); y = __dsub_rn(X, t);
}
template <int V>
__global__ void __launch_bounds__(256) bench(double *out, double q, int iters) {
  double a[16];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = (double)((tid * 16 + i) % 1000003);
  double w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = (double)(123456789 + 977 * (tid + i));
  const double qinv = 1.0 / q;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 4; s++) {   // one 16-point network = 32 butterflies
      const int d = 8 >> s;
#pragma unroll
      for (int g = 0; g < (1 << s); g++)
#pragma unroll
        for (int o = 0; o < d; o++) bf<V>(a[g * 2 * d + o], a[g * 2 * d + o + d], w[g], q, qinv);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = a[k] * 0.0625;   // keep magnitudes bounded (same in both variants)
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[tid] = s;
}
__global__ void __launch_bounds__(256) probe_frnd(double *out, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1.37 + i * 0.77;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = rint(a[i]) + 0.3;   // FRND + DADD pairs
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const double q = 68719476731.0 - 1000.0;
  const int iters = 2000, grid = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char *name, auto launch, double per_thread_iter) {
    for (int w = 0; w < 2; w++) launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 5.0 * grid * 256 * iters * per_thread_iter;
    printf("%-32s %8.1f G/s\n", name, n / (ms * 1e-3) / 1e9);
  };
  run("butterfly, magic rounding (8 DP)", [&] { bench<0><<<grid, 256>>>(out, q, iters); }, 32);
  run("butterfly, FRND.F64 (7 DP + FRND)", [&] { bench<1><<<grid, 256>>>(out, q, iters); }, 32);
  run("FRND + DADD pairs", [&] { probe_frnd<<<grid, 256>>>(out, iters); }, 96);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
