// FP64 tensor-core (DMMA) throughput on B200 and its interaction with the FP64 vector pipe / issue port.
//   1. mma.sync m8n8k4 f64, 8 independent accumulator tiles per warp
//   2. mma.sync m16n8k16 f64 (sm_90+ shape)
//   3. DMMA interleaved with independent DFMA chains (do they share a pipe?)
//   4. DMMA interleaved with integer ops (does a DMMA hold the issue port like a DFMA does?)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NF, int NI>
__global__ void k884(double *out, int iters, unsigned m) {
  double d[8][2], f[8];
  unsigned x[8];
  for (int i = 0; i < 8; i++) { d[i][0] = threadIdx.x; d[i][1] = i; f[i] = threadIdx.x + i; x[i] = threadIdx.x * 7 + i; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3, c = 1.0000001, e = 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        dmma884(d[i][0], d[i][1], a, b);
#pragma unroll
        for (int k = 0; k < NF; k++) f[(i + k) & 7] = __fma_rn(f[(i + k) & 7], c, e);
#pragma unroll
        for (int k = 0; k < NI; k++) x[(i + k) & 7] = (x[(i + k) & 7] + m) ^ x[(i + k + 3) & 7];
      }
  }
  double s = 0; unsigned y = 0;
  for (int i = 0; i < 8; i++) { s += d[i][0] + d[i][1] + f[i]; y ^= x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + y;
}

__global__ void k16816(double *out, int iters) {
  double d[4][4], a[8], b[4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[i][j] = threadIdx.x + i + j;
  for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; i++) b[i] = 1e-3 * (i + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) dmma16816(d[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(10); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(2000); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const int blocks = 148 * 8, thr = 256, iters = 2000;
  const double warps = (double)blocks * thr / 32;
  for (int wpb : {256, 128, 64, 32}) {
    float ms = timeit([&](int it) { k884<0, 0><<<148 * 8, wpb>>>(out, it, 3); });
    double fma = (double)148 * 8 * wpb / 32 * iters * 32 * 256;
    printf("m8n8k4 alone, %d thr/block x8 blocks/SM: %.3f ms  %.2f TFMA/s (DFMA peak 18.35)  %.3f DMMA/clk/SMSP\n", wpb, ms, fma / ms / 1e9,
           (double)148 * 8 * wpb / 32 * iters * 32 / (ms * 1e-3) / (148 * 4) / 1.965e9);
  }
  {
    float ms = timeit([&](int it) { k16816<<<blocks, thr>>>(out, it); });
    double fma = warps * iters * 16 * 2048;
    printf("m16n8k16 alone: %.3f ms  %.2f TFMA/s\n", ms, fma / ms / 1e9);
  }
  {
    float ms = timeit([&](int it) { k884<1, 0><<<blocks, thr>>>(out, it, 3); });
    printf("m8n8k4 + 1 DFMA each: %.3f ms  DMMA %.2f TFMA/s + DFMA %.2f Tlane/s\n", ms, warps * iters * 32 * 256 / ms / 1e9, warps * iters * 32 * 32 / ms / 1e9);
    ms = timeit([&](int it) { k884<4, 0><<<blocks, thr>>>(out, it, 3); });
    printf("m8n8k4 + 4 DFMA each: %.3f ms  DMMA %.2f TFMA/s + DFMA %.2f Tlane/s\n", ms, warps * iters * 32 * 256 / ms / 1e9, warps * iters * 32 * 128 / ms / 1e9);
    ms = timeit([&](int it) { k884<8, 0><<<blocks, thr>>>(out, it, 3); });
    printf("m8n8k4 + 8 DFMA each: %.3f ms  DMMA %.2f TFMA/s + DFMA %.2f Tlane/s\n", ms, warps * iters * 32 * 256 / ms / 1e9, warps * iters * 32 * 256 / ms / 1e9);
    ms = timeit([&](int it) { k884<0, 2><<<blocks, thr>>>(out, it, 3); });
    printf("m8n8k4 + 4 int instr each: %.3f ms  DMMA %.2f TFMA/s\n", ms, warps * iters * 32 * 256 / ms / 1e9);
    ms = timeit([&](int it) { k884<0, 6><<<blocks, thr>>>(out, it, 3); });
    printf("m8n8k4 + 12 int instr each: %.3f ms  DMMA %.2f TFMA/s\n", ms, warps * iters * 32 * 256 / ms / 1e9);
  }
  return 0;
}
