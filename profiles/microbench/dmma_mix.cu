// How should non-DMMA work be arranged around DMMA m8n8k4 so that the FP64 tensor path keeps its rate?
//   A. every warp: G DMMAs (6 independent chains) then E integer instructions (grouped, like k_bconv_mma's epilogue)
//   B. warp specialisation: even warps issue only DMMAs, odd warps only integer instructions
//   C. every warp: G DMMAs then E/2 DFMA (the real epilogue is FP64 too)
//   D. every warp: G DMMAs then L shared-memory loads
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE, int G, int E>
__global__ void k(double *out, int iters, unsigned m) {
  __shared__ double sm[1024];
  double d[6][2], f[8];
  unsigned x[8];
  for (int i = 0; i < 6; i++) { d[i][0] = threadIdx.x; d[i][1] = i; }
  for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; x[i] = threadIdx.x * 7 + i; }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3, c = 1.0000001, e = 1e-3;
  const bool mma_warp = MODE != 1 || ((threadIdx.x >> 5) & 1) == 0;
  const bool other_warp = MODE != 1 || ((threadIdx.x >> 5) & 1) == 1;
  for (int it = 0; it < iters; it++) {
    if (mma_warp) {
#pragma unroll
      for (int g = 0; g < G; g++) dmma884(d[g % 6][0], d[g % 6][1], a, b);
    }
    if (other_warp) {
      if (MODE == 0 || MODE == 1) {
#pragma unroll
        for (int k2 = 0; k2 < E; k2++) x[k2 & 7] = (x[k2 & 7] + m) ^ x[(k2 + 3) & 7];
      } else if (MODE == 2) {
#pragma unroll
        for (int k2 = 0; k2 < E / 2; k2++) f[k2 & 7] = __fma_rn(f[k2 & 7], c, e);
      } else {
#pragma unroll
        for (int k2 = 0; k2 < E; k2++) f[k2 & 7] += sm[(threadIdx.x + 32 * k2 + it) & 1023];
      }
    }
  }
  double s = 0; unsigned y = 0;
  for (int i = 0; i < 6; i++) s += d[i][0] + d[i][1];
  for (int i = 0; i < 8; i++) { s += f[i]; y ^= x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + y;
}
template <int MODE, int G, int E> void run(const char *name, double *out) {
  const int blocks = 148 * 4, thr = 256, iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, G, E><<<blocks, thr>>>(out, 10, 3); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE, G, E><<<blocks, thr>>>(out, iters, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks * thr / 32 * (MODE == 1 ? 0.5 : 1.0);
  printf("%-44s G=%2d E=%3d: %.3f ms  DMMA %.2f TFMA/s\n", name, G, E, ms, warps * iters * G * 256 / ms / 1e9);
}
int main() {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  run<0, 24, 0>("DMMA only (6 chains)", out);
  run<0, 24, 24>("same warp: DMMAs then ints", out);
  run<0, 24, 96>("same warp: DMMAs then ints", out);
  run<1, 24, 24>("specialised warps (half DMMA, half int)", out);
  run<1, 24, 96>("specialised warps (half DMMA, half int)", out);
  run<2, 24, 48>("same warp: DMMAs then 24 DFMA", out);
  run<2, 24, 96>("same warp: DMMAs then 48 DFMA", out);
  run<3, 24, 8>("same warp: DMMAs then 8 LDS", out);
  run<3, 24, 32>("same warp: DMMAs then 32 LDS", out);
  return 0;
}
