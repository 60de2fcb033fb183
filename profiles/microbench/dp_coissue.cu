// Does a non-FP64 instruction issued between two DFMAs cost extra time, or does it hide in the FP64 pipe's 2-cycle
// issue interval?  8 independent DFMA chains per thread + K independent integer (IADD3/LOP3) ops per DFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(double *out, int iters, unsigned m) {
  double a[8];
  unsigned x[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; x[i] = threadIdx.x * 7 + i; }
  const double c = 1.0000001, d = 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        a[i] = __fma_rn(a[i], c, d);
#pragma unroll
        for (int kk = 0; kk < K; kk++) x[(i + kk) & 7] = (x[(i + kk) & 7] + m) ^ x[(i + kk + 3) & 7];
      }
  }
  double s = 0; unsigned y = 0;
  for (int i = 0; i < 8; i++) { s += a[i]; y ^= x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + y;
}
template <int K> void run() {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000, blocks = 148 * 8;
  k<K><<<blocks, 256>>>(out, 10, 3); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<K><<<blocks, 256>>>(out, iters, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double dfma = (double)blocks * 256 * iters * 64;
  printf("K=%d int-ops(x2 instr) per DFMA: %.3f ms  DFMA rate %.2f Tlane/s (peak 18.35)  total instr/clk/SMSP %.2f\n", K, ms, dfma / ms / 1e9,
         dfma * (1 + 2 * K) / 32 / (ms * 1e-3) / (148 * 4) / 1.965e9);
  cudaFree(out);
}
int main() { run<0>(); run<1>(); run<2>(); run<3>(); run<4>(); return 0; }
