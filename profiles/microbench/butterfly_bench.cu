// Micro-benchmark: how many radix-2 NTT butterflies per second can B200's integer pipes sustain for
// 36-bit primes held in 64-bit words?  Decides the modmul flavour used by homulator_b200/csrc.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo butterfly_bench.cu -o butterfly_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

// ---- V1: exact Shoup, Harvey lazy ranges [0,4q) with conditional subtraction
__device__ __forceinline__ void bf_v1(u64 &x, u64 &y, u64 w, u64 wp, u64 q, u64 q2) {
  u64 X = x - (x >= q2 ? q2 : 0);
  u64 h = __umul64hi(y, wp);
  u64 T = y * w - h * q;              // [0,2q)
  x = X + T; y = X - T + q2;
}
// ---- V2: exact Shoup, no range correction (values grow by 2q per stage; 36-bit q has 28 bits of headroom)
__device__ __forceinline__ void bf_v2(u64 &x, u64 &y, u64 w, u64 wp, u64 q, u64 q2) {
  u64 h = __umul64hi(y, wp);
  u64 T = y * w - h * q;              // [0,2q)
  u64 X = x;
  x = X + T; y = X + q2 - T;
}
// ---- V3: approximate quotient from 3 partial products, no range correction (T in [0,5q))
__device__ __forceinline__ void bf_v3(u64 &x, u64 &y, u64 w, u64 wp, u64 q, u64 q5) {
  u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
  u64 h = (u64)y1 * p1 + __umulhi(y0, p1) + __umulhi(y1, p0);
  u64 T = y * w - h * q;
  u64 X = x;
  x = X + T; y = X + q5 - T;
}
// ---- V4: 32-bit word-serial Montgomery (two rounds), twiddle in Montgomery form, no companion
__device__ __forceinline__ void bf_v4(u64 &x, u64 &y, u64 w, u32 qinv_neg, u64 q, u64 q2) {
  u32 q0 = (u32)q, q1 = (u32)(q >> 32);
  u32 y0 = (u32)y, y1 = (u32)(y >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
  u64 p01 = (u64)y0 * w0;
  u64 p12 = (u64)y0 * w1 + (u64)y1 * w0 + (p01 >> 32);
  p12 += ((u64)(y1 * w1)) << 32;
  u32 pl = (u32)p01;
  // round 1
  u32 m0 = pl * qinv_neg;
  u64 t = (u64)m0 * q0 + pl;           // low word becomes 0
  u64 acc = p12 + (t >> 32) + (u64)m0 * q1;
  // round 2
  u32 a0 = (u32)acc;
  u32 m1 = a0 * qinv_neg;
  u64 t2 = (u64)m1 * q0 + a0;
  u64 T = (acc >> 32) + (t2 >> 32) + (u64)m1 * q1;
  u64 X = x;
  x = X + T; y = X + q2 - T;
}

// ---- V5: all-FP64 signed-lazy butterfly: values are doubles holding integers |v| < 2^52
__device__ __forceinline__ void bf_v5(double &x, double &y, double w, double wq, double q) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: fma(...)+MAGIC-MAGIC = rint
  double qh = __fma_rn(y, wq, MAGIC) - MAGIC;
  double h = y * w;
  double l = __fma_rn(y, w, -h);
  double r = __fma_rn(-qh, q, h);
  double T = r + l;
  double X = x;
  x = X + T; y = X - T;
}
__global__ void __launch_bounds__(256) bench_dp(double *out, const u64 *tw, u64 q, int iters) {
  double a[8];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = (double)((u64)(tid * 8 + i) % q);
  double w[7], wq[7];
  const double qd = (double)q;
#pragma unroll
  for (int i = 0; i < 7; i++) { w[i] = (double)tw[(tid + i) & 1023]; wq[i] = w[i] / qd; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const int h = 4 >> s;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if ((k & h) == 0) {
          const int ti = (1 << s) - 1 + (k >> (3 - s));
          bf_v5(a[k], a[k + h], w[ti], wq[ti], qd);
        }
      }
    }
    // bound the growth like the integer variants do (one cheap op per element per 3 stages)
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = a[k] * 0.125;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  out[tid] = s;
}
__global__ void __launch_bounds__(256) probe_dfma(u64 *out, u32 m, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  const double c = 1.0 + 1e-9 * m, d = 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __fma_rn(a[i], c, d);
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u64)s;
}
__global__ void __launch_bounds__(256) probe_imadhi(u64 *out, u32 m, int iters) {
  u32 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __umulhi(a[i], m) + a[(i + 1) & 7];   // IMAD.HI.U32
  }
  u32 s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: half the warps of each CTA do integer V2 butterflies, half do FP64 V5 (do the pipes overlap?)

template <int V>
__global__ void __launch_bounds__(256) bench(u64 *out, const u64 *tw, u64 q, int iters) {
  u64 a[8];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = (u64)(tid * 8 + i) % q;
  u64 w[7], wp[7];
#pragma unroll
  for (int i = 0; i < 7; i++) { w[i] = tw[(tid + i) & 1023]; wp[i] = tw[1024 + ((tid + i) & 1023)]; }
  const u64 q2 = 2 * q, q5 = 5 * q;
  const u32 qn = (u32)tw[2048];
  for (int it = 0; it < iters; it++) {
    // one radix-8 round = 12 butterflies
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const int h = 4 >> s;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if ((k & h) == 0) {
          const int ti = (1 << s) - 1 + (k >> (3 - s));
          if (V == 1) bf_v1(a[k], a[k + h], w[ti], wp[ti], q, q2);
          if (V == 2) bf_v2(a[k], a[k + h], w[ti], wp[ti], q, q2);
          if (V == 3) bf_v3(a[k], a[k + h], w[ti], wp[ti], q, q5);
          if (V == 4) bf_v4(a[k], a[k + h], w[ti], qn, q, q2);
        }
      }
    }
    // keep ranges bounded for the no-correction variants: cheap mask (not part of a real NTT; ~8 LOP3 per 12 bf)
    if (V != 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] &= 0x3FFFFFFFFFull;
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
  out[tid] = s;
}

// raw pipe probes: 8 independent chains per thread
__global__ void __launch_bounds__(256) probe_imad(u64 *out, u32 m, int iters) {
  u64 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = (u64)((u32)a[i]) * m + a[i];   // IMAD.WIDE.U32
  }
  u64 s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) probe_imad32(u64 *out, u32 m, int iters) {
  u32 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] * m + a[(i + 1) & 7];   // IMAD (32-bit)
  }
  u32 s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) probe_iadd(u64 *out, u32 m, int iters) {
  u32 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = (a[i] + m) ^ a[(i + 3) & 7];   // IADD3 + LOP3
  }
  u32 s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K>
static void probe(const char *name, K k, u64 *out, double ops_per_iter) {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, 12345u, 10); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); k<<<blocks, threads>>>(out, 12345u, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * ops_per_iter;
  printf("%-28s %8.3f ms  %8.2f Tlane-op/s  = %6.1f lane-ops/clk/SM at 1.9 GHz\n", name, best, ops / best / 1e9, ops / best / 1e9 * 1e3 / 148 / 1.9);
}

template <int V>
static void run(const char *name, u64 *out, u64 *tw, u64 q) {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<V><<<blocks, threads>>>(out, tw, q, 10);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    bench<V><<<blocks, threads>>>(out, tw, q, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double bf = (double)blocks * threads * iters * 12.0;
  printf("%-28s %8.3f ms  %8.2f Gbutterfly/s  -> NTT(2^16) ceiling %6.2f Mlimb/s (%5.1f%% of 6.24 M HBM roofline)\n", name,
         best, bf / best / 1e6, bf / best * 1e3 / 524288.0 / 1e6, bf / best * 1e3 / 524288.0 / 1e6 / 6.24 * 100);
}

int main() {
  u64 *out, *tw;
  cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaMalloc(&tw, 2049 * 8);
  const u64 q = 68718428161ull;  // a 36-bit prime = 1 mod 2^17
  u64 h[2049];
  for (int i = 0; i < 1024; i++) { h[i] = (0x9E3779B97F4A7C15ull * (i + 1)) % q; h[1024 + i] = (u64)(((unsigned __int128)h[i] << 64) / q); }
  h[2048] = 0xdeadbeefu | 1;
  cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
  probe("IMAD.WIDE.U32", probe_imad, out, 96.0);
  probe("IMAD 32", probe_imad32, out, 96.0);
  probe("IADD3+LOP3 (2 ops)", probe_iadd, out, 192.0);
  probe("IMAD.HI.U32", probe_imadhi, out, 96.0);
  probe("DFMA", probe_dfma, out, 96.0);
  {
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_dp<<<blocks, threads>>>((double *)out, tw, q, 10); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0); bench_dp<<<blocks, threads>>>((double *)out, tw, q, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double bf = (double)blocks * threads * iters * 12.0;
    printf("%-28s %8.3f ms  %8.2f Gbutterfly/s -> %6.3f Mlimb/s\n", "V5 all-FP64 signed lazy", best, bf / best / 1e6, bf / best / 1e6 / 524288.0 * 1e3);
  }
  run<1>("V1 shoup exact + harvey", out, tw, q);
  run<2>("V2 shoup exact, wide-lazy", out, tw, q);
  run<3>("V3 shoup approx, wide-lazy", out, tw, q);
  run<4>("V4 montgomery 2x32", out, tw, q);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
