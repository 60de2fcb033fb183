// See ewe.cuh.  All kernels are streaming (HBM-bound) except base conversion (FP64-pipe bound).
#include "ewe.cuh"

namespace hml {

constexpr int EW_THREADS = 256;  // 2 coefficients per thread, 16-byte accesses

__device__ __forceinline__ ulonglong2 ld2(const u64 *p, size_t i2) { return __ldg(reinterpret_cast<const ulonglong2 *>(p) + i2); }
__device__ __forceinline__ void st2(u64 *p, size_t i2, u64 a, u64 b) { reinterpret_cast<ulonglong2 *>(p)[i2] = make_ulonglong2(a, b); }
__device__ __forceinline__ u64 finish(double v, const ModConst &m) { return f64_to_canonical(reduce_signed(v, m.q, m.qinv), m.qi); }

static inline dim3 ew_grid(int N, int ny, int nz = 1) { return dim3((N / 2 + EW_THREADS - 1) / EW_THREADS, ny, nz); }

// ------------------------------------------------------------------------------------------------ generic EWE
__global__ void __launch_bounds__(EW_THREADS) k_ewe(const ModConst *__restrict__ mc, LimbMap lm, int N, const u64 *x1,
                                                    const u64 *x2, const u64 *x3, const u64 *x4, int subtract, u64 *out) {
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= N / 2) return;
  const int limb = blockIdx.y;
  const ModConst m = mc[lm.mod[limb]];
  const size_t o = (size_t)limb * (N / 2) + i2;
  double p0 = 0, p1 = 0, s0 = 0, s1 = 0;
  if (x1) {
    ulonglong2 a = ld2(x1, o);
    p0 = u64_to_f64(a.x); p1 = u64_to_f64(a.y);
    if (x2) {
      ulonglong2 b = ld2(x2, o);
      p0 = mulmod_var(p0, u64_to_f64(b.x), m.q, m.qinv);
      p1 = mulmod_var(p1, u64_to_f64(b.y), m.q, m.qinv);
    }
  }
  if (x3) {
    ulonglong2 a = ld2(x3, o);
    s0 = u64_to_f64(a.x); s1 = u64_to_f64(a.y);
    if (x4) {
      ulonglong2 b = ld2(x4, o);
      s0 = mulmod_var(s0, u64_to_f64(b.x), m.q, m.qinv);
      s1 = mulmod_var(s1, u64_to_f64(b.y), m.q, m.qinv);
    }
  }
  if (subtract) { s0 = -s0; s1 = -s1; }
  st2(out, o, finish(p0 + s0, m), finish(p1 + s1, m));
}

void launch_ewe(const ModConst *mc, const LimbMap &lm, int N, int n_limbs, const u64 *x1, const u64 *x2, const u64 *x3,
                const u64 *x4, int subtract, u64 *out, cudaStream_t s) {
  k_ewe<<<ew_grid(N, n_limbs), EW_THREADS, 0, s>>>(mc, lm, N, x1, x2, x3, x4, subtract, out);
}

// ------------------------------------------------------------------------------------------------ tensor product
__global__ void __launch_bounds__(EW_THREADS) k_tensor3(const ModConst *__restrict__ mc, int N, const u64 *a0, const u64 *a1,
                                                        const u64 *b0, const u64 *b1, u64 *d0, u64 *d1, u64 *d2,
                                                        long long in_stride, long long out_stride) {
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= N / 2) return;
  const int limb = blockIdx.y;  // limbs 0..L-1 are moduli q_0..q_{L-1}
  const ModConst m = mc[limb];
  const size_t o = (size_t)limb * (N / 2) + i2;
  const long long bi = (long long)blockIdx.z * in_stride, bo = (long long)blockIdx.z * out_stride;  // ciphertext of the batch
  a0 += bi; a1 += bi; b0 += bi; b1 += bi; d0 += bo; d1 += bo; d2 += bo;
  const ulonglong2 A0 = ld2(a0, o), A1 = ld2(a1, o), B0 = ld2(b0, o), B1 = ld2(b1, o);
  u64 r0[2], r1[2], r2[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double x0 = u64_to_f64(k ? A0.y : A0.x), x1 = u64_to_f64(k ? A1.y : A1.x);
    const double y0 = u64_to_f64(k ? B0.y : B0.x), y1 = u64_to_f64(k ? B1.y : B1.x);
    r0[k] = f64_to_canonical(mulmod_var(x0, y0, m.q, m.qinv), m.qi);
    r1[k] = finish(mulmod_var(x0, y1, m.q, m.qinv) + mulmod_var(x1, y0, m.q, m.qinv), m);
    r2[k] = f64_to_canonical(mulmod_var(x1, y1, m.q, m.qinv), m.qi);
  }
  st2(d0, o, r0[0], r0[1]);
  st2(d1, o, r1[0], r1[1]);
  st2(d2, o, r2[0], r2[1]);
}

void launch_tensor3(const ModConst *mc, int N, int L, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *d0,
                    u64 *d1, u64 *d2, int n_batch, long long in_stride, long long out_stride, cudaStream_t s) {
  k_tensor3<<<ew_grid(N, L, n_batch), EW_THREADS, 0, s>>>(mc, N, a0, a1, b0, b1, d0, d1, d2, in_stride, out_stride);
}

// ------------------------------------------------------------------------------------------------ key-switch inner product
// The key words of a (limb, coefficient pair) are loaded once and reused for every ciphertext of the batch
// (the key is 60% of the traffic of an unbatched inner product).
template <int IP_MAX_BETA>
__global__ void __launch_bounds__(EW_THREADS) k_inner(const ModConst *__restrict__ mc, LimbMap lm, InnerArgs a) {
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= a.N / 2) return;
  const int e = blockIdx.y;
  const ModConst m = mc[lm.mod[e]];
  const int kl = lm.pos[e], own = lm.skip[e];
  const size_t n2 = a.N / 2;
  // all loads of the first ciphertext (digits + key words) are issued together; afterwards the next ciphertext's
  // digits are in flight while the current one is multiplied
  auto digit_ptr = [&](int b, int j) {
    return j == own ? a.d + (size_t)b * a.d_batch_stride + (size_t)e * a.N
                    : a.ext + (size_t)b * a.ext_batch_stride + ((size_t)j * a.n_ext + e) * a.N;
  };
  ulonglong2 tn[IP_MAX_BETA];
#pragma unroll
  for (int j = 0; j < IP_MAX_BETA; ++j)
    if (j < a.beta) tn[j] = ld2(digit_ptr(0, j), i2);
  double k[IP_MAX_BETA][2][2];  // [digit][component][coefficient]
#pragma unroll
  for (int j = 0; j < IP_MAX_BETA; ++j)
    if (j < a.beta) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const ulonglong2 kv = ld2(a.evk, (((size_t)j * 2 + c) * a.evk_limbs + kl) * n2 + i2);
        k[j][c][0] = u64_to_f64(kv.x); k[j][c][1] = u64_to_f64(kv.y);
      }
    }
  for (int b = 0; b < a.n_batch; ++b) {
    u64 *acc = a.acc + (size_t)b * a.acc_batch_stride;
    ulonglong2 t[IP_MAX_BETA];
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j) t[j] = tn[j];
    if (b + 1 < a.n_batch) {
#pragma unroll
      for (int j = 0; j < IP_MAX_BETA; ++j)
        if (j < a.beta) tn[j] = ld2(digit_ptr(b + 1, j), i2);
    }
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;  // [component][coefficient]
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j)
      if (j < a.beta) {
        const double t0 = u64_to_f64(t[j].x), t1 = u64_to_f64(t[j].y);
        s00 += mulmod_var(t0, k[j][0][0], m.q, m.qinv);
        s01 += mulmod_var(t1, k[j][0][1], m.q, m.qinv);
        s10 += mulmod_var(t0, k[j][1][0], m.q, m.qinv);
        s11 += mulmod_var(t1, k[j][1][1], m.q, m.qinv);
      }
    st2(acc, ((size_t)0 * a.n_ext + e) * n2 + i2, finish(s00, m), finish(s01, m));
    st2(acc, ((size_t)1 * a.n_ext + e) * n2 + i2, finish(s10, m), finish(s11, m));
  }
}

void launch_inner_product(const ModConst *mc, const LimbMap &lm, const InnerArgs &a, cudaStream_t s) {
  const dim3 g = ew_grid(a.N, a.n_ext);
  if (a.beta <= 1) k_inner<1><<<g, EW_THREADS, 0, s>>>(mc, lm, a);
  else if (a.beta <= 2) k_inner<2><<<g, EW_THREADS, 0, s>>>(mc, lm, a);
  else if (a.beta <= 3) k_inner<3><<<g, EW_THREADS, 0, s>>>(mc, lm, a);
  else if (a.beta <= 4) k_inner<4><<<g, EW_THREADS, 0, s>>>(mc, lm, a);
  else k_inner<8><<<g, EW_THREADS, 0, s>>>(mc, lm, a);
}

// ------------------------------------------------------------------------------------------------ (x - y) * c (+ z)
__global__ void __launch_bounds__(EW_THREADS) k_sub_mul_add(const ModConst *__restrict__ mc, LimbMap lm, SubMulArgs a) {
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= a.N / 2) return;
  const int limb = blockIdx.y, poly = blockIdx.z;
  const ModConst m = mc[lm.mod[limb]];
  const double2 c = a.cst[limb];
  const size_t o = (size_t)limb * (a.N / 2) + i2;
  const ulonglong2 x = ld2(a.x + poly * a.x_poly_stride, o), y = ld2(a.y + poly * a.y_poly_stride, o);
  double r0 = mulmod_const(u64_to_f64(x.x) - u64_to_f64(y.x), c.x, c.y, m.q);
  double r1 = mulmod_const(u64_to_f64(x.y) - u64_to_f64(y.y), c.x, c.y, m.q);
  if (a.z) {
    const ulonglong2 z = ld2(a.z + poly * a.z_poly_stride, o);
    r0 += u64_to_f64(z.x);
    r1 += u64_to_f64(z.y);
  }
  st2(a.out + poly * a.out_poly_stride, o, finish(r0, m), finish(r1, m));
}

void launch_sub_mul_add(const ModConst *mc, const LimbMap &lm, const SubMulArgs &a, cudaStream_t s) {
  k_sub_mul_add<<<ew_grid(a.N, a.n_limbs, a.n_polys), EW_THREADS, 0, s>>>(mc, lm, a);
}

// ------------------------------------------------------------------------------------------------ automorphism
__global__ void __launch_bounds__(EW_THREADS) k_automorph(int logN, const u64 *__restrict__ in, u64 *__restrict__ out, unsigned g) {
  const unsigned N = 1u << logN;
  const unsigned k = blockIdx.x * EW_THREADS + threadIdx.x;
  if (k >= N) return;
  const unsigned j = __brev(k) >> (32 - logN);
  const unsigned e = (g * (2u * j + 1u)) & (2u * N - 1u);  // low bits of the product are exact
  const unsigned ks = __brev((e - 1u) >> 1) >> (32 - logN);
  const size_t base = (size_t)blockIdx.y << logN;
  out[base + k] = __ldg(&in[base + ks]);
}

void launch_automorph(int logN, int n_limbs, const u64 *in, u64 *out, u64 g, cudaStream_t s) {
  const unsigned N = 1u << logN;
  k_automorph<<<dim3((N + EW_THREADS - 1) / EW_THREADS, n_limbs), EW_THREADS, 0, s>>>(logN, in, out, (unsigned)(g & (2ull * N - 1)));
}

// ------------------------------------------------------------------------------------------------ base conversion
// One CTA = 256 coefficients (two per thread) x OT output limbs.  The inner loop is 3 DFMA per modular MAC
// (12-bit-split matrix, exact double sums):
//   * input words come through a 4-deep register ring (loads for sources i+1..i+4 in flight while source i is
//     accumulated);
//   * the matrix operand does not go through the LSU: the whole [n_src][n_dst][3] matrix travels in
//     kernel-parameter space (constant bank, __grid_constant__, <= 30 KB), so each matrix word is a constant
//     load feeding DFMA.  (A shared-memory broadcast costs one LSU wavefront per load and made an earlier
//     version LSU-bound — profiles/README.md.)
//   * the epilogue is branch-free with its moduli in shared memory, so the 2*OT dependent reduction chains
//     interleave instead of running one after another behind a global load each.
// OT is chosen by the launcher so that n_dst splits without padding (35 = 5x7, 45 = 9x5, 50 = 10x5).
constexpr int BC_THREADS = 128;
constexpr int BC_RING = 4;

template <int OT, bool STEP1>
__global__ void __launch_bounds__(BC_THREADS, 4) k_bconv(const ModConst *__restrict__ mc, LimbMap src_lm, LimbMap dst_lm, BConvArgs a,
                                                         const __grid_constant__ BConvMatrix mat) {
  __shared__ double qs[OT], qinvs[OT];
  __shared__ u64 qis[OT];
  const int t0 = blockIdx.y * OT;
  const int nt = min(OT, a.n_dst - t0);
  if (threadIdx.x < OT) {
    const ModConst m = mc[dst_lm.mod[min(t0 + (int)threadIdx.x, a.n_dst - 1)]];  // a padded tail repeats the last modulus
    qs[threadIdx.x] = m.q; qinvs[threadIdx.x] = m.qinv; qis[threadIdx.x] = m.qi;
  }
  const int i2 = min(blockIdx.x * BC_THREADS + threadIdx.x, (unsigned)(a.N / 2 - 1));
  const bool live = blockIdx.x * BC_THREADS + threadIdx.x < (unsigned)(a.N / 2);
  const size_t n2 = a.N / 2;
  const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(a.in + (size_t)blockIdx.z * a.in_batch_stride) + i2;
  u64 *out = a.out + (size_t)blockIdx.z * a.out_batch_stride;
  ulonglong2 ring[BC_RING];
#pragma unroll
  for (int k = 0; k < BC_RING; ++k) ring[k] = __ldg(in + (size_t)src_lm.pos[min(k, a.n_src - 1)] * n2);
  __syncthreads();
  double acc[OT][3][2];
#pragma unroll
  for (int o = 0; o < OT; ++o)
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[o][k][0] = acc[o][k][1] = 0.0;
  for (int i0 = 0; i0 < a.n_src; i0 += BC_RING) {
#pragma unroll
    for (int k = 0; k < BC_RING; ++k) {
      const int i = i0 + k;
      if (i < a.n_src) {
        double y0 = u64_to_f64(ring[k].x), y1 = u64_to_f64(ring[k].y);
        ring[k] = __ldg(in + (size_t)src_lm.pos[min(i + BC_RING, a.n_src - 1)] * n2);
        if (STEP1) {
          const ModConst m = mc[src_lm.mod[i]];
          const double2 sc = a.step1[i];
          y0 = canonicalize(mulmod_const(y0, sc.x, sc.y, m.q), m.q);
          y1 = canonicalize(mulmod_const(y1, sc.x, sc.y, m.q), m.q);
        }
        const int hb = (i * a.n_dst + t0) * 3;  // uniform: matrix row of source i, first output of this tile
#pragma unroll
        for (int o = 0; o < OT; ++o)
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const double hv = mat.h[hb + o * 3 + kk];  // the struct's padding keeps a partial last tile in bounds
            acc[o][kk][0] = __fma_rn(y0, hv, acc[o][kk][0]);
            acc[o][kk][1] = __fma_rn(y1, hv, acc[o][kk][1]);
          }
        // each term is < 2^36 * 2^12; fold every 16 sources so the exact sums stay below 2^53
        if ((i & 15) == 15 && i + 1 < a.n_src) {
#pragma unroll
          for (int o = 0; o < OT; ++o)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              acc[o][kk][0] = reduce_signed(acc[o][kk][0], qs[o], qinvs[o]);
              acc[o][kk][1] = reduce_signed(acc[o][kk][1], qs[o], qinvs[o]);
            }
        }
      }
    }
  }
  // branch-free epilogue: all 2*OT reduction chains are independent and interleave; only the store is predicated
  u64 r[OT][2];
#pragma unroll
  for (int o = 0; o < OT; ++o) {
    const double q = qs[o], qinv = qinvs[o];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      // value = S2*2^24 + S1*2^12 + S0 (mod q), folded top-down; every intermediate is an exact integer < 2^53
      double v = reduce_signed(acc[o][2][c], q, qinv);
      v = reduce_signed(__fma_rn(v, 4096.0, acc[o][1][c]), q, qinv);
      v = reduce_signed(__fma_rn(v, 4096.0, acc[o][0][c]), q, qinv);
      r[o][c] = f64_to_canonical(v, qis[o]);
    }
  }
  if (!live) return;
#pragma unroll
  for (int o = 0; o < OT; ++o) {
    if (o < nt) {
      st2(out, (size_t)dst_lm.pos[t0 + o] * n2 + i2, r[o][0], r[o][1]);
    }
  }
}

int bconv_tile_height(int n_dst) {
  // the tile height that wastes the fewest padded output limbs (ties -> the taller tile)
  int best = 7, waste = 1 << 30;
  for (int ot : {7, 6, 5}) {
    const int w = (n_dst + ot - 1) / ot * ot - n_dst;
    if (w < waste) { waste = w; best = ot; }
  }
  return best;
}

template <int OT>
static void launch_bconv_t(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a,
                           const BConvMatrix &mat, cudaStream_t s) {
  const dim3 grid((a.N / 2 + BC_THREADS - 1) / BC_THREADS, (a.n_dst + OT - 1) / OT, a.n_batches);
  if (a.step1) k_bconv<OT, true><<<grid, BC_THREADS, 0, s>>>(mc, src_lm, dst_lm, a, mat);
  else k_bconv<OT, false><<<grid, BC_THREADS, 0, s>>>(mc, src_lm, dst_lm, a, mat);
}

void launch_bconv(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a, const BConvMatrix &mat,
                  cudaStream_t s) {
  switch (bconv_tile_height(a.n_dst)) {
    case 7: launch_bconv_t<7>(mc, src_lm, dst_lm, a, mat, s); break;
    case 6: launch_bconv_t<6>(mc, src_lm, dst_lm, a, mat, s); break;
    default: launch_bconv_t<5>(mc, src_lm, dst_lm, a, mat, s); break;
  }
}

}  // namespace hml
