// See ewe.cuh.  All kernels are streaming (HBM-bound) except base conversion (FP64 tensor-core path, pipe bound).
#include "ewe.cuh"
#include "launch.h"

#include <algorithm>
#include <cstdlib>

namespace hml {

constexpr int EW_THREADS = 256;  // 2 coefficients per thread, 16-byte accesses

__device__ __forceinline__ ulonglong2 ld2(const u64 *p, size_t i2) { return __ldg(reinterpret_cast<const ulonglong2 *>(p) + i2); }
__device__ __forceinline__ void st2(u64 *p, size_t i2, u64 a, u64 b) { reinterpret_cast<ulonglong2 *>(p)[i2] = make_ulonglong2(a, b); }
__device__ __forceinline__ u64 finish(double v, const ModConst &m) { return f64_to_canonical(reduce_signed(v, m.q, m.qinv), m.qi); }

static inline dim3 ew_grid(int N, int ny, int nz = 1) { return dim3((N / 2 + EW_THREADS - 1) / EW_THREADS, ny, nz); }

// ------------------------------------------------------------------------------------------------ generic EWE
// grid.z = component of a ciphertext op: operand k advances by cs[k] words per component (0 repeats a plaintext), out by cs[4]
struct EweStrides {
  long long cs[5];
};
__global__ void __launch_bounds__(EW_THREADS) k_ewe(const ModConst *__restrict__ mc, LimbMap lm, int N, const u64 *x1,
                                                    const u64 *x2, const u64 *x3, const u64 *x4, int subtract, u64 *out,
                                                    EweStrides st) {
  pdl_wait();
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= N / 2) return;
  const int limb = blockIdx.y;
  const ModConst m = mc[lm.mod[limb]];
  const size_t o = (size_t)limb * (N / 2) + i2;
  const long long z = blockIdx.z;
  if (x1) x1 += z * st.cs[0];
  if (x2) x2 += z * st.cs[1];
  if (x3) x3 += z * st.cs[2];
  if (x4) x4 += z * st.cs[3];
  out += z * st.cs[4];
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
                const u64 *x4, int subtract, u64 *out, cudaStream_t s, int n_comp, const long long *comp_strides) {
  EweStrides st{};
  if (comp_strides) for (int k = 0; k < 5; ++k) st.cs[k] = comp_strides[k];
  launch_pdl(k_ewe, ew_grid(N, n_limbs, n_comp), EW_THREADS, 0, s, mc, lm, N, x1, x2, x3, x4, subtract, out, st);
}

// ------------------------------------------------------------------------------------------------ tensor product
__global__ void __launch_bounds__(EW_THREADS) k_tensor3(const ModConst *__restrict__ mc, int N, const u64 *a0, const u64 *a1,
                                                        const u64 *b0, const u64 *b1, u64 *d0, u64 *d1, u64 *d2,
                                                        long long in_stride, long long out_stride, int pack01) {
  pdl_wait();
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
  if (pack01) {
    st_packed2(d0 + (size_t)limb * N, N, i2, r0[0], r0[1]);
    st_packed2(d1 + (size_t)limb * N, N, i2, r1[0], r1[1]);
  } else {
    st2(d0, o, r0[0], r0[1]);
    st2(d1, o, r1[0], r1[1]);
  }
  st2(d2, o, r2[0], r2[1]);
}

void launch_tensor3(const ModConst *mc, int N, int L, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *d0,
                    u64 *d1, u64 *d2, int n_batch, long long in_stride, long long out_stride, cudaStream_t s, int pack01) {
  launch_pdl(k_tensor3, ew_grid(N, L, n_batch), EW_THREADS, 0, s, mc, N, a0, a1, b0, b1, d0, d1, d2, in_stride, out_stride, pack01);
}

// ------------------------------------------------------------------------------------------------ key-switch inner product
// The key words of a (limb, coefficient pair) are loaded once and reused for every ciphertext of the batch
// (the key is 60% of the traffic of an unbatched inner product).
// An empty asm statement that "uses" every loaded register: all the loads above it must have been ISSUED before any
// instruction below it (no instruction is emitted, so nothing waits here — the scoreboard stalls at the first real use).
#define HML_PIN4(v) "+l"(v[0].x), "+l"(v[0].y), "+l"(v[1].x), "+l"(v[1].y)
template <int NB>
__device__ __forceinline__ void pin_loads(ulonglong2 (&k)[NB][2]) {
  if constexpr (NB == 1) asm volatile("" : HML_PIN4(k[0]));
  else if constexpr (NB == 2) asm volatile("" : HML_PIN4(k[0]), HML_PIN4(k[1]));
  else if constexpr (NB == 3) asm volatile("" : HML_PIN4(k[0]), HML_PIN4(k[1]), HML_PIN4(k[2]));
  else if constexpr (NB == 4) asm volatile("" : HML_PIN4(k[0]), HML_PIN4(k[1]), HML_PIN4(k[2]), HML_PIN4(k[3]));
  else {
    asm volatile("" : HML_PIN4(k[0]), HML_PIN4(k[1]), HML_PIN4(k[2]), HML_PIN4(k[3]));
    asm volatile("" : HML_PIN4(k[4]), HML_PIN4(k[5]), HML_PIN4(k[6]), HML_PIN4(k[7]));
  }
}

// MULTI: more than one ciphertext per launch (the next ciphertext's digits are prefetched while this one is multiplied)
template <int IP_MAX_BETA, bool MULTI>
__global__ void __launch_bounds__(EW_THREADS, 3) k_inner(const ModConst *__restrict__ mc, LimbMap lm, InnerArgs a) {
  pdl_wait();
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
  // source slots of this thread's two coefficients under the automorphism (identity when galois == 0)
  unsigned s0 = 2u * i2, s1 = 2u * i2 + 1u;
  const bool gal = !MULTI && a.galois != 0;  // the batched variant stays lean: no automorphism on load, no u side output
  if (gal) {
    const unsigned sh = 32 - a.logN, m2n = (2u << a.logN) - 1u;
    s0 = __brev((((a.galois * (2u * (__brev(s0) >> sh) + 1u)) & m2n) - 1u) >> 1) >> sh;
    s1 = __brev((((a.galois * (2u * (__brev(s1) >> sh) + 1u)) & m2n) - 1u) >> 1) >> sh;
  }
  auto ld_digit = [&](int b, int j) -> ulonglong2 {
    const u64 *p = digit_ptr(b, j);
    if (!gal) return ld2(p, i2);
    return make_ulonglong2(__ldg(p + s0), __ldg(p + s1));
  };
  ulonglong2 tn[IP_MAX_BETA];
#pragma unroll
  for (int j = 0; j < IP_MAX_BETA; ++j)
    if (j < a.beta) tn[j] = ld_digit(0, j);
  // every key word of the thread is requested before the first one is used: ONE memory round trip for the 3 * beta loads
  // (converting inside the load loop made ptxas wait for each digit's pair before issuing the next)
  ulonglong2 kraw[IP_MAX_BETA][2];
  if (a.evk_packed) {  // uniform: packed key limbs — the raw halves are all requested first, then assembled (kraw.x = the pair's low
    // words, kraw.y = its two high bytes until the loop below)
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const u64 *slot = a.evk + (((size_t)j * 2 + c) * a.evk_limbs + kl) * a.N;
        kraw[j][c] = j < a.beta ? make_ulonglong2(__ldg(reinterpret_cast<const u64 *>(slot) + i2),
                                                  __ldg(reinterpret_cast<const unsigned short *>(reinterpret_cast<const unsigned char *>(slot) + 4 * (size_t)a.N) + i2))
                                : make_ulonglong2(0, 0);
      }
    pin_loads(kraw);
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const u64 lo = kraw[j][c].x, hi = kraw[j][c].y;
        kraw[j][c] = make_ulonglong2(((hi & 0xFFull) << 32) | (lo & 0xFFFFFFFFull), ((hi >> 8) << 32) | (lo >> 32));
      }
  } else {
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c)
        kraw[j][c] = j < a.beta ? ld2(a.evk, (((size_t)j * 2 + c) * a.evk_limbs + kl) * n2 + i2) : make_ulonglong2(0, 0);
    pin_loads(kraw);
  }
  double k[IP_MAX_BETA][2][2];  // [digit][component][coefficient]
#pragma unroll
  for (int j = 0; j < IP_MAX_BETA; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      k[j][c][0] = u64_to_f64(kraw[j][c].x); k[j][c][1] = u64_to_f64(kraw[j][c].y);
    }
  for (int b = 0; b < a.n_batch; ++b) {
    u64 *acc = a.acc + (size_t)b * a.acc_batch_stride;
    ulonglong2 t[IP_MAX_BETA];
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j) t[j] = tn[j];
    if (MULTI && b + 1 < a.n_batch) {
#pragma unroll
      for (int j = 0; j < IP_MAX_BETA; ++j)
        if (j < a.beta) tn[j] = ld_digit(b + 1, j);
    }
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;  // [component][coefficient]
#pragma unroll
    for (int j = 0; j < IP_MAX_BETA; ++j)
      if (j < a.beta) {
        // converted digits may arrive as raw lazy doubles straight from the transform (own-digit limbs are always words)
        const bool f64 = a.ext_f64 && j != own;
        const double t0 = f64 ? __longlong_as_double((long long)t[j].x) : u64_to_f64(t[j].x);
        const double t1 = f64 ? __longlong_as_double((long long)t[j].y) : u64_to_f64(t[j].y);
        s00 += mulmod_var(t0, k[j][0][0], m.q, m.qinv);
        s01 += mulmod_var(t1, k[j][0][1], m.q, m.qinv);
        s10 += mulmod_var(t0, k[j][1][0], m.q, m.qinv);
        s11 += mulmod_var(t1, k[j][1][1], m.q, m.qinv);
      }
    const size_t comp2 = (a.acc_comp_stride ? (size_t)a.acc_comp_stride : (size_t)a.n_ext * a.N) / 2;
    const u64 r00 = finish(s00, m), r01 = finish(s01, m), r10 = finish(s10, m), r11 = finish(s11, m);
    if (e < a.acc_pack_limbs) {  // uniform per CTA
      st_packed2(acc + (size_t)e * a.N, a.N, i2, r00, r01);
      st_packed2(acc + 2 * comp2 + (size_t)e * a.N, a.N, i2, r10, r11);
    } else {
      st2(acc, (size_t)e * n2 + i2, r00, r01);
      st2(acc, comp2 + (size_t)e * n2 + i2, r10, r11);
    }
    if (!MULTI && e == a.u_limb) {  // uniform per CTA: u = acc * P^-1 + d on the limb the rescale drops (no extra launch, no re-read)
      const u64 *ua = a.u_add + (size_t)b * a.u_add_batch_stride + (size_t)e * a.N;
      const ulonglong2 d0v = ld_packed2(ua, a.N, i2), d1v = ld_packed2(ua + a.u_add_comp_stride, a.N, i2);
      const double2 c = a.u_cst;
      st2(acc, (size_t)a.u_slot * n2 + i2, finish(mulmod_const(u64_to_f64(r00), c.x, c.y, m.q) + u64_to_f64(d0v.x), m),
          finish(mulmod_const(u64_to_f64(r01), c.x, c.y, m.q) + u64_to_f64(d0v.y), m));
      st2(acc, comp2 + (size_t)a.u_slot * n2 + i2, finish(mulmod_const(u64_to_f64(r10), c.x, c.y, m.q) + u64_to_f64(d1v.x), m),
          finish(mulmod_const(u64_to_f64(r11), c.x, c.y, m.q) + u64_to_f64(d1v.y), m));
    }
  }
}

template <bool MULTI>
static void launch_inner_t(const ModConst *mc, const LimbMap &lm, const InnerArgs &a, cudaStream_t s) {
  const dim3 g = ew_grid(a.N, a.n_ext);
  if (a.beta <= 1) launch_pdl(k_inner<1, MULTI>, g, EW_THREADS, 0, s, mc, lm, a);
  else if (a.beta <= 2) launch_pdl(k_inner<2, MULTI>, g, EW_THREADS, 0, s, mc, lm, a);
  else if (a.beta <= 3) launch_pdl(k_inner<3, MULTI>, g, EW_THREADS, 0, s, mc, lm, a);
  else if (a.beta <= 4) launch_pdl(k_inner<4, MULTI>, g, EW_THREADS, 0, s, mc, lm, a);
  else launch_pdl(k_inner<8, MULTI>, g, EW_THREADS, 0, s, mc, lm, a);
}
void launch_inner_product(const ModConst *mc, const LimbMap &lm, const InnerArgs &a, cudaStream_t s) {
  if (a.n_batch > 1) launch_inner_t<true>(mc, lm, a, s);
  else launch_inner_t<false>(mc, lm, a, s);
}

// ------------------------------------------------------------------------------------------------ (x - y) * c (+ z)
__global__ void __launch_bounds__(EW_THREADS) k_sub_mul_add(const ModConst *__restrict__ mc, LimbMap lm, SubMulArgs a) {
  pdl_wait();
  const int i2 = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= a.N / 2) return;
  const int limb = blockIdx.y, poly = blockIdx.z;
  const ModConst m = mc[lm.mod[limb]];
  const double2 c = a.cst[limb];
  const size_t o = (size_t)limb * (a.N / 2) + i2;
  const ulonglong2 x = a.x_packed ? ld_packed2(a.x + poly * a.x_poly_stride + (size_t)limb * a.N, a.N, i2) : ld2(a.x + poly * a.x_poly_stride, o);
  const ulonglong2 y = a.y ? ld2(a.y + poly * a.y_poly_stride, o) : make_ulonglong2(0, 0);
  double r0 = mulmod_const(u64_to_f64(x.x) - u64_to_f64(y.x), c.x, c.y, m.q);
  double r1 = mulmod_const(u64_to_f64(x.y) - u64_to_f64(y.y), c.x, c.y, m.q);
  if (a.z && (a.z_mask == 0 || ((a.z_mask >> poly) & 1u))) {
    const ulonglong2 z = a.z_packed ? ld_packed2(a.z + poly * a.z_poly_stride + (size_t)limb * a.N, a.N, i2) : ld2(a.z + poly * a.z_poly_stride, o);
    r0 += u64_to_f64(z.x);
    r1 += u64_to_f64(z.y);
  }
  st2(a.out + poly * a.out_poly_stride, o, finish(r0, m), finish(r1, m));
}

void launch_sub_mul_add(const ModConst *mc, const LimbMap &lm, const SubMulArgs &a, cudaStream_t s) {
  launch_pdl(k_sub_mul_add, ew_grid(a.N, a.n_limbs, a.n_polys), EW_THREADS, 0, s, mc, lm, a);
}

// ------------------------------------------------------------------------------------------------ automorphism
__global__ void __launch_bounds__(EW_THREADS) k_automorph(int logN, const u64 *__restrict__ in, u64 *__restrict__ out, unsigned g) {
  pdl_wait();
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
  launch_pdl(k_automorph, dim3((N + EW_THREADS - 1) / EW_THREADS, n_limbs), EW_THREADS, 0, s, logN, in, out, (unsigned)(g & (2ull * N - 1)));
}

// words -> packed limbs, slot geometry unchanged (evaluation keys: hml_key_pack)
__global__ void __launch_bounds__(EW_THREADS) k_pack_limbs(int N, const u64 *__restrict__ in, u64 *__restrict__ out) {
  pdl_wait();
  const size_t i2 = (size_t)blockIdx.x * EW_THREADS + threadIdx.x;
  if (i2 >= (size_t)N / 2) return;
  const size_t limb = blockIdx.y + (size_t)blockIdx.z * gridDim.y;
  const ulonglong2 v = ld2(in + limb * N, i2);
  st_packed2(out + limb * N, N, i2, v.x, v.y);
}
void launch_pack_limbs(int N, size_t n_limbs, const u64 *in, u64 *out, cudaStream_t s) {
  for (size_t done = 0; done < n_limbs;) {  // grid.y * grid.z limbs per launch
    const size_t rest = n_limbs - done, gy = rest < 256 ? rest : 256, gz = std::min<size_t>(rest / gy, 65535);
    launch_pdl(k_pack_limbs, dim3((unsigned)((N / 2 + EW_THREADS - 1) / EW_THREADS), (unsigned)gy, (unsigned)gz), EW_THREADS, 0, s, N, in + done * N, out + done * N);
    done += gy * gz;
  }
}

// ------------------------------------------------------------------------------------------------ base conversion
// out[t][m] = sum_i y_i[m] * H[i][t]  (mod q_t)  is a matrix product [coefficients x sources] x [sources x targets]
// with exact integer entries.  It runs on the FP64 tensor-core path (DMMA m8n8k4): H is split into three 12-bit pieces
// so that every product (36 + 12 bits) and every 16-term partial sum (<= 52 bits) is an exactly representable double,
// whatever the accumulation order inside the MMA.  Measured on B200 (profiles/microbench/dmma_bench.cu): DMMA reaches
// the FP64 pipe's full 18.5 T FMA/s with ONE instruction per 256 FMAs, and integer / load instructions issue in its
// shadow, whereas a DFMA holds the issue port for two cycles per 32 FMAs — the earlier DFMA kernel spent 25% of its
// issue slots on matrix-operand loads alone.
//   CTA  = `tm` coefficients (256, or 64 when there are many sources) x all targets; the sources of the tile are
//          staged once in shared memory as doubles, pitch tm + 4 (conflict-free fragment reads), zero rows as padding
//   warp = one block of 8 targets: B fragments (sources x 8 targets x 3 pieces) live in registers for <= 16 sources
//          (KS = number of 4-source k-steps), else they are re-read from the L1-resident matrix and the accumulators are
//          folded mod q_t every 16 sources
//   lane = coefficient lane/4 of each 8-coefficient m-tile, targets 2*(lane%4), +1 of the block: the three pieces of a
//          target meet in one lane, so the epilogue (Horner over the pieces, 3 reductions) needs no exchange
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int KS, bool STEP1>
__global__ void __launch_bounds__(512) k_bconv_mma(const ModConst *__restrict__ mc, LimbMap src_lm, LimbMap dst_lm, BConvArgs a,
                                                   const double *__restrict__ mat, int n_src_pad, int n_dst_pad, int tm) {
  extern __shared__ __align__(16) double ys[];  // [n_src_pad][tm + 4]
  const int pitch = tm + 4;
  const int lane = threadIdx.x & 31, tb = threadIdx.x >> 5;
  const size_t m_base = (size_t)blockIdx.x * tm;
  const u64 *in = a.in + (size_t)blockIdx.y * a.in_batch_stride;
  u64 *out = a.out + (size_t)blockIdx.y * a.out_batch_stride;
  // matrix fragments and moduli are constant tables: fetched before the programmatic dependency is resolved
  const bool mma_warp = tb * 8 < n_dst_pad;  // a launch with fewer than four target blocks has staging-only helper warps
  const int kq = lane & 3, rq = lane >> 2;
  const int t0 = tb * 8 + 2 * kq, t1 = t0 + 1;
  const bool live0 = t0 < a.n_dst, live1 = t1 < a.n_dst;
  const ModConst m0 = mc[dst_lm.mod[live0 ? t0 : a.n_dst - 1]], m1 = mc[dst_lm.mod[live1 ? t1 : a.n_dst - 1]];
  u64 *o0 = out + (size_t)dst_lm.pos[live0 ? t0 : a.n_dst - 1] * a.N + m_base + rq;
  u64 *o1 = out + (size_t)dst_lm.pos[live1 ? t1 : a.n_dst - 1] * a.N + m_base + rq;
  const double *bsrc = mat + ((size_t)kq * n_dst_pad + tb * 8 + rq) * 3;  // B[k = lane%4][n = lane/4], k-step stride 4 rows
  const size_t brow = (size_t)4 * n_dst_pad * 3;
  double bf[KS > 0 ? KS : 1][3];
  if constexpr (KS > 0) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int p = 0; p < 3; ++p) bf[ks][p] = mma_warp ? __ldg(bsrc + ks * brow + p) : 0.0;
  }
  pdl_wait();  // the sources are another kernel's output
  {  // stage the tile's sources (step 1 applied here when requested), zero the padding rows; four loads in flight per thread
    const int half = tm >> 1, total = n_src_pad * half;
    for (int e0 = threadIdx.x; e0 < total; e0 += 4 * blockDim.x) {
      ulonglong2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * blockDim.x, i = e / half, m2 = e - i * half;
        v[k] = make_ulonglong2(0, 0);
        if (e < total && i < a.n_src) v[k] = __ldg(reinterpret_cast<const ulonglong2 *>(in + (size_t)src_lm.pos[i] * a.N + m_base) + m2);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * blockDim.x, i = e / half, m2 = e - i * half;
        if (e >= total) break;
        double y0 = u64_to_f64(v[k].x), y1 = u64_to_f64(v[k].y);
        if (STEP1 && i < a.n_src) {
          const ModConst m = mc[src_lm.mod[i]];
          const double2 sc = a.step1[i];
          y0 = canonicalize(mulmod_const(y0, sc.x, sc.y, m.q), m.q);
          y1 = canonicalize(mulmod_const(y1, sc.x, sc.y, m.q), m.q);
        }
        *reinterpret_cast<double2 *>(&ys[i * pitch + 2 * m2]) = make_double2(y0, y1);
      }
    }
  }
  __syncthreads();
  if (a.fold) {  // uniform branch: every thread of the CTA takes it
    const ModConst mf = mc[a.fold_mod];
    const int last = a.n_src - 1;
    for (int m = threadIdx.x; m < tm; m += blockDim.x) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;  // terms of < 2^48 each: exact as long as at most 16 are summed unreduced
      for (int i = 0; i < last; ++i) {
        const double y = ys[i * pitch + m];
        s0 = __fma_rn(y, __ldg(a.fold + 3 * i), s0);
        s1 = __fma_rn(y, __ldg(a.fold + 3 * i + 1), s1);
        s2 = __fma_rn(y, __ldg(a.fold + 3 * i + 2), s2);
        if ((i & 15) == 15) {  // alpha > 16 (DMMA path only): fold like the main loop does, every 16 sources
          s0 = reduce_signed(s0, mf.q, mf.qinv);
          s1 = reduce_signed(s1, mf.q, mf.qinv);
          s2 = reduce_signed(s2, mf.q, mf.qinv);
        }
      }
      double v = reduce_signed(s2, mf.q, mf.qinv);
      v = reduce_signed(__fma_rn(v, 4096.0, s1), mf.q, mf.qinv);
      v = reduce_signed(__fma_rn(v, 4096.0, s0 + ys[last * pitch + m]), mf.q, mf.qinv);
      ys[last * pitch + m] = canonicalize(v, mf.q);
    }
    __syncthreads();
  }
  if (!mma_warp) return;
  const double nh0 = a.out_f64 ? 0.0 : -(m0.q - 1.0) * 0.5, nh1 = a.out_f64 ? 0.0 : -(m1.q - 1.0) * 0.5;
  const double hb0 = 4503599627370496.0 - nh0, hb1 = 4503599627370496.0 - nh1;  // 2^52 + h
  const double *yl = ys + kq * pitch + rq;  // A[row = lane/4][col = lane%4]
  const int n_ks = n_src_pad >> 2;
  // two m-tiles (16 coefficients) per iteration: six independent accumulator chains keep the tensor pipe fed
  for (int mt = 0; mt < (tm >> 3); mt += 2) {
    // piece-0 accumulators start at -h, h = (q_t - 1) / 2: the final centred remainder is then (value - h) in [-h, h] and
    // canonicalisation is one add of 2^52 + h (no sign fix-up)
    double acc[2][3][2] = {{{nh0, nh1}, {0.0, 0.0}, {0.0, 0.0}}, {{nh0, nh1}, {0.0, 0.0}, {0.0, 0.0}}};
    if constexpr (KS > 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double af0 = yl[ks * 4 * pitch + mt * 8], af1 = yl[ks * 4 * pitch + mt * 8 + 8];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          dmma884(acc[0][p][0], acc[0][p][1], af0, bf[ks][p]);
          dmma884(acc[1][p][0], acc[1][p][1], af1, bf[ks][p]);
        }
      }
    } else {
      for (int ks = 0; ks < n_ks; ++ks) {
        const double af0 = yl[ks * 4 * pitch + mt * 8], af1 = yl[ks * 4 * pitch + mt * 8 + 8];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const double b = __ldg(bsrc + ks * brow + p);
          dmma884(acc[0][p][0], acc[0][p][1], af0, b);
          dmma884(acc[1][p][0], acc[1][p][1], af1, b);
        }
        // each term is < 2^36 * 2^12: fold every 16 sources so the exact sums stay below 2^53
        if ((ks & 3) == 3 && ks + 1 < n_ks) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              acc[h][p][0] = reduce_signed(acc[h][p][0], m0.q, m0.qinv);
              acc[h][p][1] = reduce_signed(acc[h][p][1], m1.q, m1.qinv);
            }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // value = S2*2^24 + S1*2^12 + S0 (mod q), folded top-down; every intermediate is an exact integer < 2^53
      double v0 = reduce_signed(acc[h][2][0], m0.q, m0.qinv), v1 = reduce_signed(acc[h][2][1], m1.q, m1.qinv);
      v0 = reduce_signed(__fma_rn(v0, 4096.0, acc[h][1][0]), m0.q, m0.qinv);
      v1 = reduce_signed(__fma_rn(v1, 4096.0, acc[h][1][1]), m1.q, m1.qinv);
      v0 = reduce_signed(__fma_rn(v0, 4096.0, acc[h][0][0]), m0.q, m0.qinv);
      v1 = reduce_signed(__fma_rn(v1, 4096.0, acc[h][0][1]), m1.q, m1.qinv);
      if (a.out_f64) {  // uniform
        if (live0) reinterpret_cast<double *>(o0)[(mt + h) * 8] = v0;
        if (live1) reinterpret_cast<double *>(o1)[(mt + h) * 8] = v1;
      } else {
        if (live0) o0[(mt + h) * 8] = (u64)__double_as_longlong(v0 + hb0) & 0x000FFFFFFFFFFFFFull;
        if (live1) o1[(mt + h) * 8] = (u64)__double_as_longlong(v1 + hb1) & 0x000FFFFFFFFFFFFFull;
      }
    }
  }
}

template <int KS>
static void launch_bconv_t(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a, const double *mat,
                           int n_src_pad, int n_dst_pad, int tm, cudaStream_t s) {
  const dim3 grid(a.N / tm, a.n_batches);
  const int threads = std::max(128, 32 * (n_dst_pad / 8));  // at least four warps so that staging has loads in flight
  const size_t smem = (size_t)n_src_pad * (tm + 4) * sizeof(double);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(k_bconv_mma<KS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 68 * 8);
    cudaFuncSetAttribute(k_bconv_mma<KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 68 * 8);
  }
  if (a.step1) launch_pdl(k_bconv_mma<KS, true>, grid, threads, smem, s, mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm);
  else launch_pdl(k_bconv_mma<KS, false>, grid, threads, smem, s, mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm);
}

int bconv_umma_enabled() {
  static const int v = [] {
    const char *e = getenv("HML_BCONV_UMMA");
    return e && atoi(e) == 0 ? 0 : 1;
  }();
  return v;
}

void launch_bconv(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a, const double *mat, cudaStream_t s,
                  const BConvImage *im) {
  if (im && im->img && a.N >= 128 && a.N % 128 == 0 && (a.fold != nullptr) == (im->fold != 0) && bconv_umma_enabled()) {
    launch_bconv_umma(mc, src_lm, dst_lm, a, *im, s);
    return;
  }
  const int n_src_pad = bconv_pad_src(a.n_src), n_dst_pad = bconv_pad_dst(a.n_dst);
  static const int tm_env = [] {
    const char *e = getenv("HML_BCONV_TM");  // tuning knob: coefficients per CTA (<= 16 sources); 0 = adaptive
    const int v = e ? atoi(e) : 0;
    return (v == 64 || v == 128 || v == 256) ? v : 0;
  }();
  // big tiles amortise the per-CTA set-up (matrix fragments, moduli); small launches need more, shorter CTAs to fill 148 SMs
  int tm = 64;
  if (n_src_pad <= 16) {
    tm = tm_env ? tm_env : 256;
    while (!tm_env && tm > 64 && (long long)(a.N / tm) * a.n_batches < 1480) tm >>= 1;
  }
  while (tm > a.N) tm >>= 1;  // tiny rings (tests): N >= 16
  switch (n_src_pad <= 16 ? n_src_pad / 4 : 0) {
    case 1: launch_bconv_t<1>(mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm, s); break;
    case 2: launch_bconv_t<2>(mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm, s); break;
    case 3: launch_bconv_t<3>(mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm, s); break;
    case 4: launch_bconv_t<4>(mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm, s); break;
    default: launch_bconv_t<0>(mc, src_lm, dst_lm, a, mat, n_src_pad, n_dst_pad, tm, s); break;
  }
}

}  // namespace hml
