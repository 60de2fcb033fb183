// Base conversion on the 5th-generation tensor cores: tcgen05.mma kind::i8 (u8 x u8 -> s32), accumulators in TMEM.
//
// Reference instruction class: BCONV_STEP2, InsGen::GenBCONV reference src/InsGen.cpp:263-313 (simulated unit BCONVU,
// reference src/Components.cpp:268-362); call sites KeySwitch::ModUpBConvStep2 reference src/Operation.cpp:137-188 and
// KeySwitch::ModDownBConvStep2 :489-519.
//
//   out[t][m] = sum_i y_i[m] * H[i][t]   (mod q_t),      y_i < 2^36,  H[i][t] < q_t < 2^36
//
// is an integer matrix product [coefficients x sources] x [sources x targets].  Integer split: y_i = sum_a Y_ia 2^(8a)
// (five bytes), and for every byte position a the host precomputes H_a[i][t] = H[i][t] * 2^(8a) mod q_t and splits THAT
// into five bytes B_b.  Then
//   out[t] = sum_b 2^(8b) T_b[t],     T_b[t] = sum_{(i,a)} Y_ia * byte_b(H_a[i][t])
// so ONE u8 x u8 -> s32 GEMM with K = 5 * n_src and five s32 columns per target does all the multiply work, and the
// epilogue is a 5-term shift-add (exact: < 2^52 for <= 48 sources, because byte 4 of a 36-bit value has 4 bits) plus a
// single reduction mod q_t on the FP64 pipe.  SURVEY.md / BASELINE.json north_star: "tensor cores are used only if an
// integer-split BConv beats CUDA cores in ncu" — measured numbers in profiles/README.md.
//
// Kernel shape (one persistent CTA per SM, 512 threads, no warp specialisation, one __syncthreads per tile):
//   tile    128 coefficients (UMMA M = 128: TMEM lane = coefficient) x all targets (UMMA N = NP <= 256 s32 columns,
//           column = byte level b * ND + target) x K = 80 bytes per 16 sources, padded to a multiple of 32
//   A       [K/16][128 rows][16 B] in shared memory = the canonical K-major no-swizzle UMMA layout (core matrix = 8 rows x
//           16 B contiguous; SBO = 128 B between 8-row groups, LBO = 2048 B between 16-byte K chunks).  K order: chunk
//           (i/16)*5 + a holds byte a of sources 16*(i/16) .. +15, so a thread that holds four consecutive sources of one
//           coefficient emits one 32-bit word per byte position (3 PRMT) and a warp's 32 words are conflict-free
//   B       the host-built image, same layout with NP rows, copied once per CTA
//   D       two TMEM buffers of 256 columns: the MMA of tile n+1 runs while tile n's accumulators are drained
//   loop    pack(n+1) -> sync -> [thread 0: MMA(n+1), commit] -> global loads(n+2) in flight -> wait MMA(n) ->
//           epilogue(n): tcgen05.ld 4 targets x 5 levels, shift-add on the integer pipe, one FP64 reduction, 256-byte
//           coalesced stores per warp and target
// The optional fold of hmult's merged ModDown + Rescale (context.cu) is one more (virtual) target: its remainder r is
// computed first by every warp for its rows and added to every real target's sum before the reduction.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ewe.cuh"
#include "launch.h"

namespace hml {

// ------------------------------------------------------------------------------------------------ host: operand image
static inline u64 h_mulmod64(u64 a, u64 b, u64 q) { return (u64)((unsigned __int128)a * b % q); }

static inline int umma_k_index(int i, int a) { return ((i >> 4) * 5 + a) * 16 + (i & 15); }

bool bconv_image_shape(int n_src, int n_dst, int fold, BConvImage &im) {
  im = BConvImage{};
  if (n_src < 1 || n_src > 48 || n_dst < 1) return false;
  const int ntv = n_dst + (fold ? 1 : 0);
  const int ND = (ntv + 7) & ~7;
  const int NP = (5 * ND + 15) & ~15;
  if (NP > 256) return false;
  im.n16 = (n_src + 15) >> 4;
  im.K = (im.n16 * 80 + 31) & ~31;
  im.NP = NP; im.ND = ND; im.fold = fold ? 1 : 0;
  return true;
}

// hat [n_src][n_dst] row-major; dst_q[t] modulus of target t.  With fold != null the LAST source is the folded remainder's
// seed (see BConvArgs::fold): its matrix entries for the real targets are ignored (the remainder is added in the
// epilogue), fold[i] (i < n_src - 1) are the constants of the virtual target, modulus fold_q.
bool bconv_image_build(const u64 *hat, int n_src, int n_dst, const u64 *dst_q, const u64 *fold, u64 fold_q,
                       std::vector<uint8_t> &img, BConvImage &im) {
  if (!bconv_image_shape(n_src, n_dst, fold != nullptr, im)) return false;
  img.assign((size_t)im.K * im.NP, 0);
  const int ntv = n_dst + im.fold;
  for (int t = 0; t < ntv; ++t) {
    const bool virt = t == n_dst;
    const u64 q = virt ? fold_q : dst_q[t];
    if (q >> 36) return false;
    for (int i = 0; i < n_src; ++i) {
      u64 h;
      if (virt) h = i + 1 < n_src ? fold[i] % q : 1;
      else h = (fold && i + 1 == n_src) ? 0 : hat[(size_t)i * n_dst + t] % q;
      for (int a = 0; a < 5; ++a) {
        const u64 ha = h_mulmod64(h, (1ull << (8 * a)) % q, q);
        const int k = umma_k_index(i, a);
        for (int b = 0; b < 5; ++b) {
          const int n = b * im.ND + t;
          img[(size_t)(k >> 4) * im.NP * 16 + (size_t)n * 16 + (k & 15)] = (uint8_t)(ha >> (8 * b));
        }
      }
    }
  }
  return true;
}

// CPU model of the kernel's data path (A packing order, image addressing, epilogue arithmetic), used by the CPU tests to
// check the host-built image and the exactness argument without a GPU.  y [n_src][M] canonical, out [n_dst][M] canonical.
extern "C" int hml_dbg_bconv_umma_model(const uint64_t *hat, int n_src, int n_dst, const uint64_t *dst_q, const uint64_t *fold,
                                        uint64_t fold_q, const uint64_t *y, int M, uint64_t *out) {
  BConvImage im;
  std::vector<uint8_t> img;
  if (!bconv_image_build((const u64 *)hat, n_src, n_dst, (const u64 *)dst_q, (const u64 *)fold, fold_q, img, im)) return 1;
  std::vector<uint8_t> arow(im.K);
  std::vector<int32_t> T(im.NP);
  auto value = [&](int t) {  // exact integer sum_b 2^(8b) T_b[t] as the epilogue forms it
    u64 acc = ((u64)(0x43300000u + (uint32_t)T[4 * im.ND + t]) << 32) | (uint32_t)T[t];
    acc += (u64)(uint32_t)T[1 * im.ND + t] << 8;
    acc += (u64)(uint32_t)T[2 * im.ND + t] << 16;
    acc += (u64)(uint32_t)T[3 * im.ND + t] << 24;
    double d;
    memcpy(&d, &acc, 8);
    return d - 4503599627370496.0;
  };
  auto reduce = [](double v, double q) {
    const double qinv = 1.0 / q;
    const double qh = __builtin_fma(v, qinv, 6755399441055744.0) - 6755399441055744.0;
    return __builtin_fma(-qh, q, v);
  };
  for (int m = 0; m < M; ++m) {
    std::fill(arow.begin(), arow.end(), 0);
    for (int i = 0; i < n_src; ++i)
      for (int a = 0; a < 5; ++a) arow[umma_k_index(i, a)] = (uint8_t)(y[(size_t)i * M + m] >> (8 * a));
    for (int n = 0; n < im.NP; ++n) {
      int64_t s = 0;
      for (int k = 0; k < im.K; ++k) s += (int64_t)arow[k] * img[(size_t)(k >> 4) * im.NP * 16 + (size_t)n * 16 + (k & 15)];
      if (s > 0x7FFFFFFFll) return 2;
      T[n] = (int32_t)s;
    }
    double rf = 0.0;
    if (im.fold) {
      rf = reduce(value(n_dst), (double)fold_q);
      if (rf < 0.0) rf += (double)fold_q;
    }
    for (int t = 0; t < n_dst; ++t) {
      const double v = value(t) + rf;
      if (!(v < 9007199254740992.0)) return 3;
      double r = reduce(v, (double)dst_q[t]);
      if (r < 0.0) r += (double)dst_q[t];
      out[(size_t)t * M + m] = (uint64_t)r;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle (layout_type 0), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();  // a tensor-core launch that never completes must not hang the device
  } while (!ok);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the loaded registers are operands of the wait, so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[5][4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0][0]), "+r"(v[0][1]), "+r"(v[0][2]), "+r"(v[0][3]), "+r"(v[1][0]), "+r"(v[1][1]), "+r"(v[1][2]), "+r"(v[1][3]),
                 "+r"(v[2][0]), "+r"(v[2][1]), "+r"(v[2][2]), "+r"(v[2][3]), "+r"(v[3][0]), "+r"(v[3][1]), "+r"(v[3][2]), "+r"(v[3][3]),
                 "+r"(v[4][0]), "+r"(v[4][1]), "+r"(v[4][2]), "+r"(v[4][3])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&f)[5]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(f[0]), "+r"(f[1]), "+r"(f[2]), "+r"(f[3]), "+r"(f[4]) : : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// sum_b 2^(8b) T_b as an exact double (< 2^52): the 64-bit integer is assembled directly inside the mantissa of 2^52
__device__ __forceinline__ double level_sum(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4) {
  u64 acc = ((u64)(0x43300000u + t4) << 32) | t0;
  acc += (u64)t1 * 256u;
  acc += (u64)t2 * 65536u;
  acc += (u64)t3 * 16777216u;
  return __longlong_as_double((long long)acc) - 4503599627370496.0;
}

constexpr int UMMA_THREADS = 512;
constexpr int UMMA_TM = 128;

template <int N16>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
k_bconv_umma(const ModConst *__restrict__ mc, LimbMap src_lm, LimbMap dst_lm, BConvArgs a, const uint8_t *__restrict__ img, int K, int NP,
             int ND, int has_fold, int tiles_per_batch, int n_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (uint32_t)UMMA_TM * K;
  unsigned char *As = smem;                                  // two buffers
  unsigned char *Bs = smem + 2 * a_bytes;                    // NP * K
  double2 *tq = reinterpret_cast<double2 *>(Bs + (size_t)NP * K);   // [ND] (q, 1/q) per target
  uint64_t *bars = reinterpret_cast<uint64_t *>(tq + ND);    // two mbarriers
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);

  // ---- constant set-up (before the programmatic dependency is resolved)
  for (uint32_t e = tid; e < 2 * a_bytes / 16; e += UMMA_THREADS) reinterpret_cast<uint4 *>(As)[e] = make_uint4(0, 0, 0, 0);
  for (uint32_t e = tid; e < (uint32_t)NP * K / 16; e += UMMA_THREADS) reinterpret_cast<uint4 *>(Bs)[e] = __ldg(reinterpret_cast<const uint4 *>(img) + e);
  const int ntv = a.n_dst + has_fold;
  for (int t = tid; t < ND; t += UMMA_THREADS) {
    double2 c = make_double2(1.0, 1.0);
    if (t < ntv) {
      const ModConst m = mc[t < a.n_dst ? dst_lm.mod[t] : a.fold_mod];
      c = make_double2(m.q, m.qinv);
    }
    tq[t] = c;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = 0x20u | ((uint32_t)(NP >> 3) << 17) | (8u << 24);   // D = s32, A = B = u8, K-major, N = NP, M = 128
  const uint32_t bar0 = smem_addr(&bars[0]);

  // loader role: coefficient row lr, source group g (sources 16 s + 4 g .. + 3 of every 16-source slab s)
  const int g = tid & 3, lr = tid >> 2;
  // epilogue role: TMEM lanes 32 (warp % 4) .. + 31 = coefficient rows, target blocks sub, sub + 4, ...
  const int row = ((warp & 3) << 5) | lane, sub = warp >> 2;
  const uint32_t t_lane = (uint32_t)((warp & 3) << 5) << 16;

  u64 y[N16][4];
  auto load_tile = [&](int tile) {
    const int batch = tile / tiles_per_batch;
    const u64 *in = a.in + (size_t)batch * a.in_batch_stride + (size_t)(tile - batch * tiles_per_batch) * UMMA_TM + lr;
#pragma unroll
    for (int s = 0; s < N16; ++s)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = s * 16 + g * 4 + k;
        y[s][k] = i < a.n_src ? __ldg(in + (size_t)src_lm.pos[i] * a.N) : 0ull;
      }
    if (a.step1) {  // uniform: per-source scaling inside the conversion (primitive entry point)
#pragma unroll
      for (int s = 0; s < N16; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = s * 16 + g * 4 + k;
          if (i < a.n_src) {
            const ModConst m = mc[src_lm.mod[i]];
            const double2 sc = a.step1[i];
            y[s][k] = f64_to_canonical(mulmod_const(u64_to_f64(y[s][k]), sc.x, sc.y, m.q), m.qi);
          }
        }
    }
  };
  auto pack_tile = [&](unsigned char *Ab) {
#pragma unroll
    for (int s = 0; s < N16; ++s) {
      const uint32_t l0 = (uint32_t)y[s][0], l1 = (uint32_t)y[s][1], l2 = (uint32_t)y[s][2], l3 = (uint32_t)y[s][3];
      const uint32_t h0 = (uint32_t)(y[s][0] >> 32), h1 = (uint32_t)(y[s][1] >> 32), h2 = (uint32_t)(y[s][2] >> 32), h3 = (uint32_t)(y[s][3] >> 32);
      uint32_t *dst = reinterpret_cast<uint32_t *>(Ab + ((size_t)(s * 5) * UMMA_TM + lr) * 16 + 4 * g);
      dst[0 * UMMA_TM * 4] = __byte_perm(__byte_perm(l0, l1, 0x0040), __byte_perm(l2, l3, 0x0040), 0x5410);
      dst[1 * UMMA_TM * 4] = __byte_perm(__byte_perm(l0, l1, 0x0051), __byte_perm(l2, l3, 0x0051), 0x5410);
      dst[2 * UMMA_TM * 4] = __byte_perm(__byte_perm(l0, l1, 0x0062), __byte_perm(l2, l3, 0x0062), 0x5410);
      dst[3 * UMMA_TM * 4] = __byte_perm(__byte_perm(l0, l1, 0x0073), __byte_perm(l2, l3, 0x0073), 0x5410);
      dst[4 * UMMA_TM * 4] = __byte_perm(__byte_perm(h0, h1, 0x0040), __byte_perm(h2, h3, 0x0040), 0x5410);
    }
  };
  auto issue_mma = [&](int buf) {  // one thread
    const uint32_t a_addr = smem_addr(As + (size_t)buf * a_bytes), b_addr = smem_addr(Bs);
    const uint32_t a_lbo = UMMA_TM * 16, b_lbo = (uint32_t)NP * 16;
    for (int ks = 0; ks < (K >> 5); ++ks)
      umma_i8(tmem_base + (uint32_t)buf * 256u, umma_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128), umma_desc(b_addr + ks * 2 * b_lbo, b_lbo, 128), idesc,
              ks > 0);
    umma_commit(bar0 + 8u * buf);
  };
  auto epilogue = [&](int tile, int buf) {
    const int batch = tile / tiles_per_batch;
    u64 *out = a.out + (size_t)batch * a.out_batch_stride + (size_t)(tile - batch * tiles_per_batch) * UMMA_TM + row;
    const uint32_t tb = tmem_base + t_lane + (uint32_t)buf * 256u;
    double rf = 0.0;
    if (has_fold) {  // uniform
      uint32_t f[5];
#pragma unroll
      for (int b = 0; b < 5; ++b) f[b] = tmem_ld1(tb + b * ND + a.n_dst);
      tmem_ld_wait(f);
      const double2 c = tq[a.n_dst];
      rf = canonicalize(reduce_signed(level_sum(f[0], f[1], f[2], f[3], f[4]), c.x, c.y), c.x);
    }
    const int n_blk = (a.n_dst + 3) >> 2;
    uint32_t v[2][5][4];
    if (sub < n_blk) {
#pragma unroll
      for (int b = 0; b < 5; ++b) tmem_ld4(tb + b * ND + 4 * sub, v[0][b]);
    }
#pragma unroll 1
    for (int blk = sub; blk < n_blk; blk += 8) {  // two blocks per trip: the second block's accumulators load while the first is reduced
      tmem_ld_wait(v[0]);
      const bool more1 = blk + 4 < n_blk;
      if (more1) {
#pragma unroll
        for (int b = 0; b < 5; ++b) tmem_ld4(tb + b * ND + 4 * (blk + 4), v[1][b]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int bl = blk + 4 * h;
        if (h == 1) {
          if (!more1) break;
          tmem_ld_wait(v[1]);
          if (bl + 4 < n_blk) {
#pragma unroll
            for (int b = 0; b < 5; ++b) tmem_ld4(tb + b * ND + 4 * (bl + 4), v[0][b]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = 4 * bl + j;
          if (t < a.n_dst) {
            const double2 c = tq[t];
            const double r = reduce_signed(level_sum(v[h][0][j], v[h][1][j], v[h][2][j], v[h][3][j], v[h][4][j]) + rf, c.x, c.y);
            u64 *o = out + (size_t)dst_lm.pos[t] * a.N;
            if (a.out_f64) *reinterpret_cast<double *>(o) = r;
            else *o = f64_to_canonical(r, (u64)c.x);
          }
        }
      }
    }
    tmem_ld_wait();
  };

  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  pdl_wait();  // the sources are another kernel's output; the outputs may still be read by one
  if (n_my > 0) {
    load_tile(blockIdx.x);
    pack_tile(As);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_mma(0);
    }
    if (n_my > 1) load_tile(blockIdx.x + gridDim.x);
    for (int it = 0; it < n_my; ++it) {
      const int buf = it & 1;
      const bool next = it + 1 < n_my;
      if (next) {
        pack_tile(As + (size_t)(buf ^ 1) * a_bytes);
        fence_async_smem();
      }
      tc_fence_before();
      __syncthreads();
      if (next) {
        if (tid == 0) {
          tc_fence_after();
          issue_mma(buf ^ 1);
        }
        if (it + 2 < n_my) load_tile(blockIdx.x + (it + 2) * gridDim.x);
      }
      mbar_wait(bar0 + 8u * buf, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      epilogue(blockIdx.x + it * gridDim.x, buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int N16>
static void launch_umma_t(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a, const BConvImage &im, cudaStream_t s) {
  const int tiles_per_batch = a.N / UMMA_TM, n_tiles = tiles_per_batch * a.n_batches;
  // at least half of the SM's shared memory: ONE CTA per SM (a second one would only sit in tcgen05.alloc until the first ends)
  const size_t smem = std::max<size_t>((size_t)2 * UMMA_TM * im.K + (size_t)im.NP * im.K + (size_t)im.ND * sizeof(double2) + 32, 116 * 1024);
  static PerDeviceOnce once;
  static int n_sm[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (once.first()) {
    cudaFuncSetAttribute(k_bconv_umma<N16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * UMMA_TM * 256 + 256 * 256 + 64 * 16 + 32);  // 132 KB
    cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = std::min(n_tiles, std::max(1, n_sm[dev & 63]));
  launch_pdl(k_bconv_umma<N16>, dim3(grid), dim3(UMMA_THREADS), smem, s, mc, src_lm, dst_lm, a, im.img, im.K, im.NP, im.ND, im.fold, tiles_per_batch, n_tiles);
}

void launch_bconv_umma(const ModConst *mc, const LimbMap &src_lm, const LimbMap &dst_lm, const BConvArgs &a, const BConvImage &im, cudaStream_t s) {
  switch (im.n16) {
    case 1: launch_umma_t<1>(mc, src_lm, dst_lm, a, im, s); break;
    case 2: launch_umma_t<2>(mc, src_lm, dst_lm, a, im, s); break;
    default: launch_umma_t<3>(mc, src_lm, dst_lm, a, im, s); break;
  }
}

}  // namespace hml
