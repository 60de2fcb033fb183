// See ntt.cuh for the algorithm.  sm_100a only.
#include "ntt.cuh"

namespace hml {

// ---- asynchronous bulk copy (TMA engine, no tensor map) global -> shared, completion on an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// A "round" = S consecutive radix-2 stages (S <= 3) of a length-2^LOGR sub-NTT, starting at stage ST,
// executed on 8 register-resident points per thread.  Thread unit u in [0, 2^LOGR / 8) owns 8/2^S
// butterfly groups of 2^S points each; register r = (group gi, point kk).
// ------------------------------------------------------------------------------------------------
template <int LOGR, int ST, int S>
struct Round {
  static constexpr int TL = (1 << LOGR) >> (ST + S);  // smallest butterfly distance in the round
  static constexpr int G = 8 >> S;                    // groups per thread
  __device__ static __forceinline__ int point(int u, int r) {
    const int gi = r >> S, kk = r & ((1 << S) - 1);
    const int g = u * G + gi;
    const int hi = g / TL, lo = g % TL;
    return hi * (TL << S) + kk * TL + lo;
  }
};

// twiddle index of the butterfly whose lower point is p, at stage i of a sub-NTT whose twiddle block
// starts at tw_base (1 for the column pass; R1 + row for the row pass): (tw_base << i) + p / (2t)
// One stage of a round: JS = position in the 3-stage register pattern (distance 4 >> JS).
template <int LOGR, int ST, int S, int JS, bool INV, bool TW_SMEM>
__device__ __forceinline__ void round_stage(double (&a)[8], int u, const double2 *__restrict__ tw, unsigned tw_base, double q) {
  using RD = Round<LOGR, ST, S>;
  constexpr int h = 4 >> JS, i = ST + JS - (3 - S);
#pragma unroll
  for (int sg = 0; sg < (1 << JS); ++sg) {
    const int r0 = sg << (3 - JS);
    const int p0 = RD::point(u, r0);
    const unsigned ti = (tw_base << i) + (p0 >> (LOGR - i));
    const double2 w = TW_SMEM ? tw[ti] : __ldg(&tw[ti]);
#pragma unroll
    for (int o = 0; o < h; ++o) {
      if constexpr (INV) gs_butterfly(a[r0 + o], a[r0 + o + h], w.x, w.y, q);
      else ct_butterfly(a[r0 + o], a[r0 + o + h], w.x, w.y, q);
    }
  }
}

template <int LOGR, int ST, int S, bool TW_SMEM = false>
__device__ __forceinline__ void ct_round(double (&a)[8], int u, const double2 *__restrict__ tw, unsigned tw_base, double q) {
  if constexpr (S >= 3) round_stage<LOGR, ST, S, 0, false, TW_SMEM>(a, u, tw, tw_base, q);
  if constexpr (S >= 2) round_stage<LOGR, ST, S, 1, false, TW_SMEM>(a, u, tw, tw_base, q);
  round_stage<LOGR, ST, S, 2, false, TW_SMEM>(a, u, tw, tw_base, q);
}

template <int LOGR, int ST, int S, bool TW_SMEM = false>
__device__ __forceinline__ void gs_round(double (&a)[8], int u, const double2 *__restrict__ tw, unsigned tw_base, double q) {
  round_stage<LOGR, ST, S, 2, true, TW_SMEM>(a, u, tw, tw_base, q);
  if constexpr (S >= 2) round_stage<LOGR, ST, S, 1, true, TW_SMEM>(a, u, tw, tw_base, q);
  if constexpr (S >= 3) round_stage<LOGR, ST, S, 0, true, TW_SMEM>(a, u, tw, tw_base, q);
}

// round split of a length-2^LOGR sub-NTT: S1 = 3, then S2, S3 (S3 may be 0)
template <int LOGR> struct Split {
  static constexpr int S1 = 3;
  static constexpr int S2 = (LOGR <= 6) ? (LOGR - 3) : (LOGR - 3 + 1) / 2;
  static constexpr int S3 = LOGR - 3 - S2;
};

struct LimbCtx {
  const u64 *in;
  u64 *out;
  const double2 *tw;
  double q, qinv;
  u64 qi;
  int limb;
};

__device__ __forceinline__ LimbCtx limb_ctx(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, bool inverse,
                                            int limb, int poly, int batch = 0) {
  LimbCtx c;
  c.limb = limb;
  const int mi = lm.mod[c.limb];
  const ModConst mc = t.mc[mi];
  c.q = mc.q; c.qinv = mc.qinv; c.qi = mc.qi;
  c.tw = (inverse ? t.inv : t.fwd) + ((size_t)mi << logN);
  const long long slot = lm.pos[c.limb];
  c.in = l.in + (long long)batch * l.in_batch_stride + (long long)poly * l.in_poly_stride + slot * l.in_limb_stride;
  c.out = l.out + (long long)batch * l.out_batch_stride + (long long)poly * l.out_poly_stride + slot * l.out_limb_stride;
  return c;
}

// ================================================================================ forward, columns
template <int LOGR1>
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_fwd_cols(NttTables t, int logN, LimbMap lm, NttLaunch l) {
  constexpr int R1 = 1 << LOGR1, C = NTT_TILE / R1, R2 = 1 << NTT_ROW_LOG;
  using SP = Split<LOGR1>;
  extern __shared__ __align__(16) double sm[];  // [R1][C] tile, then the R1 twiddles of the column pass
  double2 *stw = reinterpret_cast<double2 *>(sm + NTT_TILE);
  __shared__ __align__(8) unsigned long long bar;
  const int limb = blockIdx.y % l.n_limbs, poly = (blockIdx.y / l.n_limbs) % l.n_polys, batch = blockIdx.y / (l.n_limbs * l.n_polys);
  if (poly == lm.skip[limb]) return;
  const LimbCtx lc = limb_ctx(t, logN, lm, l, false, limb, poly, batch);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {  // every column shares the same 2^LOGR1 twiddles: one 4 KB bulk copy
    mbar_expect_tx(&bar, R1 * 16u);
    bulk_g2s(stw, lc.tw, R1 * 16u, &bar);
  }
  const int c = threadIdx.x % C, u = threadIdx.x / C;
  const int col = blockIdx.x * C + c;
  double a[8];
  {
    using RD = Round<LOGR1, 0, SP::S1>;
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = u64_to_f64(__ldg(&lc.in[(size_t)RD::point(u, r) * R2 + col]));
    mbar_wait(&bar, 0);
    ct_round<LOGR1, 0, SP::S1, true>(a, u, stw, 1u, lc.q);
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[RD::point(u, r) * C + c] = a[r];
  }
  __syncthreads();
  double *outd = reinterpret_cast<double *>(lc.out);
  if constexpr (SP::S3 == 0) {
    using RD = Round<LOGR1, SP::S1, SP::S2>;
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = sm[RD::point(u, r) * C + c];
    ct_round<LOGR1, SP::S1, SP::S2, true>(a, u, stw, 1u, lc.q);
#pragma unroll
    for (int r = 0; r < 8; ++r) outd[(size_t)RD::point(u, r) * R2 + col] = a[r];
  } else {
    {
      using RD = Round<LOGR1, SP::S1, SP::S2>;
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = sm[RD::point(u, r) * C + c];
      ct_round<LOGR1, SP::S1, SP::S2, true>(a, u, stw, 1u, lc.q);
#pragma unroll
      for (int r = 0; r < 8; ++r) sm[RD::point(u, r) * C + c] = a[r];
    }
    __syncthreads();
    using RD = Round<LOGR1, SP::S1 + SP::S2, SP::S3>;
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = sm[RD::point(u, r) * C + c];
    ct_round<LOGR1, SP::S1 + SP::S2, SP::S3, true>(a, u, stw, 1u, lc.q);
#pragma unroll
    for (int r = 0; r < 8; ++r) outd[(size_t)RD::point(u, r) * R2 + col] = a[r];
  }
}

// ================================================================================ row passes: one warp per row
// The last 8 forward stages (first 8 inverse stages) stay inside a contiguous 256-point row.  A warp owns a row:
// lane l holds points kk*32 + l (kk = register index), so the three widest stages (t = 128, 64, 32) are
// register-only and global accesses are 256-byte coalesced.  The five narrow stages (t = 16 .. 1) pair points in
// different lanes; instead of shared memory each stage swaps HALF of every lane's registers with lane ^ t
// (register bit 2 <-> lane bit b), after which each lane owns four complete butterflies (registers j, j+4).
// The swaps are never undone: the logical point of (lane, register) is tracked in closed form
//   stage b (= 4..0):  kk = (r & 3) | (lane bit 4) << 2,   lane-point bits: P_m = lane bit m-1 (m > b), P_b = r bit 2,
//                      P_m = lane bit m (m < b)
// and after the last stage registers (j, j+4) hold the adjacent points p, p+1 with
//   p = ((j + 4*(lane>>4)) * 32 + 2*(lane & 15)),  i.e. 16-byte stores that tile two 256-byte segments per warp.
// No shared memory, no block barrier.
constexpr int ROW_WARPS = 8;  // rows per CTA

__device__ __forceinline__ void swap_half(double (&a)[8], int b, bool upper) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double send = upper ? a[j] : a[j + 4];
    const double recv = __shfl_xor_sync(0xffffffffu, send, 1 << b);
    if (upper) a[j] = recv; else a[j + 4] = recv;
  }
}

// twiddle index of pair j at lane-stage b (NTT stage i = 7 - b): (tw_base << i) + kk * 2^(4-b) + ((lane & 15) >> b)
__device__ __forceinline__ unsigned lane_stage_tw(unsigned tw_base, int b, int j, int lane) {
  const int kk = j + ((lane >> 4) << 2);
  return (tw_base << (7 - b)) + (kk << (4 - b)) + ((lane & 15) >> b);
}

// The twiddles of the CTA's ROW_WARPS consecutive rows form ONE contiguous table segment per stage
// ((R1 + row0) << i .. (R1 + row0 + 8) << i), so eight bulk copies bring all of them (32 KB) into shared memory
// while the warps are still waiting for their data rows; every poly of the limb then reuses them.
// Three mbarriers, small segments first: stages 0-4 (4 KB) land almost immediately, stages 5-6 (12 KB) and
// stage 7 (16 KB) are only awaited right before the butterflies that need them.
constexpr int ROW_TW_ENTRIES = ROW_WARPS * 255;
__device__ __forceinline__ int row_tw_off(int i) { return ROW_WARPS * ((1 << i) - 1); }
__device__ __forceinline__ int row_tw_bar(int i) { return i <= 4 ? 0 : (i <= 6 ? 1 : 2); }

__device__ __forceinline__ void stage_row_twiddles(double2 *stw, unsigned long long *bar, const double2 *tw, unsigned tw_base0) {
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (ROW_WARPS * 16u) * 31);
    mbar_expect_tx(bar + 1, (ROW_WARPS * 16u) * 96);
    mbar_expect_tx(bar + 2, (ROW_WARPS * 16u) * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      bulk_g2s(stw + row_tw_off(i), tw + ((size_t)tw_base0 << i), (ROW_WARPS * 16u) << i, bar + row_tw_bar(i));
  }
}

// index inside the shared twiddle block: row-local w, lane-stage b (NTT stage i = 7 - b), butterfly pair j
__device__ __forceinline__ int lane_stage_stw(int w, int b, int j, int lane) {
  const int i = 7 - b, kk = j + ((lane >> 4) << 2);
  return row_tw_off(i) + (w << i) + (kk << (4 - b)) + ((lane & 15) >> b);
}

// Polys that share the limb's modulus (the beta digits of ModUp, the two key-switch accumulators, the two
// rescaled polys) are processed back to back by the same warp with the same shared twiddles.
// PIPE: software-pipelined item loop (next item's row prefetched into registers; 3 CTAs/SM) for launches with
// several items per limb; the plain variant keeps 4 CTAs/SM for single-item launches.
template <bool PIPE>
__global__ void __launch_bounds__(ROW_WARPS * 32, PIPE ? 3 : 4) ntt_fwd_rows(NttTables t, int logN, LimbMap lm, NttLaunch l) {
  constexpr int LR = NTT_ROW_LOG, R2 = 1 << LR;
  __shared__ __align__(16) double2 stw[ROW_TW_ENTRIES];
  __shared__ __align__(8) unsigned long long bar[3];
  const int limb = blockIdx.y;
  const unsigned R1 = 1u << (logN - LR);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, row = blockIdx.x * ROW_WARPS + w;
  stage_row_twiddles(stw, bar, t.fwd + ((size_t)lm.mod[limb] << logN), R1 + blockIdx.x * ROW_WARPS);
  bool ready = false;
  // items = (ciphertext b, poly p) pairs sharing this limb's modulus; the next item's row is in flight (registers)
  // while the current one is transformed
  const int skip = lm.skip[limb], n_items = l.n_polys * l.n_batch;
  auto next_item = [&](int it) { ++it; while (it < n_items && (it % l.n_polys) == skip) ++it; return it; };
  int item = next_item(-1);
  double nx[PIPE ? 8 : 1];
  if (PIPE && item < n_items) {
    const LimbCtx lc0 = limb_ctx(t, logN, lm, l, false, limb, item % l.n_polys, item / l.n_polys);
    const double *ind = reinterpret_cast<const double *>(lc0.out) + (size_t)row * R2;  // pass 1 left raw doubles in `out`
#pragma unroll
    for (int r = 0; r < 8; ++r) nx[PIPE ? r : 0] = ind[r * 32 + lane];
  }
  while (item < n_items) {
    const LimbCtx lc = limb_ctx(t, logN, lm, l, false, limb, item % l.n_polys, item / l.n_polys);
    double a[8];
    if (PIPE) {
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = nx[PIPE ? r : 0];
    } else {
      const double *ind = reinterpret_cast<const double *>(lc.out) + (size_t)row * R2;
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = ind[r * 32 + lane];
    }
    item = next_item(item);
    if (PIPE && item < n_items) {
      const LimbCtx ln = limb_ctx(t, logN, lm, l, false, limb, item % l.n_polys, item / l.n_polys);
      const double *ind = reinterpret_cast<const double *>(ln.out) + (size_t)row * R2;
#pragma unroll
      for (int r = 0; r < 8; ++r) nx[PIPE ? r : 0] = ind[r * 32 + lane];
    }
    if (!ready) mbar_wait(bar, 0);
    {  // stages t = 128, 64, 32: register-only, twiddles broadcast from shared memory
      const double2 w0 = stw[row_tw_off(0) + w];
#pragma unroll
      for (int o = 0; o < 4; ++o) ct_butterfly(a[o], a[o + 4], w0.x, w0.y, lc.q);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const double2 w1 = stw[row_tw_off(1) + (w << 1) + g];
#pragma unroll
        for (int o = 0; o < 2; ++o) ct_butterfly(a[4 * g + o], a[4 * g + o + 2], w1.x, w1.y, lc.q);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const double2 w2 = stw[row_tw_off(2) + (w << 2) + g];
        ct_butterfly(a[2 * g], a[2 * g + 1], w2.x, w2.y, lc.q);
      }
    }
#pragma unroll
    for (int b = 4; b >= 0; --b) {
      if (!ready && b == 2) mbar_wait(bar + 1, 0);  // stages 5, 6
      if (!ready && b == 0) mbar_wait(bar + 2, 0);  // stage 7
      double2 tw4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tw4[j] = stw[lane_stage_stw(w, b, j, lane)];
      swap_half(a, b, (lane >> b) & 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) ct_butterfly(a[j], a[j + 4], tw4[j].x, tw4[j].y, lc.q);
    }
    ready = true;
    u64 *outp = lc.out + (size_t)row * R2 + (lane >> 4) * 128 + 2 * (lane & 15);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u64 v0 = f64_to_canonical(reduce_signed(a[j], lc.q, lc.qinv), lc.qi);
      const u64 v1 = f64_to_canonical(reduce_signed(a[j + 4], lc.q, lc.qinv), lc.qi);
      *reinterpret_cast<ulonglong2 *>(outp + j * 32) = make_ulonglong2(v0, v1);
    }
  }
  if (!ready) { mbar_wait(bar, 0); mbar_wait(bar + 1, 0); mbar_wait(bar + 2, 0); }  // never exit with a bulk copy in flight
}

template <bool PIPE>
__global__ void __launch_bounds__(ROW_WARPS * 32, PIPE ? 3 : 4) ntt_inv_rows(NttTables t, int logN, LimbMap lm, NttLaunch l) {
  constexpr int LR = NTT_ROW_LOG, R2 = 1 << LR;
  __shared__ __align__(16) double2 stw[ROW_TW_ENTRIES];
  __shared__ __align__(8) unsigned long long bar[3];
  const int limb = blockIdx.y;
  const unsigned R1 = 1u << (logN - LR);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, row = blockIdx.x * ROW_WARPS + w;
  stage_row_twiddles(stw, bar, t.inv + ((size_t)lm.mod[limb] << logN), R1 + blockIdx.x * ROW_WARPS);
  bool ready = false;
  const int skip = lm.skip[limb], n_items = l.n_polys * l.n_batch;
  auto next_item = [&](int it) { ++it; while (it < n_items && (it % l.n_polys) == skip) ++it; return it; };
  int item = next_item(-1);
  ulonglong2 nx[PIPE ? 4 : 1];
  if (PIPE && item < n_items) {
    const LimbCtx lc0 = limb_ctx(t, logN, lm, l, true, limb, item % l.n_polys, item / l.n_polys);
    const u64 *inp = lc0.in + (size_t)row * R2 + (lane >> 4) * 128 + 2 * (lane & 15);
#pragma unroll
    for (int j = 0; j < 4; ++j) nx[PIPE ? j : 0] = __ldg(reinterpret_cast<const ulonglong2 *>(inp + j * 32));
  }
  while (item < n_items) {
    const LimbCtx lc = limb_ctx(t, logN, lm, l, true, limb, item % l.n_polys, item / l.n_polys);
    double a[8];
    if (PIPE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = u64_to_f64(nx[PIPE ? j : 0].x);
        a[j + 4] = u64_to_f64(nx[PIPE ? j : 0].y);
      }
    } else {
      const u64 *inp = lc.in + (size_t)row * R2 + (lane >> 4) * 128 + 2 * (lane & 15);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(inp + j * 32));
        a[j] = u64_to_f64(v.x);
        a[j + 4] = u64_to_f64(v.y);
      }
    }
    item = next_item(item);
    if (PIPE && item < n_items) {
      const LimbCtx ln = limb_ctx(t, logN, lm, l, true, limb, item % l.n_polys, item / l.n_polys);
      const u64 *inp = ln.in + (size_t)row * R2 + (lane >> 4) * 128 + 2 * (lane & 15);
#pragma unroll
      for (int j = 0; j < 4; ++j) nx[PIPE ? j : 0] = __ldg(reinterpret_cast<const ulonglong2 *>(inp + j * 32));
    }
    if (!ready) { mbar_wait(bar + 2, 0); mbar_wait(bar + 1, 0); mbar_wait(bar, 0); ready = true; }  // the inverse starts with stage 7
#pragma unroll
    for (int b = 0; b <= 4; ++b) {
      double2 tw4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tw4[j] = stw[lane_stage_stw(w, b, j, lane)];
#pragma unroll
      for (int j = 0; j < 4; ++j) gs_butterfly(a[j], a[j + 4], tw4[j].x, tw4[j].y, lc.q);
      swap_half(a, b, (lane >> b) & 1);
    }
    {  // stages t = 32, 64, 128
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const double2 w2 = stw[row_tw_off(2) + (w << 2) + g];
        gs_butterfly(a[2 * g], a[2 * g + 1], w2.x, w2.y, lc.q);
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const double2 w1 = stw[row_tw_off(1) + (w << 1) + g];
#pragma unroll
        for (int o = 0; o < 2; ++o) gs_butterfly(a[4 * g + o], a[4 * g + o + 2], w1.x, w1.y, lc.q);
      }
      const double2 w0 = stw[row_tw_off(0) + w];
#pragma unroll
      for (int o = 0; o < 4; ++o) gs_butterfly(a[o], a[o + 4], w0.x, w0.y, lc.q);
    }
    double *outd = reinterpret_cast<double *>(lc.out) + (size_t)row * R2;
    // the sums have grown to <= 2^8 q: bring them back to |v| <= q/2 before the column pass doubles them again
#pragma unroll
    for (int r = 0; r < 8; ++r) outd[r * 32 + lane] = reduce_signed(a[r], lc.q, lc.qinv);
  }
  if (!ready) { mbar_wait(bar, 0); mbar_wait(bar + 1, 0); mbar_wait(bar + 2, 0); }
}

// ================================================================================ inverse, columns (second)
template <int LOGR1>
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_inv_cols(NttTables t, int logN, LimbMap lm, NttLaunch l) {
  constexpr int R1 = 1 << LOGR1, C = NTT_TILE / R1, R2 = 1 << NTT_ROW_LOG;
  using SP = Split<LOGR1>;
  extern __shared__ __align__(16) double sm[];  // [R1][C] tile, then the R1 twiddles of the column pass
  double2 *stw = reinterpret_cast<double2 *>(sm + NTT_TILE);
  __shared__ __align__(8) unsigned long long bar;
  const int limb = blockIdx.y % l.n_limbs, poly = (blockIdx.y / l.n_limbs) % l.n_polys, batch = blockIdx.y / (l.n_limbs * l.n_polys);
  if (poly == lm.skip[limb]) return;
  const LimbCtx lc = limb_ctx(t, logN, lm, l, true, limb, poly, batch);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {  // every column shares the same 2^LOGR1 twiddles: one 4 KB bulk copy
    mbar_expect_tx(&bar, R1 * 16u);
    bulk_g2s(stw, lc.tw, R1 * 16u, &bar);
  }
  const int c = threadIdx.x % C, u = threadIdx.x / C;
  const int col = blockIdx.x * C + c;
  const double *ind = reinterpret_cast<const double *>(lc.out);
  double a[8];
  if constexpr (SP::S3 != 0) {
    {
      using RD = Round<LOGR1, SP::S1 + SP::S2, SP::S3>;
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = ind[(size_t)RD::point(u, r) * R2 + col];
      mbar_wait(&bar, 0);
      gs_round<LOGR1, SP::S1 + SP::S2, SP::S3, true>(a, u, stw, 1u, lc.q);
#pragma unroll
      for (int r = 0; r < 8; ++r) sm[RD::point(u, r) * C + c] = a[r];
    }
    __syncthreads();
    using RD = Round<LOGR1, SP::S1, SP::S2>;
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = sm[RD::point(u, r) * C + c];
    gs_round<LOGR1, SP::S1, SP::S2, true>(a, u, stw, 1u, lc.q);
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[RD::point(u, r) * C + c] = a[r];
  } else {
    using RD = Round<LOGR1, SP::S1, SP::S2>;
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = ind[(size_t)RD::point(u, r) * R2 + col];
    mbar_wait(&bar, 0);
    gs_round<LOGR1, SP::S1, SP::S2, true>(a, u, stw, 1u, lc.q);
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[RD::point(u, r) * C + c] = a[r];
  }
  __syncthreads();
  using RD = Round<LOGR1, 0, SP::S1>;
#pragma unroll
  for (int r = 0; r < 8; ++r) a[r] = sm[RD::point(u, r) * C + c];
  gs_round<LOGR1, 0, SP::S1, true>(a, u, stw, 1u, lc.q);
  double2 sc;
  if (l.post_scale) sc = l.post_scale[lc.limb];
  else { const ModConst mc = t.mc[lm.mod[lc.limb]]; sc = make_double2(mc.ninv, mc.ninv_q); }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    lc.out[(size_t)RD::point(u, r) * R2 + col] = f64_to_canonical(mulmod_const(a[r], sc.x, sc.y, lc.q), lc.qi);
}

// ================================================================================ small N (<= 4096): one CTA per limb
__global__ void __launch_bounds__(256) ntt_small(NttTables t, int logN, LimbMap lm, NttLaunch l, int inverse) {
  extern __shared__ double sm[];
  const int limb = blockIdx.y % l.n_limbs, poly = (blockIdx.y / l.n_limbs) % l.n_polys, batch = blockIdx.y / (l.n_limbs * l.n_polys);
  if (poly == lm.skip[limb]) return;
  const LimbCtx lc = limb_ctx(t, logN, lm, l, inverse != 0, limb, poly, batch);
  const int N = 1 << logN, half = N >> 1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sm[i] = u64_to_f64(__ldg(&lc.in[i]));
  __syncthreads();
  if (!inverse) {
    for (int s = 0; s < logN; ++s) {
      const int tt = N >> (s + 1);
      for (int b = threadIdx.x; b < half; b += blockDim.x) {
        const int grp = b / tt, j = b % tt, i0 = grp * 2 * tt + j;
        const double2 w = __ldg(&lc.tw[(1 << s) + grp]);
        double x = sm[i0], y = sm[i0 + tt];
        ct_butterfly(x, y, w.x, w.y, lc.q);
        sm[i0] = x; sm[i0 + tt] = y;
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x)
      lc.out[i] = f64_to_canonical(reduce_signed(sm[i], lc.q, lc.qinv), lc.qi);
  } else {
    for (int s = logN - 1; s >= 0; --s) {
      const int tt = N >> (s + 1);
      for (int b = threadIdx.x; b < half; b += blockDim.x) {
        const int grp = b / tt, j = b % tt, i0 = grp * 2 * tt + j;
        const double2 w = __ldg(&lc.tw[(1 << s) + grp]);
        double x = sm[i0], y = sm[i0 + tt];
        gs_butterfly(x, y, w.x, w.y, lc.q);
        if (((logN - s) & 3) == 0) { x = reduce_signed(x, lc.q, lc.qinv); }  // bound the doubling every 4 stages
        sm[i0] = x; sm[i0 + tt] = y;
      }
      __syncthreads();
    }
    double2 sc;
    if (l.post_scale) sc = l.post_scale[lc.limb];
    else { const ModConst mc = t.mc[lm.mod[lc.limb]]; sc = make_double2(mc.ninv, mc.ninv_q); }
    for (int i = threadIdx.x; i < N; i += blockDim.x)
      lc.out[i] = f64_to_canonical(mulmod_const(sm[i], sc.x, sc.y, lc.q), lc.qi);
  }
}

// ================================================================================ host launchers
template <int LOGR1>
static void launch_fwd_t(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s) {
  constexpr int C = NTT_TILE >> LOGR1;
  const dim3 g1((1 << NTT_ROW_LOG) / C, l.n_limbs * l.n_polys * l.n_batch), g2((1 << LOGR1) / ROW_WARPS, l.n_limbs);
  ntt_fwd_cols<LOGR1><<<g1, NTT_THREADS, NTT_TILE * sizeof(double) + (16u << LOGR1), s>>>(t, logN, lm, l);
  if (l.n_polys * l.n_batch >= 3) ntt_fwd_rows<true><<<g2, ROW_WARPS * 32, 0, s>>>(t, logN, lm, l);
  else ntt_fwd_rows<false><<<g2, ROW_WARPS * 32, 0, s>>>(t, logN, lm, l);
}
template <int LOGR1>
static void launch_inv_t(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s) {
  constexpr int C = NTT_TILE >> LOGR1;
  const dim3 g1((1 << LOGR1) / ROW_WARPS, l.n_limbs), g2((1 << NTT_ROW_LOG) / C, l.n_limbs * l.n_polys * l.n_batch);
  if (l.n_polys * l.n_batch >= 3) ntt_inv_rows<true><<<g1, ROW_WARPS * 32, 0, s>>>(t, logN, lm, l);
  else ntt_inv_rows<false><<<g1, ROW_WARPS * 32, 0, s>>>(t, logN, lm, l);
  ntt_inv_cols<LOGR1><<<g2, NTT_THREADS, NTT_TILE * sizeof(double) + (16u << LOGR1), s>>>(t, logN, lm, l);
}

void launch_ntt_forward(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s) {
  if (logN <= 12) {
    ntt_small<<<dim3(1, l.n_limbs * l.n_polys * l.n_batch), 256, sizeof(double) << logN, s>>>(t, logN, lm, l, 0);
    return;
  }
  switch (logN - NTT_ROW_LOG) {
    case 5: launch_fwd_t<5>(t, logN, lm, l, s); break;
    case 6: launch_fwd_t<6>(t, logN, lm, l, s); break;
    case 7: launch_fwd_t<7>(t, logN, lm, l, s); break;
    case 8: launch_fwd_t<8>(t, logN, lm, l, s); break;
  }
}

void launch_ntt_inverse(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s) {
  if (logN <= 12) {
    ntt_small<<<dim3(1, l.n_limbs * l.n_polys * l.n_batch), 256, sizeof(double) << logN, s>>>(t, logN, lm, l, 1);
    return;
  }
  switch (logN - NTT_ROW_LOG) {
    case 5: launch_inv_t<5>(t, logN, lm, l, s); break;
    case 6: launch_inv_t<6>(t, logN, lm, l, s); break;
    case 7: launch_inv_t<7>(t, logN, lm, l, s); break;
    case 8: launch_inv_t<8>(t, logN, lm, l, s); break;
  }
}

}  // namespace hml
