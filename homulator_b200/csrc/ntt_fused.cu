// Single-launch negacyclic NTT / inverse NTT for two-pass rings (N >= 8192): both passes in ONE persistent kernel.
//
// Reference instruction classes NTT / INTT (InsGen::GenNTT, reference src/InsGen.cpp:17-44); the simulated NTTU is itself
// "phase 1 -> transpose -> phase 2" inside one unit (reference src/Components.cpp:397-431), which is what this kernel
// restores on the GPU: the two-kernel version (ntt.cu) wrote the 8-byte intermediate of every limb to HBM between the passes
// (2 W extra per limb, 2.03x the algorithmic traffic in ncu) and paid two launch ramps and tails per transform.
//
// Structure.  The transform of one polynomial limb (a "member") is T = N / 4096 column tiles followed by T row tiles (forward;
// the inverse runs rows first), every tile 4096 points = 32 KB, the same arithmetic as ntt.cu (16 points per thread, two
// rounds of four radix-2 stages, FP64 datapath of modarith.cuh).  Members that share a modulus are grouped into super-jobs of
// up to FUSED_G; a work ITEM is (super-job, tile index, pass) and processes that tile of every member of the group, so the
// 32 KB row-twiddle blob of a (modulus, tile) is fetched once per item.  All items of a launch sit in ONE global queue,
//   [first-pass items of the first `lag` super-jobs] [first-pass items of super-job k + lag interleaved with second-pass
//   items of super-job k] ... [second-pass items of the last `lag` super-jobs]
// and the persistent CTAs (two per SM) pop it with an atomic counter.  A second-pass item waits (acquire-spin on a
// per-super-job counter, released by the first-pass items) until all T first-pass tiles of its super-job are done; because
// every item it can wait for was popped EARLIER by a CTA that is running and never waits itself, the queue cannot deadlock.
// With `lag` super-jobs (a few MB) between producer and consumer the intermediate is still in the 126 MB L2 when it is
// read back and is overwritten there by the final result: HBM sees one read and one write per limb.
// Inside a CTA the tiles ("units") are software-pipelined exactly like the two-kernel version: the next unit's cp.async
// loads are issued before the current unit is transformed.
#include "ntt_core.cuh"

#include <algorithm>
#include <cstdlib>

namespace hml {

constexpr int FUSED_G = 8;          // members per super-job
constexpr int FUSED_MAX_SJ = 8192;  // super-jobs per launch (counters: ctrl[2 + sj], ctrl[2 + FUSED_MAX_SJ + sj])
constexpr int FUSED_STAGE_BYTES = NTT_TILE * 8 + 2048 + 64;   // tile | column twiddles (<= 256) | ModConst (48 B) | post-scale (16 B)
constexpr int FUSED_SMEM_BYTES = ROW_TILE_BYTES + 2 * FUSED_STAGE_BYTES;

struct FusedPlan {
  int G;        // members per super-job (<= FUSED_G; small launches use 1 so that every tile is its own work item)
  int T;        // tiles per member and pass
  int n_sj;     // super-jobs
  int lag;      // super-jobs between the first pass and the second
  int n_items;  // 2 * n_sj * T
  uint16_t sj_first[NTT_MAX_LIMBS + 1];  // first super-job of limb l (prefix sums), [n_limbs] = n_sj
  uint16_t members[NTT_MAX_LIMBS];       // members of limb l (items (b, p) with p != skip)
};

struct ItemDesc {  // decoded by thread 0, broadcast through shared memory
  int valid, second, sj, t, limb, mi, m0, cnt, ready;
};

// queue position -> (pass, super-job, tile)
__device__ __forceinline__ void decode_pos(const FusedPlan &fp, int x, int &second, int &sj, int &t) {
  const int T = fp.T, a = fp.lag * T;
  if (x < a) { second = 0; sj = x / T; t = x - sj * T; return; }
  x -= a;
  const int nb = fp.n_sj - fp.lag, b = 2 * T * nb;
  if (x < b) {
    const int blk = x / (2 * T), r = x - blk * 2 * T;
    second = r & 1; t = r >> 1; sj = second ? blk : blk + fp.lag;
    return;
  }
  x -= b;
  second = 1; sj = nb + x / T; t = x % T;
}

// INV = false: forward (column pass first, then rows); INV = true: inverse (rows first, then columns)
template <int LOGR1, bool INV, bool IN_F64, bool FUSE>
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_fused(NttTables t, int logN, LimbMap lm, NttLaunch l, FusedPlan fp, unsigned *ctrl) {
  constexpr int NT = NTT_THREADS, R1 = 1 << LOGR1, TILE = NTT_TILE, C = TILE / R1, SKIP = 8 - LOGR1;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ ItemDesc next_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *blob = reinterpret_cast<const double *>(smem);
  unsigned char *stage0 = smem + ROW_TILE_BYTES;
  const unsigned stage0_u32 = smem_u32(stage0);
  if (tid == 0) mbar_init(&bar, 1);
  unsigned blob_phase = 0;
  int blob_mi = -1, blob_t = -1;
  __syncthreads();

  // ---- geometry of the two unit kinds
  const int cc = tid % C, cu = tid / C;                       // column unit: thread (column, row group)
  const RowAddr ad = row_addr(lane, warp);                    // row unit: warp-local addressing
  const int l16 = lane & 15, rr = 2 * warp + (lane >> 4);

  auto member_bp = [&](int limb, int m, int &b, int &p) {
    const int skip = lm.skip[limb];
    if (skip == 0xFF) { b = m / l.n_polys; p = m - b * l.n_polys; }
    else { const int np1 = l.n_polys - 1; b = m / np1; const int pp = m - b * np1; p = pp + (pp >= skip ? 1 : 0); }
  };
  auto base_in = [&](int limb, int b, int p) -> const u64 * {
    return l.in + (long long)b * l.in_batch_stride + (long long)p * l.in_poly_stride + (long long)lm.pos[limb] * l.in_limb_stride;
  };
  auto base_out = [&](int limb, int b, int p) -> u64 * {
    return l.out + (long long)b * l.out_batch_stride + (long long)p * l.out_poly_stride + (long long)lm.pos[limb] * l.out_limb_stride;
  };
  // is this item a column-type item?  forward: first pass = columns; inverse: second pass = columns
  auto is_col = [&](int second) { return INV ? second != 0 : second == 0; };

  // thread 0: pop + decode the next item into shared memory
  int popped = -1;  // thread 0: queue position fetched ahead of its use (the atomic's latency hides behind a unit's arithmetic)
  auto pop = [&]() {
    const int x = popped >= 0 ? popped : (int)atomicAdd(&ctrl[0], 1u);
    popped = -1;
    ItemDesc d{};
    if (x < fp.n_items) {
      d.valid = 1;
      decode_pos(fp, x, d.second, d.sj, d.t);
      int lo = 0, hi = l.n_limbs;  // largest limb with sj_first[limb] <= sj
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fp.sj_first[mid] <= d.sj) lo = mid; else hi = mid; }
      d.limb = lo;
      d.mi = lm.mod[lo];
      d.m0 = (d.sj - fp.sj_first[lo]) * fp.G;
      d.cnt = min(fp.G, (int)fp.members[lo] - d.m0);
      d.ready = 1;
      if (d.second) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctrl + 2 + d.sj) : "memory");
        d.ready = v >= (unsigned)fp.T;
      }
    }
    next_item = d;
  };
  auto wait_ready = [&](int sj) {  // all threads; thread 0 spins, the barrier publishes
    if (tid == 0) {
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctrl + 2 + sj) : "memory"); } while (v < (unsigned)fp.T);
    }
    __syncthreads();
  };

  // ---- loads of one unit into a stage
  auto issue_col = [&](const ItemDesc &it, int m, unsigned st) {
    int b, p;
    member_bp(it.limb, it.m0 + m, b, p);
    // forward: the first pass reads the caller's input; inverse: the second pass works in place on `out`
    const u64 *src = (INV ? (const u64 *)base_out(it.limb, b, p) : base_in(it.limb, b, p)) + it.t * C;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int qd = tid + NT * k, row = qd / (C / 2), c2 = qd % (C / 2);
      cp_async16(st + qd * 16, src + (size_t)row * (1 << NTT_ROW_LOG) + 2 * c2);
    }
    const double *tw_table = INV ? t.inv : t.fwd;
    for (int x = tid; x < R1 / 2 + 4; x += NT) {
      const int y = x - R1 / 2;
      if (y < 0) cp_async16(st + TILE * 8 + x * 16, tw_table + ((size_t)it.mi << logN) + 2 * x);
      else if (y < 3) cp_async16(st + TILE * 8 + 2048 + y * 16, reinterpret_cast<const char *>(t.mc + it.mi) + y * 16);
      else if (INV && l.post_scale) cp_async16(st + TILE * 8 + 2048 + 48, l.post_scale + it.limb);
    }
  };
  auto issue_row = [&](const ItemDesc &it, int m, unsigned st) {
    int b, p;
    member_bp(it.limb, it.m0 + m, b, p);
    // forward: rows read the raw doubles the column pass left in `out`; inverse: rows read the caller's input words
    const u64 *src = (INV ? base_in(it.limb, b, p) : (const u64 *)base_out(it.limb, b, p)) + (size_t)it.t * TILE;
    row_issue(src, st, lane, warp);
  };
  auto issue_blob = [&](const ItemDesc &it) {  // one thread, after a barrier that ends every generic-proxy read of the old blob
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar, ROW_TILE_BYTES);
    bulk_g2s(smem, (INV ? t.inv_rows : t.fwd_rows) + ((size_t)it.mi << logN) + (size_t)it.t * TILE, ROW_TILE_BYTES, &bar);
  };

  // ---- transforms of one unit
  auto compute_col_fwd = [&](const ItemDesc &it, int m, unsigned char *stg) {
    int b, p;
    member_bp(it.limb, it.m0 + m, b, p);
    double *data = reinterpret_cast<double *>(stg);
    const double *tw = data + TILE;
    const ModConst &mc = *reinterpret_cast<const ModConst *>(tw + 256);
    const double q = mc.q, qinv = mc.qinv;
    double a[16], w[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = IN_F64 ? data[tid + NT * j] : u64_to_f64(reinterpret_cast<const u64 *>(data)[tid + NT * j]);
    // a constant added to coefficient 0 shows up unchanged in every slot: -h = -(q-1)/2 makes the row pass's centred
    // reduction land in [-h, h] = [0, q-1] - h, so its canonicalisation is one add (see ntt.cu)
    if (it.t == 0 && tid == 0 && !FUSE && !l.out_f64) a[0] -= (q - 1.0) * 0.5;
    lds_run<1>(w, tw + 1); ct_level<0>(a, w, q, qinv);
    lds_run<2>(w, tw + 2); ct_level<1>(a, w, q, qinv);
    lds_run<4>(w, tw + 4); ct_level<2>(a, w, q, qinv);
    lds_run<8>(w, tw + 8); ct_level<3>(a, w, q, qinv);
#pragma unroll
    for (int j = 0; j < 16; ++j) data[tid + NT * j] = a[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = data[(16 * cu + j) * C + cc];
    if constexpr (SKIP <= 0) { lds_run<1>(w, tw + (1 << (LOGR1 - 4)) + cu); ct_level<0>(a, w, q, qinv); }
    if constexpr (SKIP <= 1) { lds_run<2>(w, tw + (1 << (LOGR1 - 3)) + 2 * cu); ct_level<1>(a, w, q, qinv); }
    if constexpr (SKIP <= 2) { lds_run<4>(w, tw + (1 << (LOGR1 - 2)) + 4 * cu); ct_level<2>(a, w, q, qinv); }
    lds_run<8>(w, tw + (1 << (LOGR1 - 1)) + 8 * cu); ct_level<3>(a, w, q, qinv);
    double *outd = reinterpret_cast<double *>(base_out(it.limb, b, p)) + it.t * C + cc;
#pragma unroll
    for (int j = 0; j < 16; ++j) outd[(size_t)(16 * cu + j) << NTT_ROW_LOG] = a[j];
  };
  auto compute_col_inv = [&](const ItemDesc &it, int m, unsigned char *stg) {
    int b, p;
    member_bp(it.limb, it.m0 + m, b, p);
    double *data = reinterpret_cast<double *>(stg);
    const double *tw = data + TILE;
    const ModConst &mc = *reinterpret_cast<const ModConst *>(tw + 256);
    const double q = mc.q, qinv = mc.qinv;
    double a[16], w[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = data[(16 * cu + j) * C + cc];
    lds_run<8>(w, tw + (1 << (LOGR1 - 1)) + 8 * cu); gs_level<3>(a, w, q, qinv);
    if constexpr (SKIP <= 2) { lds_run<4>(w, tw + (1 << (LOGR1 - 2)) + 4 * cu); gs_level<2>(a, w, q, qinv); }
    if constexpr (SKIP <= 1) { lds_run<2>(w, tw + (1 << (LOGR1 - 3)) + 2 * cu); gs_level<1>(a, w, q, qinv); }
    if constexpr (SKIP <= 0) { lds_run<1>(w, tw + (1 << (LOGR1 - 4)) + cu); gs_level<0>(a, w, q, qinv); }
#pragma unroll
    for (int j = 0; j < 16; ++j) data[(16 * cu + j) * C + cc] = a[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = data[tid + NT * j];
    lds_run<8>(w, tw + 8); gs_level<3>(a, w, q, qinv);
    lds_run<4>(w, tw + 4); gs_level<2>(a, w, q, qinv);
    lds_run<2>(w, tw + 2); gs_level<1>(a, w, q, qinv);
    // last level (one twiddle w1 for the whole limb) with the post-scale c folded in: (x + y) * c and (x - y) * (w1 * c)
    double2 sc;
    if (l.post_scale) sc = *reinterpret_cast<const double2 *>(tw + 256 + 6);
    else sc = make_double2(mc.ninv, mc.ninv_q);
    const double w1c = mulmod_var(tw[1], sc.x, q, qinv);
    const u64 qi = mc.qi;
    u64 *outp = base_out(it.limb, b, p) + it.t * C + cc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double sum = __dadd_rn(a[j], a[j + 8]), dif = __dsub_rn(a[j], a[j + 8]);
      outp[(size_t)(cu + (R1 / 16) * j) << NTT_ROW_LOG] = f64_to_canonical(mulmod_const(sum, sc.x, sc.y, q), qi);
      outp[(size_t)(cu + (R1 / 16) * (j + 8)) << NTT_ROW_LOG] = f64_to_canonical(mulmod_var(dif, w1c, q, qinv), qi);
    }
  };
  auto compute_row = [&](const ItemDesc &it, int m, unsigned char *data) {
    int b, p;
    member_bp(it.limb, it.m0 + m, b, p);
    const ModConst mc = t.mc[it.mi];
    const double q = mc.q, qinv = mc.qinv;
    const double hb = 4503599627370496.0 + (q - 1.0) * 0.5;  // 2^52 + h
    const double *tw_a = blob + rr * 16;
    const size_t tile_off = (size_t)it.t * TILE;
    const int limb = it.limb;
    double a[16], w[8];
    if constexpr (FUSE) {
      // the epilogue's operands (this warp's two rows = 512 coefficients of x and z) start their trip to L2 now
      const NttFuse &f = l.fuse;
      const long long fb = p / f.n_c, fc = p % f.n_c;
      const size_t nn = (size_t)1 << logN, c0 = tile_off + (size_t)warp * 512;
      auto prefetch = [&](const u64 *slot, int packed) {
        const unsigned char *bp = reinterpret_cast<const unsigned char *>(slot);
        const unsigned char *pp = nullptr;
        if (!packed) pp = bp + c0 * 8 + lane * 128;
        else if (lane < 16) pp = bp + c0 * 4 + lane * 128;
        else if (lane < 20) pp = bp + 4 * nn + c0 + (lane - 16) * 128;
        if (pp) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
      };
      prefetch(f.x + fc * f.x_c_stride + fb * f.x_b_stride + (size_t)limb * nn, f.x_packed);
      if (f.z != nullptr && ((f.z_mask >> fc) & 1u)) prefetch(f.z + fc * f.z_c_stride + fb * f.z_b_stride + (size_t)limb * nn, f.z_packed);
    }
    if constexpr (!INV) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = *reinterpret_cast<const double *>(data + ad.A(j));
      lds_run<1>(w, tw_a + 1); ct_level<0>(a, w, q, qinv);
      lds_run<2>(w, tw_a + 2); ct_level<1>(a, w, q, qinv);
      lds_run<4>(w, tw_a + 4); ct_level<2>(a, w, q, qinv);
      lds_run<8>(w, tw_a + 8); ct_level<3>(a, w, q, qinv);
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<double *>(data + ad.A(j)) = a[j];
      __syncwarp();
#pragma unroll
      for (int mm = 0; mm < 8; ++mm) {
        const double2 v = *reinterpret_cast<const double2 *>(data + ad.B(mm));
        a[2 * mm] = v.x; a[2 * mm + 1] = v.y;
      }
      w[0] = blob[256 + tid]; ct_level<0>(a, w, q, qinv);
      lds_run<2>(w, blob + 512 + 2 * tid); ct_level<1>(a, w, q, qinv);
      lds_run<2>(w, blob + 1024 + 2 * tid);
      { const double2 v = *reinterpret_cast<const double2 *>(blob + 1536 + 2 * tid); w[2] = v.x; w[3] = v.y; }
      ct_level<2>(a, w, q, qinv);
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) {
        const double2 v = *reinterpret_cast<const double2 *>(blob + 2048 + 512 * mm + 2 * tid);
        w[2 * mm] = v.x; w[2 * mm + 1] = v.y;
      }
      ct_level<3>(a, w, q, qinv);
      if constexpr (!FUSE) {
        if (l.out_f64) {  // uniform: the consumer takes the lazy sums as they are
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) *reinterpret_cast<double2 *>(data + ad.B(mm)) = make_double2(a[2 * mm], a[2 * mm + 1]);
        } else {
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            // values are (true - h) mod q (bias planted by the column pass): centred remainder + h is canonical
            const u64 v0 = (u64)__double_as_longlong(reduce_signed(a[2 * mm], q, qinv) + hb) & 0x000FFFFFFFFFFFFFull;
            const u64 v1 = (u64)__double_as_longlong(reduce_signed(a[2 * mm + 1], q, qinv) + hb) & 0x000FFFFFFFFFFFFFull;
            *reinterpret_cast<ulonglong2 *>(data + ad.B(mm)) = make_ulonglong2(v0, v1);
          }
        }
        __syncwarp();
        u64 *outp = base_out(limb, b, p) + tile_off + (size_t)rr * 256 + (l16 >> 3) * 16 + (l16 & 7) * 2;
#pragma unroll
        for (int mm = 0; mm < 8; ++mm)
          *reinterpret_cast<ulonglong2 *>(outp + 32 * mm) = *reinterpret_cast<const ulonglong2 *>(data + ad.C(mm));
      } else {
        // fused epilogue (see NttFuse): the raw lazy sums go through the swizzled tile, then every lane handles the 16-byte
        // chunks it will store, so x, z and dst are all accessed 256 B per half-warp
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) *reinterpret_cast<double2 *>(data + ad.B(mm)) = make_double2(a[2 * mm], a[2 * mm + 1]);
        __syncwarp();
        const NttFuse &f = l.fuse;
        const long long fb = p / f.n_c, fc = p % f.n_c;  // n_batch == 1 for fused launches
        const size_t nn = (size_t)1 << logN;
        const size_t ci = tile_off + (size_t)rr * 256 + (l16 >> 3) * 16 + (l16 & 7) * 2;
        const u64 *xs = f.x + fc * f.x_c_stride + fb * f.x_b_stride + (size_t)limb * nn;
        const bool has_z = f.z != nullptr && ((f.z_mask >> fc) & 1u);
        const u64 *zs = has_z ? f.z + fc * f.z_c_stride + fb * f.z_b_stride + (size_t)limb * nn : nullptr;
        u64 *dp = f.dst + fc * f.dst_c_stride + fb * f.dst_b_stride + (size_t)limb * nn + ci;
        const double2 cst = f.cst[limb];
        auto load_chunk = [&](const u64 *slot, int packed, int mm) -> ulonglong2 {
          return packed ? ld_packed2(slot, nn, (ci >> 1) + 16 * mm) : __ldg(reinterpret_cast<const ulonglong2 *>(slot + ci + 32 * mm));
        };
        ulonglong2 xv[8];
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) xv[mm] = load_chunk(xs, f.x_packed, mm);
        if (f.cst2 != nullptr) {  // ((x * c + z) - y) * c2: z is mandatory
          const double2 cst2 = f.cst2[limb];
          ulonglong2 zv[8];
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) zv[mm] = load_chunk(zs, f.z_packed, mm);
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            const double2 y = *reinterpret_cast<const double2 *>(data + ad.C(mm));
            const double u0 = mulmod_const(u64_to_f64(xv[mm].x), cst.x, cst.y, q) + u64_to_f64(zv[mm].x) - y.x;
            const double u1 = mulmod_const(u64_to_f64(xv[mm].y), cst.x, cst.y, q) + u64_to_f64(zv[mm].y) - y.y;
            const u64 v0 = f64_to_canonical(mulmod_const(u0, cst2.x, cst2.y, q), mc.qi);
            const u64 v1 = f64_to_canonical(mulmod_const(u1, cst2.x, cst2.y, q), mc.qi);
            *reinterpret_cast<ulonglong2 *>(dp + 32 * mm) = make_ulonglong2(v0, v1);
          }
        } else if (has_z) {
          ulonglong2 zv[8];
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) zv[mm] = load_chunk(zs, f.z_packed, mm);
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            const double2 y = *reinterpret_cast<const double2 *>(data + ad.C(mm));
            // z - h rides on the integer -> double conversion: (2^52 | z) - (2^52 + h)
            const double z0 = __longlong_as_double((long long)(zv[mm].x | 0x4330000000000000ull)) - hb;
            const double z1 = __longlong_as_double((long long)(zv[mm].y | 0x4330000000000000ull)) - hb;
            const double r0 = mulmod_const(u64_to_f64(xv[mm].x) - y.x, cst.x, cst.y, q) + z0;
            const double r1 = mulmod_const(u64_to_f64(xv[mm].y) - y.y, cst.x, cst.y, q) + z1;
            const u64 v0 = (u64)__double_as_longlong(reduce_signed(r0, q, qinv) + hb) & 0x000FFFFFFFFFFFFFull;
            const u64 v1 = (u64)__double_as_longlong(reduce_signed(r1, q, qinv) + hb) & 0x000FFFFFFFFFFFFFull;
            *reinterpret_cast<ulonglong2 *>(dp + 32 * mm) = make_ulonglong2(v0, v1);
          }
        } else {
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            const double2 y = *reinterpret_cast<const double2 *>(data + ad.C(mm));
            const u64 v0 = f64_to_canonical(mulmod_const(u64_to_f64(xv[mm].x) - y.x, cst.x, cst.y, q), mc.qi);
            const u64 v1 = f64_to_canonical(mulmod_const(u64_to_f64(xv[mm].y) - y.y, cst.x, cst.y, q), mc.qi);
            *reinterpret_cast<ulonglong2 *>(dp + 32 * mm) = make_ulonglong2(v0, v1);
          }
        }
      }
    } else {
#pragma unroll
      for (int mm = 0; mm < 8; ++mm) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(data + ad.B(mm));
        a[2 * mm] = u64_to_f64(v.x); a[2 * mm + 1] = u64_to_f64(v.y);
      }
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) {
        const double2 v = *reinterpret_cast<const double2 *>(blob + 2048 + 512 * mm + 2 * tid);
        w[2 * mm] = v.x; w[2 * mm + 1] = v.y;
      }
      gs_level<3>(a, w, q, qinv);
      lds_run<2>(w, blob + 1024 + 2 * tid);
      { const double2 v = *reinterpret_cast<const double2 *>(blob + 1536 + 2 * tid); w[2] = v.x; w[3] = v.y; }
      gs_level<2>(a, w, q, qinv);
      lds_run<2>(w, blob + 512 + 2 * tid); gs_level<1>(a, w, q, qinv);
      w[0] = blob[256 + tid]; gs_level<0>(a, w, q, qinv);
#pragma unroll
      for (int mm = 0; mm < 8; ++mm) *reinterpret_cast<double2 *>(data + ad.B(mm)) = make_double2(a[2 * mm], a[2 * mm + 1]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = *reinterpret_cast<const double *>(data + ad.A(j));
      lds_run<8>(w, tw_a + 8); gs_level<3>(a, w, q, qinv);
      lds_run<4>(w, tw_a + 4); gs_level<2>(a, w, q, qinv);
      lds_run<2>(w, tw_a + 2); gs_level<1>(a, w, q, qinv);
      lds_run<1>(w, tw_a + 1); gs_level<0>(a, w, q, qinv);
      // the sums have grown to <= 2^8 q: bring them back to |v| <= q/2 before the column pass doubles them again
      double *outd = reinterpret_cast<double *>(base_out(limb, b, p)) + tile_off + (size_t)rr * 256 + l16;
#pragma unroll
      for (int j = 0; j < 16; ++j) outd[16 * j] = reduce_signed(a[j], q, qinv);
    }
  };

  // ---- the unit pipeline
  pdl_wait();  // the queue counters and every buffer belong to the stream's earlier kernels until here
  if (tid == 0) pop();
  __syncthreads();
  ItemDesc cur = next_item;
  int cur_m = 0, stg = 0;
  bool blob_pending = false;  // a blob copy is in flight for `cur`
  auto start_item_loads = [&](const ItemDesc &it, int m, unsigned st, bool blob_free) -> bool {
    // returns true when the item's blob (row items) still has to be requested later (the buffer is in use)
    bool need_blob_later = false;
    if (is_col(it.second)) issue_col(it, m, st);
    else {
      issue_row(it, m, st);
      if (m == 0 && (it.mi != blob_mi || it.t != blob_t)) {
        if (blob_free) { if (tid == 0) issue_blob(it); blob_mi = it.mi; blob_t = it.t; blob_pending = true; }
        else need_blob_later = true;
      }
    }
    return need_blob_later;
  };
  if (cur.valid) {
    if (!cur.ready) wait_ready(cur.sj);
    start_item_loads(cur, 0, stage0_u32, true);
    if (tid == 0 && cur.cnt == 1) popped = (int)atomicAdd(&ctrl[0], 1u);
  }
  cp_async_commit();
  while (cur.valid) {
    // the unit after `cur`: the next member of the item, or the first member of the next item of the queue
    ItemDesc nxt = cur;
    int nxt_m = cur_m + 1;
    const bool last_member = nxt_m == cur.cnt;
    if (last_member && tid == 0) pop();
    cp_async_wait<0>();
    __syncthreads();  // `cur` has landed for every thread; everybody is done with the other stage; next_item is visible
    if (last_member) { nxt = next_item; nxt_m = 0; }
    bool issued = false, blob_later = false;
    const bool cur_uses_blob = !is_col(cur.second);
    if (nxt.valid && (nxt_m > 0 || nxt.ready)) {
      blob_later = start_item_loads(nxt, nxt_m, stage0_u32 + (stg ^ 1) * FUSED_STAGE_BYTES, !cur_uses_blob || !last_member);
      issued = true;
    }
    cp_async_commit();
    if (tid == 0 && nxt.valid && nxt_m + 1 == nxt.cnt) popped = (int)atomicAdd(&ctrl[0], 1u);  // needed at the next loop top
    unsigned char *sp = stage0 + stg * FUSED_STAGE_BYTES;
    if (is_col(cur.second)) {
      if constexpr (INV) compute_col_inv(cur, cur_m, sp); else compute_col_fwd(cur, cur_m, sp);
    } else {
      if (blob_pending) { mbar_wait(&bar, blob_phase); blob_phase ^= 1; blob_pending = false; }
      compute_row(cur, cur_m, sp);
    }
    if (last_member) {
      // item done: first-pass items release their super-job's counter; second-pass items count down to the reset
      __syncthreads();
      if (tid == 0) {
        if (!cur.second) {
          __threadfence();
          atomicAdd(&ctrl[2 + cur.sj], 1u);
        } else if (atomicAdd(&ctrl[2 + FUSED_MAX_SJ + cur.sj], 1u) == (unsigned)fp.T - 1) {
          ctrl[2 + cur.sj] = 0; ctrl[2 + FUSED_MAX_SJ + cur.sj] = 0;  // every consumer of the super-job has long passed its check
        }
      }
      if (nxt.valid && !issued) {  // rare: the producers of the next item were still running when it was popped
        wait_ready(nxt.sj);
        blob_later = start_item_loads(nxt, 0, stage0_u32 + (stg ^ 1) * FUSED_STAGE_BYTES, true);
        cp_async_commit();
      } else if (blob_later) {  // the blob buffer was busy with `cur`
        if (tid == 0) issue_blob(nxt);
        blob_mi = nxt.mi; blob_t = nxt.t; blob_pending = true;
      }
    }
    cur = nxt; cur_m = nxt_m; stg ^= 1;
  }
  cp_async_wait<0>();
  // the last CTA to leave resets the queue for the next launch
  __syncthreads();
  if (tid == 0) {
    if (atomicAdd(&ctrl[1], 1u) == gridDim.x - 1) { ctrl[0] = 0; ctrl[1] = 0; __threadfence(); }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int fused_sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}

int ntt_fused_enabled() {
  static const int v = [] {
    // Opt-in (HML_NTT_FUSED=1).  Measured on B200 (profiles/ntt_fused_r2.md): the single launch brings the DRAM traffic of a
    // transform down to the algorithmic 2 W per limb, but one kernel has to run both unit kinds at the row pass's register
    // budget (two CTAs per SM instead of three for the column units) and pays a block barrier per unit: 0.438 vs 0.370 us per
    // limb batched, 69 vs 60 us for one ciphertext's 115-limb ModUp launch.  The two-kernel transforms stay the default.
    const char *e = getenv("HML_NTT_FUSED");
    return e && atoi(e) == 1 ? 1 : 0;
  }();
  return v;
}

static bool make_plan(int logN, const LimbMap &lm, const NttLaunch &l, FusedPlan &fp) {
  fp = FusedPlan{};
  fp.T = 1 << (logN - 12);
  static const int g_env = [] { const char *e = getenv("HML_NTT_G"); return e ? atoi(e) : 0; }();
  // grouping members amortises the row-twiddle blob but makes the work items coarser: group only when the queue stays long
  const int resident = 2 * fused_sm_count();
  for (int G = g_env > 0 ? std::min(g_env, FUSED_G) : FUSED_G;; G >>= 1) {
    int sj = 0;
    bool ok = true;
    for (int i = 0; i < l.n_limbs && ok; ++i) {
      const int members = l.n_batch * (l.n_polys - (lm.skip[i] != 0xFF && lm.skip[i] < l.n_polys ? 1 : 0));
      fp.sj_first[i] = (uint16_t)sj;
      fp.members[i] = (uint16_t)members;
      ok = members <= 65535;
      sj += (members + G - 1) / G;
      ok = ok && sj <= FUSED_MAX_SJ;
    }
    if (G > 1 && g_env <= 0 && (!ok || sj * fp.T < 8 * resident)) continue;
    if (!ok) return false;
    fp.sj_first[l.n_limbs] = (uint16_t)sj;
    fp.n_sj = sj;
    fp.G = G;
    break;
  }
  fp.n_items = 2 * fp.n_sj * fp.T;
  // enough first-pass items ahead of the first consumer that (a) consumers practically never wait and (b) the intermediate
  // of the super-jobs in between (lag * G * 0.5 MB) stays far below the L2 capacity
  static const int lag_env = [] { const char *e = getenv("HML_NTT_LAG"); return e ? atoi(e) : 0; }();
  const int lag = lag_env > 0 ? lag_env : (3 * resident + fp.T - 1) / fp.T;
  fp.lag = std::max(1, std::min(lag, fp.n_sj));
  return fp.n_sj > 0;
}

template <int LOGR1, bool INV>
static void launch_fused_t(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, const FusedPlan &fp, unsigned *ctrl, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(ntt_fused<LOGR1, INV, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
    if constexpr (!INV) {
      cudaFuncSetAttribute(ntt_fused<LOGR1, INV, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
      cudaFuncSetAttribute(ntt_fused<LOGR1, INV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
      cudaFuncSetAttribute(ntt_fused<LOGR1, INV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
    }
  }
  const int grid = std::min(fp.n_items, 2 * fused_sm_count());
  const bool f64 = !INV && l.in_f64, fuse = !INV && l.fuse.x != nullptr;
  if constexpr (INV) {
    launch_pdl(ntt_fused<LOGR1, true, false, false>, grid, NTT_THREADS, FUSED_SMEM_BYTES, s, t, logN, lm, l, fp, ctrl);
  } else {
    if (f64 && fuse) launch_pdl(ntt_fused<LOGR1, false, true, true>, grid, NTT_THREADS, FUSED_SMEM_BYTES, s, t, logN, lm, l, fp, ctrl);
    else if (f64) launch_pdl(ntt_fused<LOGR1, false, true, false>, grid, NTT_THREADS, FUSED_SMEM_BYTES, s, t, logN, lm, l, fp, ctrl);
    else if (fuse) launch_pdl(ntt_fused<LOGR1, false, false, true>, grid, NTT_THREADS, FUSED_SMEM_BYTES, s, t, logN, lm, l, fp, ctrl);
    else launch_pdl(ntt_fused<LOGR1, false, false, false>, grid, NTT_THREADS, FUSED_SMEM_BYTES, s, t, logN, lm, l, fp, ctrl);
  }
}

// returns false when the launch shape is outside the fused kernel (the caller falls back to the two-kernel path)
bool launch_ntt_fused(bool inverse, const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, unsigned *ctrl, cudaStream_t s) {
  if (!ctrl || logN <= NTT_SMALL_LOG || logN > 16 || !ntt_fused_enabled()) return false;
  if (l.in_galois || l.side_out || l.fuse.z_galois) return false;  // automorphism on load: two-kernel path only
  FusedPlan fp;
  if (!make_plan(logN, lm, l, fp)) return false;
  switch (logN - NTT_ROW_LOG) {
    case 5: inverse ? launch_fused_t<5, true>(t, logN, lm, l, fp, ctrl, s) : launch_fused_t<5, false>(t, logN, lm, l, fp, ctrl, s); break;
    case 6: inverse ? launch_fused_t<6, true>(t, logN, lm, l, fp, ctrl, s) : launch_fused_t<6, false>(t, logN, lm, l, fp, ctrl, s); break;
    case 7: inverse ? launch_fused_t<7, true>(t, logN, lm, l, fp, ctrl, s) : launch_fused_t<7, false>(t, logN, lm, l, fp, ctrl, s); break;
    case 8: inverse ? launch_fused_t<8, true>(t, logN, lm, l, fp, ctrl, s) : launch_fused_t<8, false>(t, logN, lm, l, fp, ctrl, s); break;
    default: return false;
  }
  return true;
}

size_t ntt_fused_ctrl_words() { return 2 + 2 * (size_t)FUSED_MAX_SJ; }

}  // namespace hml
