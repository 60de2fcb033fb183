// Negacyclic NTT / inverse NTT per RNS limb on B200 (sm_100a), FP64 datapath (see modarith.cuh).
//
// Executes the reference's NTT / INTT instruction classes (InsGen::GenNTT, reference
// src/InsGen.cpp:17-44; simulated unit NTTU, reference src/Components.cpp:380-569 — which is itself a
// two-phase + transpose structure, :397-431) on real data.
//
// Algorithm: merged-psi Cooley-Tukey (forward, natural in -> bit-reversed out) / Gentleman-Sande
// (inverse), split 4-step style into TWO passes so that every pass works on a 4096-element tile held in
// registers + shared memory:
//   forward  pass 1 "columns": the first log2(R1) stages couple elements R2 apart (R2 = 256, R1 = N/256);
//            a CTA owns C = 4096/R1 adjacent columns (>= 128 B contiguous per row -> full-line accesses).
//   forward  pass 2 "rows":    the last 8 stages stay inside contiguous 256-element rows; a CTA owns 16 rows.
//   inverse  = the mirror image (rows first, then columns).
// Inside a pass each thread keeps 8 points in registers and runs up to 3 radix-2 stages per "round";
// rounds exchange data through shared memory laid out [point][column] (conflict-free: a warp always
// touches whole rows; the row pass uses pitch C+1 for its transposing accesses).
// The buffer between the passes holds raw signed-lazy doubles (no conversion / reduction cost).
#pragma once
#include "modarith.cuh"

namespace hml {

constexpr int NTT_TILE = 4096;     // elements per CTA
constexpr int NTT_THREADS = 512;   // 8 elements per thread
constexpr int NTT_ROW_LOG = 8;     // R2 = 256: contiguous row length handled by the row pass
constexpr int NTT_MAX_LIMBS = 128; // limbs (x polys) per launch

struct LimbMap {
  uint16_t mod[NTT_MAX_LIMBS];  // modulus index of each limb of the launch
  uint16_t pos[NTT_MAX_LIMBS];  // NTT launches: limb slot inside the buffer (address = base + pos * limb_stride)
  uint8_t skip[NTT_MAX_LIMBS];  // NTT launches: poly index NOT to transform for this limb (0xFF = none); used by
                                // ModUp, where digit j's own limbs stay as they are
};

// Twiddle tables: per modulus, N entries of (w, RN(w/q)) as double2, index = bit-reversed exponent
struct NttTables {
  const double2 *fwd;      // [n_mod][N]
  const double2 *inv;      // [n_mod][N]
  const ModConst *mc;      // [n_mod]
};

// ---- launch descriptors (host side, ntt.cu)
struct NttLaunch {
  const u64 *in;        // [n_polys][n_limbs][N] (poly stride / limb stride in elements below)
  u64 *out;
  long long in_poly_stride, in_limb_stride;    // in_limb_stride = 0 broadcasts one source limb (rescale)
  long long out_poly_stride, out_limb_stride;
  int n_limbs, n_polys;
  int n_batch;                                  // independent ciphertexts: item (b, p) lives at base + b * batch_stride + p * poly_stride
  long long in_batch_stride, out_batch_stride;
  // inverse only: per-limb post-scale constant c (folded with N^-1 on the host): out = INTT(in) * c, canonical.
  const double2 *post_scale;  // [n_limbs] (c*ninv mod q, RN(that / q)) or nullptr for plain N^-1
};

void launch_ntt_forward(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s);
void launch_ntt_inverse(const NttTables &t, int logN, const LimbMap &lm, const NttLaunch &l, cudaStream_t s);

}  // namespace hml
