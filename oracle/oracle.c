/*
 * TEST INFRASTRUCTURE ONLY — scalar textbook RNS-CKKS oracle (plain C).  See oracle.h for the
 * parity statement and the reference file:line each composite op follows.
 *
 * Style: deliberately naive.  Every modular product is a 128-bit product followed by `%`.
 * No Montgomery/Barrett/Shoup tricks, no lazy ranges — those live in the CUDA path this file checks.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

struct orc_params {
  uint32_t N, logN, word_bits, max_level, alpha, n_mod;
  uint64_t *mod;      /* [n_mod] q_0..q_{max_level-1}, p_0..p_{alpha-1} */
  uint64_t *psi;      /* [n_mod] primitive 2N-th root */
  uint64_t *psi_inv;  /* [n_mod] */
  uint64_t *n_inv;    /* [n_mod] N^-1 */
  uint64_t **psi_brv;     /* [n_mod][N] psi^bitrev(i) */
  uint64_t **psi_inv_brv; /* [n_mod][N] psi^-bitrev(i) */
  int use_direct;
};

static int g_threads = 1;

/* ---------------------------------------------------------------- scalar modular arithmetic */
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((u128)a * b) % m); }
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t m) { uint64_t s = a + b; return s >= m ? s - m : s; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t m) { return a >= b ? a - b : a + m - b; }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t m) {
  uint64_t r = 1 % m;
  a %= m;
  while (e) { if (e & 1) r = mulmod(r, a, m); a = mulmod(a, a, m); e >>= 1; }
  return r;
}
static uint64_t invmod(uint64_t a, uint64_t m) { return powmod(a, m - 2, m); } /* m prime */

static int is_prime(uint64_t n) {
  static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (unsigned i = 0; i < 12; i++) { if (n == bases[i]) return 1; if (n % bases[i] == 0) return 0; }
  uint64_t d = n - 1; int s = 0;
  while ((d & 1) == 0) { d >>= 1; s++; }
  for (unsigned i = 0; i < 12; i++) {
    uint64_t x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; r++) { x = mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
    if (comp) return 0;
  }
  return 1;
}

static uint32_t bitrev(uint32_t x, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

/* ---------------------------------------------------------------- parameters */
orc_params *orc_create(uint32_t N, uint32_t word_bits, uint32_t max_level, uint32_t alpha) {
  if (N < 4 || (N & (N - 1)) || word_bits < 20 || word_bits > 60 || max_level == 0 || alpha == 0) return NULL;
  orc_params *p = (orc_params *)calloc(1, sizeof(*p));
  p->N = N; p->word_bits = word_bits; p->max_level = max_level; p->alpha = alpha;
  p->n_mod = max_level + alpha;
  while ((1u << p->logN) < N) p->logN++;
  p->mod = (uint64_t *)calloc(p->n_mod, 8);
  p->psi = (uint64_t *)calloc(p->n_mod, 8);
  p->psi_inv = (uint64_t *)calloc(p->n_mod, 8);
  p->n_inv = (uint64_t *)calloc(p->n_mod, 8);
  p->psi_brv = (uint64_t **)calloc(p->n_mod, sizeof(uint64_t *));
  p->psi_inv_brv = (uint64_t **)calloc(p->n_mod, sizeof(uint64_t *));
  const uint64_t twoN = 2ull * N, hi = 1ull << word_bits, lo = 1ull << (word_bits - 1);
  /* largest candidate = 1 (mod 2N) strictly below 2^w */
  uint64_t cand = ((hi - 2) / twoN) * twoN + 1;
  uint32_t found = 0;
  while (found < p->n_mod && cand > lo) {
    if (is_prime(cand)) p->mod[found++] = cand;
    cand -= twoN;
  }
  if (found < p->n_mod) { orc_destroy(p); return NULL; }
  for (uint32_t i = 0; i < p->n_mod; i++) {
    const uint64_t m = p->mod[i];
    uint64_t psi = 0;
    for (uint64_t x = 2; x < m; x++) {
      uint64_t c = powmod(x, (m - 1) / twoN, m);
      if (powmod(c, N, m) == m - 1) { psi = c; break; } /* c^N = -1  <=>  order exactly 2N */
    }
    p->psi[i] = psi;
    p->psi_inv[i] = invmod(psi, m);
    p->n_inv[i] = invmod(N % m, m);
    p->psi_brv[i] = (uint64_t *)malloc(8ull * N);
    p->psi_inv_brv[i] = (uint64_t *)malloc(8ull * N);
    uint64_t pw = 1, pwi = 1;
    for (uint32_t e = 0; e < N; e++) {
      uint32_t r = bitrev(e, p->logN);
      p->psi_brv[i][r] = pw;
      p->psi_inv_brv[i][r] = pwi;
      pw = mulmod(pw, psi, m);
      pwi = mulmod(pwi, p->psi_inv[i], m);
    }
  }
  return p;
}

void orc_destroy(orc_params *p) {
  if (!p) return;
  for (uint32_t i = 0; i < p->n_mod; i++) {
    if (p->psi_brv && p->psi_brv[i]) free(p->psi_brv[i]);
    if (p->psi_inv_brv && p->psi_inv_brv[i]) free(p->psi_inv_brv[i]);
  }
  free(p->psi_brv); free(p->psi_inv_brv);
  free(p->mod); free(p->psi); free(p->psi_inv); free(p->n_inv);
  free(p);
}

uint32_t orc_n_moduli(const orc_params *p) { return p->n_mod; }
uint64_t orc_modulus(const orc_params *p, uint32_t idx) { return p->mod[idx]; }
uint64_t orc_psi(const orc_params *p, uint32_t idx) { return p->psi[idx]; }
void orc_set_direct(orc_params *p, int use_direct) { p->use_direct = use_direct; }
int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n < 1) n = omp_get_max_threads();
  g_threads = n;
#else
  (void)n; g_threads = 1;
#endif
  return g_threads;
}

/* ---------------------------------------------------------------- NTT: definition (T1) and fast (T2) */
void orc_ntt_direct(const orc_params *p, uint32_t mi, const uint64_t *a, uint64_t *out) {
  const uint64_t m = p->mod[mi];
  const uint32_t N = p->N;
  for (uint32_t k = 0; k < N; k++) {
    /* evaluation point psi^(2*brv(k)+1); Horner */
    uint64_t x = powmod(p->psi[mi], 2ull * bitrev(k, p->logN) + 1, m);
    uint64_t acc = 0;
    for (uint32_t n = N; n-- > 0;) acc = addmod(mulmod(acc, x, m), a[n] % m, m);
    out[k] = acc;
  }
}

void orc_intt_direct(const orc_params *p, uint32_t mi, const uint64_t *a, uint64_t *out) {
  /* a[n] = N^-1 * sum_k ahat[k] * x_k^-n,  x_k = psi^(2 brv(k)+1) */
  const uint64_t m = p->mod[mi];
  const uint32_t N = p->N;
  uint64_t *xinv = (uint64_t *)malloc(8ull * N);
  for (uint32_t k = 0; k < N; k++) xinv[k] = powmod(p->psi_inv[mi], 2ull * bitrev(k, p->logN) + 1, m);
  uint64_t *cur = (uint64_t *)malloc(8ull * N); /* x_k^-n, updated per n */
  for (uint32_t k = 0; k < N; k++) cur[k] = 1;
  for (uint32_t n = 0; n < N; n++) {
    uint64_t acc = 0;
    for (uint32_t k = 0; k < N; k++) {
      acc = addmod(acc, mulmod(a[k] % m, cur[k], m), m);
      cur[k] = mulmod(cur[k], xinv[k], m);
    }
    out[n] = mulmod(acc, p->n_inv[mi], m);
  }
  free(cur); free(xinv);
}

/* Cooley–Tukey, natural in -> bit-reversed out, psi powers merged (textbook negacyclic NTT). */
void orc_ntt(const orc_params *p, uint32_t mi, uint64_t *a) {
  const uint64_t m = p->mod[mi];
  const uint64_t *w = p->psi_brv[mi];
  uint32_t t = p->N;
  for (uint32_t mm = 1; mm < p->N; mm <<= 1) {
    t >>= 1;
    for (uint32_t i = 0; i < mm; i++) {
      const uint64_t s = w[mm + i];
      const uint32_t j1 = 2 * i * t;
      for (uint32_t j = j1; j < j1 + t; j++) {
        uint64_t u = a[j], v = mulmod(a[j + t], s, m);
        a[j] = addmod(u, v, m);
        a[j + t] = submod(u, v, m);
      }
    }
  }
}

/* Gentleman–Sande, bit-reversed in -> natural out, times N^-1. */
void orc_intt(const orc_params *p, uint32_t mi, uint64_t *a) {
  const uint64_t m = p->mod[mi];
  const uint64_t *w = p->psi_inv_brv[mi];
  uint32_t t = 1;
  for (uint32_t mm = p->N >> 1; mm >= 1; mm >>= 1) {
    for (uint32_t i = 0; i < mm; i++) {
      const uint64_t s = w[mm + i];
      const uint32_t j1 = 2 * i * t;
      for (uint32_t j = j1; j < j1 + t; j++) {
        uint64_t u = a[j], v = a[j + t];
        a[j] = addmod(u, v, m);
        a[j + t] = mulmod(submod(u, v, m), s, m);
      }
    }
    t <<= 1;
  }
  for (uint32_t j = 0; j < p->N; j++) a[j] = mulmod(a[j], p->n_inv[mi], m);
}

static void ntt_any(const orc_params *p, uint32_t mi, uint64_t *a) {
  if (!p->use_direct) { orc_ntt(p, mi, a); return; }
  uint64_t *t = (uint64_t *)malloc(8ull * p->N);
  orc_ntt_direct(p, mi, a, t); memcpy(a, t, 8ull * p->N); free(t);
}
static void intt_any(const orc_params *p, uint32_t mi, uint64_t *a) {
  if (!p->use_direct) { orc_intt(p, mi, a); return; }
  uint64_t *t = (uint64_t *)malloc(8ull * p->N);
  orc_intt_direct(p, mi, a, t); memcpy(a, t, 8ull * p->N); free(t);
}

void orc_negacyclic_schoolbook(const orc_params *p, uint32_t mi, const uint64_t *a, const uint64_t *b, uint64_t *out) {
  const uint64_t m = p->mod[mi];
  const uint32_t N = p->N;
  memset(out, 0, 8ull * N);
  for (uint32_t i = 0; i < N; i++)
    for (uint32_t j = 0; j < N; j++) {
      uint64_t pr = mulmod(a[i] % m, b[j] % m, m);
      uint32_t k = i + j;
      if (k < N) out[k] = addmod(out[k], pr, m);
      else out[k - N] = submod(out[k - N], pr, m);
    }
}

/* ---------------------------------------------------------------- EWE / automorphism */
void orc_ewe(const orc_params *p, uint32_t mi, const uint64_t *x1, const uint64_t *x2, const uint64_t *x3,
             const uint64_t *x4, int sub, uint64_t *out) {
  const uint64_t m = p->mod[mi];
  for (uint32_t n = 0; n < p->N; n++) {
    uint64_t a = 0, b = 0;
    if (x1) a = x2 ? mulmod(x1[n] % m, x2[n] % m, m) : x1[n] % m;
    if (x3) b = x4 ? mulmod(x3[n] % m, x4[n] % m, m) : x3[n] % m;
    out[n] = sub ? submod(a, b, m) : addmod(a, b, m);
  }
}

void orc_automorph_index(const orc_params *p, uint64_t g, uint32_t *perm) {
  const uint64_t twoN = 2ull * p->N;
  for (uint32_t k = 0; k < p->N; k++) {
    uint64_t e = (g % twoN) * (2ull * bitrev(k, p->logN) + 1) % twoN; /* odd */
    perm[k] = bitrev((uint32_t)((e - 1) / 2), p->logN);
  }
}

void orc_automorph_eval(const orc_params *p, uint64_t g, const uint64_t *in, uint64_t *out) {
  uint32_t *perm = (uint32_t *)malloc(4ull * p->N);
  orc_automorph_index(p, g, perm);
  for (uint32_t k = 0; k < p->N; k++) out[k] = in[perm[k]];
  free(perm);
}

void orc_automorph_coeff(const orc_params *p, uint32_t mi, uint64_t g, const uint64_t *in, uint64_t *out) {
  /* a(X) -> a(X^g): coefficient n moves to g*n mod 2N, negated when it wraps past N */
  const uint64_t m = p->mod[mi], twoN = 2ull * p->N;
  for (uint32_t n = 0; n < p->N; n++) {
    uint64_t e = (g % twoN) * n % twoN;
    if (e < p->N) out[e] = in[n] % m;
    else out[e - p->N] = submod(0, in[n] % m, m);
  }
}

/* ---------------------------------------------------------------- base conversion */
void orc_bconv(const orc_params *p, const uint32_t *src, uint32_t n_src, uint32_t dst, const uint64_t *x, uint64_t *out) {
  const uint64_t m = p->mod[dst];
  const uint32_t N = p->N;
  uint64_t *hat_inv = (uint64_t *)malloc(8ull * n_src); /* (D/s_i)^-1 mod s_i */
  uint64_t *hat_m = (uint64_t *)malloc(8ull * n_src);   /* (D/s_i) mod m */
  for (uint32_t i = 0; i < n_src; i++) {
    const uint64_t si = p->mod[src[i]];
    uint64_t a = 1, b = 1;
    for (uint32_t j = 0; j < n_src; j++) {
      if (j == i) continue;
      a = mulmod(a, p->mod[src[j]] % si, si);
      b = mulmod(b, p->mod[src[j]] % m, m);
    }
    hat_inv[i] = invmod(a, si);
    hat_m[i] = b;
  }
  for (uint32_t n = 0; n < N; n++) {
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n_src; i++) {
      const uint64_t si = p->mod[src[i]];
      uint64_t y = mulmod(x[(size_t)i * N + n] % si, hat_inv[i], si); /* step 1 */
      acc = addmod(acc, mulmod(y % m, hat_m[i], m), m);              /* step 2 */
    }
    out[n] = acc;
  }
  free(hat_inv); free(hat_m);
}

/* ---------------------------------------------------------------- key switch (K1..K10) */
#define EXT_MOD(e) ((e) < L ? (e) : p->max_level + ((e) - L))
#define EVK_LIMB(e) ((e) < L ? (e) : evk_q_limbs + ((e) - L))

/* K1..K4: the beta extended digits of d, evaluation form: t [beta][L + alpha][N].  Digit j keeps the residues of its own
 * limbs and is base-converted (K2 + K3) to every other modulus of E = (q_0..q_{L-1}, p_0..p_{alpha-1}). */
void orc_modup(const orc_params *p, uint32_t L, const uint64_t *d, uint64_t *t) {
  const uint32_t N = p->N, A = p->alpha, E = L + A, beta = (L + A - 1) / A;
  const size_t W = N;
  uint64_t *dc = (uint64_t *)malloc(8 * W * L);             /* K1: coefficient form of d */
  memcpy(dc, d, 8 * W * L);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
  for (uint32_t i = 0; i < L; i++) intt_any(p, i, dc + i * W); /* K1 */
  for (uint32_t j = 0; j < beta; j++) {
    const uint32_t lo = j * A, aj = (L - lo < A) ? (L - lo) : A;
    uint32_t src[64 * 8];
    for (uint32_t i = 0; i < aj; i++) src[i] = lo + i;
    uint64_t *tj = t + (size_t)j * E * W;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (uint32_t e = 0; e < E; e++) {
      if (e >= lo && e < lo + aj) memcpy(tj + e * W, dc + e * W, 8 * W);
      else orc_bconv(p, src, aj, EXT_MOD(e), dc + (size_t)lo * W, tj + e * W);
      ntt_any(p, EXT_MOD(e), tj + e * W); /* K4 */
    }
  }
  free(dc);
}

/* K5..K10 from given extended digits t [beta][E][N] (evaluation form) */
void orc_keyswitch_digits(const orc_params *p, uint32_t L, const uint64_t *t, const uint64_t *evk, uint32_t evk_q_limbs,
                          uint64_t *out0, uint64_t *out1) {
  const uint32_t N = p->N, A = p->alpha, E = L + A, beta = (L + A - 1) / A;
  const size_t W = N;
  const uint32_t evk_limbs = evk_q_limbs + A;
  uint64_t *acc = (uint64_t *)calloc(2 * W * E, 8);         /* K5 accumulators [2][E][N] */
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
  for (uint32_t e = 0; e < E; e++) {
    const uint64_t m = p->mod[EXT_MOD(e)];
    for (uint32_t j = 0; j < beta; j++)
      for (uint32_t c = 0; c < 2; c++) {   /* K5: acc_c[e] += t_j[e] * evk[j][c][e] */
        const uint64_t *k = evk + (((size_t)j * 2 + c) * evk_limbs + EVK_LIMB(e)) * W;
        const uint64_t *tj = t + ((size_t)j * E + e) * W;
        uint64_t *a = acc + ((size_t)c * E + e) * W;
        for (uint32_t n = 0; n < N; n++) a[n] = addmod(a[n], mulmod(tj[n], k[n] % m, m), m);
      }
  }
  /* ModDown, per accumulator c */
  uint32_t psrc[64 * 8];
  for (uint32_t j = 0; j < A; j++) psrc[j] = p->max_level + j;
  for (uint32_t c = 0; c < 2; c++) {
    uint64_t *a = acc + (size_t)c * E * W;
    uint64_t *u = (uint64_t *)malloc(8 * W * A);
    memcpy(u, a + (size_t)L * W, 8 * W * A);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (uint32_t j = 0; j < A; j++) intt_any(p, p->max_level + j, u + j * W); /* K6 */
    uint64_t *out = c == 0 ? out0 : out1;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (uint32_t i = 0; i < L; i++) {
      const uint64_t m = p->mod[i];
      uint64_t *v = out + i * W;
      orc_bconv(p, psrc, A, i, u, v); /* K7+K8 */
      ntt_any(p, i, v);               /* K9 */
      uint64_t pinv = 1;              /* P^-1 mod q_i */
      for (uint32_t j = 0; j < A; j++) pinv = mulmod(pinv, p->mod[p->max_level + j] % m, m);
      pinv = invmod(pinv, m);
      for (uint32_t n = 0; n < N; n++) v[n] = mulmod(submod(a[i * W + n], v[n], m), pinv, m); /* K10 */
    }
    free(u);
  }
  free(acc);
}

void orc_keyswitch(const orc_params *p, uint32_t L, const uint64_t *d, const uint64_t *evk, uint32_t evk_q_limbs,
                   uint64_t *out0, uint64_t *out1) {
  const uint32_t A = p->alpha, E = L + A, beta = (L + A - 1) / A;
  uint64_t *t = (uint64_t *)malloc(8 * (size_t)p->N * E * beta);
  orc_modup(p, L, d, t);
  orc_keyswitch_digits(p, L, t, evk, evk_q_limbs, out0, out1);
  free(t);
}

/* Hoisted rotations: n_rot rotations of one ciphertext share ONE ModUp of c1 (its own definition, NOT bit-identical to
 * orc_hrotate: the approximate base conversion does not commute with the automorphism limb by limb).  With
 * t = ModUp(c1) (K1..K4):   out_r = (sigma_r(c0) + ks0, ks1),  (ks0, ks1) = K5..K10( sigma_r(t), rotkeys[r] ),
 * sigma_r applied to every limb of every extended digit in evaluation form.  Replaces n_rot runs of the reference's
 * HROTATE (src/Operation.cpp:1271-1358), which cannot share work between operations (:636,:675,:714). */
void orc_hrotate_hoisted(const orc_params *p, uint32_t L, const uint64_t *ct, uint32_t n_rot, const uint64_t *const *rotkeys,
                         uint32_t evk_q_limbs, const uint64_t *galois, uint64_t *const *ct_outs) {
  const uint32_t A = p->alpha, E = L + A, beta = (L + A - 1) / A;
  const size_t W = p->N, PL = W * L;
  uint64_t *t = (uint64_t *)malloc(8 * W * E * beta), *ts = (uint64_t *)malloc(8 * W * E * beta);
  uint64_t *s0 = (uint64_t *)malloc(8 * PL);
  orc_modup(p, L, ct + PL, t);
  for (uint32_t r = 0; r < n_rot; r++) {
    for (size_t k = 0; k < (size_t)beta * E; k++) orc_automorph_eval(p, galois[r], t + k * W, ts + k * W);
    for (uint32_t l = 0; l < L; l++) orc_automorph_eval(p, galois[r], ct + l * W, s0 + l * W);
    uint64_t *out = ct_outs[r];
    orc_keyswitch_digits(p, L, ts, rotkeys[r], evk_q_limbs, out, out + PL);
    for (uint32_t l = 0; l < L; l++) orc_ewe(p, l, out + l * W, NULL, s0 + l * W, NULL, 0, out + l * W);
  }
  free(t); free(ts); free(s0);
}
#undef EXT_MOD
#undef EVK_LIMB

/* ---------------------------------------------------------------- rescale */
void orc_rescale(const orc_params *p, uint32_t L, const uint64_t *in, uint64_t *out) {
  const uint32_t N = p->N;
  const size_t W = N;
  const uint64_t ql = p->mod[L - 1];
  uint64_t *r = (uint64_t *)malloc(8 * W);
  memcpy(r, in + (size_t)(L - 1) * W, 8 * W);
  intt_any(p, L - 1, r);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
  for (uint32_t l = 0; l < L - 1; l++) {
    const uint64_t m = p->mod[l];
    uint64_t *o = out + l * W;
    for (uint32_t n = 0; n < N; n++) o[n] = r[n] % m; /* plain non-negative reduction */
    ntt_any(p, l, o);
    const uint64_t qinv = invmod(ql % m, m);
    for (uint32_t n = 0; n < N; n++) o[n] = mulmod(submod(in[l * W + n] % m, o[n], m), qinv, m);
  }
  free(r);
}

/* ---------------------------------------------------------------- top-level ops */
void orc_hmult(const orc_params *p, uint32_t L, const uint64_t *a, const uint64_t *b, const uint64_t *evk,
               uint32_t evk_q_limbs, uint64_t *ct_out) {
  const size_t W = p->N, PL = W * L;
  uint64_t *d0 = (uint64_t *)malloc(8 * PL), *d1 = (uint64_t *)malloc(8 * PL), *d2 = (uint64_t *)malloc(8 * PL);
  uint64_t *k0 = (uint64_t *)malloc(8 * PL), *k1 = (uint64_t *)malloc(8 * PL);
  for (uint32_t l = 0; l < L; l++) {
    const uint64_t *a0 = a + l * W, *a1 = a + PL + l * W, *b0 = b + l * W, *b1 = b + PL + l * W;
    orc_ewe(p, l, a0, b0, NULL, NULL, 0, d0 + l * W);
    orc_ewe(p, l, a0, b1, a1, b0, 0, d1 + l * W);
    orc_ewe(p, l, a1, b1, NULL, NULL, 0, d2 + l * W);
  }
  orc_keyswitch(p, L, d2, evk, evk_q_limbs, k0, k1);
  for (uint32_t l = 0; l < L; l++) {
    orc_ewe(p, l, d0 + l * W, NULL, k0 + l * W, NULL, 0, d0 + l * W);
    orc_ewe(p, l, d1 + l * W, NULL, k1 + l * W, NULL, 0, d1 + l * W);
  }
  orc_rescale(p, L, d0, ct_out);
  orc_rescale(p, L, d1, ct_out + W * (L - 1));
  free(d0); free(d1); free(d2); free(k0); free(k1);
}

void orc_hrotate(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *rotkey, uint32_t evk_q_limbs,
                 uint64_t g, uint64_t *ct_out) {
  const size_t W = p->N, PL = W * L;
  uint64_t *s0 = (uint64_t *)malloc(8 * PL), *s1 = (uint64_t *)malloc(8 * PL);
  for (uint32_t l = 0; l < L; l++) {
    orc_automorph_eval(p, g, ct + l * W, s0 + l * W);
    orc_automorph_eval(p, g, ct + PL + l * W, s1 + l * W);
  }
  orc_keyswitch(p, L, s1, rotkey, evk_q_limbs, ct_out, ct_out + PL);
  for (uint32_t l = 0; l < L; l++) orc_ewe(p, l, ct_out + l * W, NULL, s0 + l * W, NULL, 0, ct_out + l * W);
  free(s0); free(s1);
}

void orc_hadd(const orc_params *p, uint32_t L, const uint64_t *a, const uint64_t *b, uint64_t *out) {
  const size_t W = p->N;
  for (uint32_t k = 0; k < 2; k++)
    for (uint32_t l = 0; l < L; l++) {
      size_t o = ((size_t)k * L + l) * W;
      orc_ewe(p, l, a + o, NULL, b + o, NULL, 0, out + o);
    }
}
void orc_pmult(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *pt, uint64_t *out) {
  const size_t W = p->N;
  for (uint32_t k = 0; k < 2; k++)
    for (uint32_t l = 0; l < L; l++) {
      size_t o = ((size_t)k * L + l) * W;
      orc_ewe(p, l, ct + o, pt + l * W, NULL, NULL, 0, out + o);
    }
}
/* The reference adds the plaintext to BOTH components (src/Operation.cpp:1650-1672); mirrored here. */
void orc_padd(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *pt, uint64_t *out) {
  const size_t W = p->N;
  for (uint32_t k = 0; k < 2; k++)
    for (uint32_t l = 0; l < L; l++) {
      size_t o = ((size_t)k * L + l) * W;
      orc_ewe(p, l, ct + o, NULL, pt + l * W, NULL, 0, out + o);
    }
}
