/*
 * TEST INFRASTRUCTURE ONLY — scalar textbook RNS-CKKS oracle (plain C, CPU).
 *
 * Nothing under homulator_b200/ may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / the reported
 * CPU baseline.
 *
 * PARITY STATUS: the reference (FHE-ACCELE/Homulator) computes *timing*, never ciphertext values
 * (reference include/Context.h:14,23 — Polynominal::data is never written; there is no modular
 * arithmetic anywhere in its tree) and ships no tests / golden vectors.  VALUE parity is therefore
 * "unpinned by the reference": this oracle is pinned instead against (a) first-principles O(N^2)
 * definitions implemented here (direct evaluation NTT, schoolbook negacyclic product, coefficient-domain
 * automorphism, big-integer CRT base conversion in tests/), and (b) committed golden vectors generated
 * by it (tests/golden/).  TRACE parity (instruction counts) *is* pinned by the reference itself: see
 * oracle/ref_count_harness.cpp and tests/golden/ref_counts.json.
 *
 * What it restates (structure, not values) and where the reference defines it:
 *   keyswitch stage order K1..K10     reference src/Operation.cpp:9-54, :63-590
 *   tensor product d0,d1,d2           reference src/Operation.cpp:592-739
 *   rescale                           reference src/Operation.cpp:741-911
 *   hmult composition                 reference src/Operation.cpp:913-1023
 *   hrotate composition               reference src/Operation.cpp:1271-1358
 *   hadd / pmult / padd               reference src/Operation.cpp:1114-1176, :1453-1523, :1618-1680
 *   primitive operand shapes          reference src/InsGen.cpp:17-173 (NTT/AUTO/EWE), :263-354 (BCONV)
 * The arithmetic itself follows SURVEY.md Appendix A (the shared mathematical specification).
 */
#ifndef HOMULATOR_ORACLE_H
#define HOMULATOR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params orc_params;

/* N power of two (>= 4), word_bits in [20, 60], n moduli = max_level + alpha.
 * Moduli: every prime q with 2^(w-1) < q < 2^w and q = 1 (mod 2N), scanned DOWNWARD from 2^w;
 * the first max_level are q_0.., the next alpha are p_0..   psi_m = x^((m-1)/2N) for the smallest
 * x >= 2 for which that power has order exactly 2N.  Returns NULL on failure. */
orc_params *orc_create(uint32_t N, uint32_t word_bits, uint32_t max_level, uint32_t alpha);
void orc_destroy(orc_params *p);

uint32_t orc_n_moduli(const orc_params *p);
uint64_t orc_modulus(const orc_params *p, uint32_t idx); /* idx < max_level: q_idx, else p_{idx-max_level} */
uint64_t orc_psi(const orc_params *p, uint32_t idx);

/* Evaluation form: ahat[k] = a(psi^(2*bitrev(k)+1)) mod m, bitrev over log2(N) bits. */
void orc_ntt(const orc_params *p, uint32_t mod_idx, uint64_t *a);          /* fast, in place */
void orc_intt(const orc_params *p, uint32_t mod_idx, uint64_t *a);         /* fast, in place, includes N^-1 */
void orc_ntt_direct(const orc_params *p, uint32_t mod_idx, const uint64_t *a, uint64_t *out);  /* O(N^2) definition */
void orc_intt_direct(const orc_params *p, uint32_t mod_idx, const uint64_t *a, uint64_t *out); /* O(N^2) definition */
/* tier switch for the composite ops below: 0 = fast NTT (T2), 1 = direct O(N^2) definitions (T1) */
void orc_set_direct(orc_params *p, int use_direct);
/* number of OpenMP threads the composite ops may use (1 = scalar port); returns threads in effect */
int orc_set_threads(int n);

/* out = a * b in Z_m[X]/(X^N+1), schoolbook, coefficient form */
void orc_negacyclic_schoolbook(const orc_params *p, uint32_t mod_idx, const uint64_t *a, const uint64_t *b, uint64_t *out);

/* EWE (reference opcode MULT, src/InsGen.cpp:90-102): out = x1*x2 +/- x3*x4 mod m.
 * NULL x2 / x4 = multiplier 1; NULL x1 or x3 = that product is absent (0). sub != 0 -> minus. */
void orc_ewe(const orc_params *p, uint32_t mod_idx, const uint64_t *x1, const uint64_t *x2,
             const uint64_t *x3, const uint64_t *x4, int sub, uint64_t *out);

/* Galois automorphism X -> X^g (g odd).  eval form: out[k] = in[k'], 2*brv(k')+1 = g*(2*brv(k)+1) mod 2N */
void orc_automorph_eval(const orc_params *p, uint64_t g, const uint64_t *in, uint64_t *out);
void orc_automorph_coeff(const orc_params *p, uint32_t mod_idx, uint64_t g, const uint64_t *in, uint64_t *out);
void orc_automorph_index(const orc_params *p, uint64_t g, uint32_t *perm); /* perm[k] = k' */

/* Fast base conversion, BOTH steps (step 1 = per-limb scaling, an EWE op in the reference;
 * step 2 = opcode BCONV_STEP2):  out[n] = sum_i [x_i[n] * (D/s_i)^-1]_{s_i} * [D/s_i]_m  mod m,
 * no overflow correction.  src_idx: n_src modulus indices (D = their product); x: [n_src][N]. */
void orc_bconv(const orc_params *p, const uint32_t *src_idx, uint32_t n_src, uint32_t dst_idx,
               const uint64_t *x, uint64_t *out);

/* Hybrid key switch at level L (L Q-limbs), digits of alpha limbs.
 * d: [L][N] evaluation form.  evk: [beta][2][evk_q_limbs + alpha][N], evk_q_limbs in {L, max_level}
 * (Q-limbs first, P-limbs after).  out0,out1: [L][N]. */
void orc_keyswitch(const orc_params *p, uint32_t L, const uint64_t *d, const uint64_t *evk,
                   uint32_t evk_q_limbs, uint64_t *out0, uint64_t *out1);

/* The two halves of the key switch, exposed because the hoisted rotation is defined on them:
 * orc_modup = K1..K4: t [beta][L+alpha][N], the extended digits of d in evaluation form;
 * orc_keyswitch_digits = K5..K10 from given digits.  orc_keyswitch(d) == orc_keyswitch_digits(orc_modup(d)). */
void orc_modup(const orc_params *p, uint32_t L, const uint64_t *d, uint64_t *t);
void orc_keyswitch_digits(const orc_params *p, uint32_t L, const uint64_t *t, const uint64_t *evk, uint32_t evk_q_limbs,
                          uint64_t *out0, uint64_t *out1);
/* Hoisted rotations: n_rot rotations of ONE ciphertext share one ModUp of c1; the automorphism is
 * applied to the extended digits.  Own definition — NOT bit-identical to orc_hrotate (approximate base conversion is not
 * equivariant under the automorphism), equally valid: tests/test_oracle_semantics.py decrypts both.
 * rotkeys / galois / ct_outs: n_rot entries; ct_outs[r] is [2][L][N]. */
void orc_hrotate_hoisted(const orc_params *p, uint32_t L, const uint64_t *ct, uint32_t n_rot, const uint64_t *const *rotkeys,
                         uint32_t evk_q_limbs, const uint64_t *galois, uint64_t *const *ct_outs);

/* Rescale one polynomial: in [L][N] -> out [L-1][N]; non-negative reduction of the dropped limb. */
void orc_rescale(const orc_params *p, uint32_t L, const uint64_t *in, uint64_t *out);

/* ct layout [2][L][N] (c0 limbs then c1 limbs).  out: [2][L-1][N]. */
void orc_hmult(const orc_params *p, uint32_t L, const uint64_t *ct_a, const uint64_t *ct_b,
               const uint64_t *evk, uint32_t evk_q_limbs, uint64_t *ct_out);
/* out: [2][L][N]:  (sigma(c0) + ks0, ks1), ks = KeySwitch(sigma(c1)) */
void orc_hrotate(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *rotkey,
                 uint32_t evk_q_limbs, uint64_t galois_elt, uint64_t *ct_out);
void orc_hadd(const orc_params *p, uint32_t L, const uint64_t *ct_a, const uint64_t *ct_b, uint64_t *ct_out);
void orc_pmult(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *pt, uint64_t *ct_out);
void orc_padd(const orc_params *p, uint32_t L, const uint64_t *ct, const uint64_t *pt, uint64_t *ct_out);

#ifdef __cplusplus
}
#endif
#endif
