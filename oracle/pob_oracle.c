/* oracle/pob_oracle.c -- CPU restatement of the worm-privacy/proof-of-burn circuits as a witness calculator.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/fr.h).  The product path never links, imports or runs this file.
 *
 * What it restates.  The reference "calculator" is the C++ program the external circom compiler emits
 * from the circuits/ .circom sources under --O0 (Makefile:2-3, tests/test.py:32); that program is not in the tree.
 * This file evaluates every template of the include closure of circuits/main_proof_of_burn.circom:27 and
 * circuits/main_spend.circom:6 directly, one C function per circom template, each citing the template it
 * follows, and writes every signal into a flat witness vector using circom's --O0 numbering rules
 * (SURVEY.md Appendix C):
 *   R1  witness[0] = 1; every signal is kept; index = depth-first signal id.
 *   R2  one component = [its own signals: outputs, inputs, intermediates, each in declaration order,
 *       arrays row-major] followed by the complete block of each sub-component.
 *   R3  sub-components are ordered by the moment their last input is assigned ("H-complete", what
 *       circom >= 2.1 does because a sub-template is only instantiated once its input tags are known);
 *       `hcreate=1` switches the two sites where creation order differs (Num2Bits_strict, MultiAND n>=3)
 *       to creation order.
 * PARITY UNPINNED for the whole-witness ORDER: the reference holds no .wtns/.sym/.r1cs golden, and circom
 * cannot run here.  Signal VALUES, output signals and accept/reject are pinned by the reference's own
 * test tables (tests/golden/reference_testcases.json, generated from /root/reference/tests/testcases).
 *
 * Failure semantics: the reference aborts on the first failing `===` (tests/test.py:65-68).  The oracle
 * keeps evaluating and reports status = 1 + (smallest witness index of a component that owns a failing
 * constraint), 0 = accepted -- a definition that does not depend on evaluation order.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include "fr.h"
#include "poseidon_constants_data.h"

typedef struct {
    Fr *w;            /* witness vector (mmap'ed, zero-filled) */
    size_t pos, cap;
    uint64_t status;  /* 0 ok, else 1 + min failing component base */
    int hcreate;
} Ctx;

static inline size_t alloc_sig(Ctx *c, size_t n) {
    size_t b = c->pos; c->pos += n;
    if (c->pos > c->cap) { fprintf(stderr, "pob_oracle: witness capacity exceeded\n"); abort(); }
    return b;
}
static inline void fail_at(Ctx *c, size_t base) { if (c->status == 0 || base + 1 < c->status) c->status = base + 1; }
static inline void check(Ctx *c, int ok, size_t base) { if (!ok) fail_at(c, base); }
#define WW (c->w)
static inline void cpy(Ctx *c, size_t dst, const Fr *src, size_t n) { memcpy(WW + dst, src, n * sizeof(Fr)); }
static Fr *tmp_alloc(size_t n) { Fr *p = (Fr *)calloc(n ? n : 1, sizeof(Fr)); if (!p) abort(); return p; }

/* ============================ circomlib/circuits/gates.circom ============================ */
/* XOR :21-27  out <== a + b - 2*a*b */
static inline size_t T_XOR(Ctx *c, Fr a, Fr b) {
    size_t o = alloc_sig(c, 3); Fr *s = WW + o; s[1] = a; s[2] = b;
    if (fr_fits64(a) && fr_fits64(b) && a.l[0] <= 1 && b.l[0] <= 1) s[0] = fr_u64(a.l[0] ^ b.l[0]);
    else { Fr ab = fr_mul(a, b); s[0] = fr_sub(fr_add(a, b), fr_add(ab, ab)); }
    return o;
}
/* AND :29-35  out <== a*b */
static inline size_t T_AND(Ctx *c, Fr a, Fr b) {
    size_t o = alloc_sig(c, 3); Fr *s = WW + o; s[1] = a; s[2] = b; s[0] = fr_mul(a, b); return o;
}
/* OR :37-43  out <== a + b - a*b */
static inline size_t T_OR(Ctx *c, Fr a, Fr b) {
    size_t o = alloc_sig(c, 3); Fr *s = WW + o; s[1] = a; s[2] = b;
    s[0] = fr_sub(fr_add(a, b), fr_mul(a, b)); return o;
}
/* MultiAND(n) :68-96 */
static size_t T_MultiAND(Ctx *c, int n, const Fr *in) {
    size_t o = alloc_sig(c, 1 + (size_t)n); cpy(c, o + 1, in, (size_t)n);
    if (n == 1) { WW[o] = in[0]; }
    else if (n == 2) { size_t a = T_AND(c, in[0], in[1]); WW[o] = WW[a]; }
    else {
        int n1 = n / 2, n2 = n - n / 2;
        if (c->hcreate) {               /* creation order: and2, ands[0], ands[1] */
            size_t a2 = alloc_sig(c, 3);
            size_t x0 = T_MultiAND(c, n1, in), x1 = T_MultiAND(c, n2, in + n1);
            WW[a2 + 1] = WW[x0]; WW[a2 + 2] = WW[x1]; WW[a2] = fr_mul(WW[x0], WW[x1]); WW[o] = WW[a2];
        } else {                        /* completion order: ands[0], ands[1], and2 */
            size_t x0 = T_MultiAND(c, n1, in), x1 = T_MultiAND(c, n2, in + n1);
            size_t a2 = T_AND(c, WW[x0], WW[x1]); WW[o] = WW[a2];
        }
    }
    return o;
}

/* ============================ circomlib/circuits/bitify.circom ============================ */
/* Num2Bits(n) :25-39: out[i] <-- (in >> i) & 1; out[i]*(out[i]-1) === 0; sum(out[i]*2^i) === in */
static size_t T_Num2Bits(Ctx *c, int n, Fr in) {
    size_t o = alloc_sig(c, (size_t)n + 1); Fr *s = WW + o;
    Fr lc = FR_ZERO;
    for (int i = 0; i < n; i++) {
        int b = fr_bit(in, (unsigned)i); s[i] = fr_u64((uint64_t)b);
        if (b) lc = fr_add(lc, fr_pow2((unsigned)i));
    }
    s[n] = in;
    check(c, fr_eq(lc, in), o);
    return o;
}
/* Bits2Num(n) :55-67 */
static size_t T_Bits2Num(Ctx *c, int n, const Fr *in) {
    size_t o = alloc_sig(c, (size_t)n + 1); cpy(c, o + 1, in, (size_t)n);
    Fr lc = FR_ZERO;
    for (int i = 0; i < n; i++) lc = fr_add(lc, fr_mul(in[i], fr_pow2((unsigned)i)));
    WW[o] = lc; return o;
}
/* compconstant.circom CompConstant(ct) :25-73, instantiated only with ct = -1 = p-1 (aliascheck.circom:28) */
static size_t T_CompConstant(Ctx *c, Fr ct, const Fr *in) {
    size_t o = alloc_sig(c, 1 + 254 + 127 + 1); Fr *s = WW + o;   /* out, in[254], parts[127], sout */
    cpy(c, o + 1, in, 254);
    Fr *parts = s + 255;
    Fr b = fr_sub(fr_pow2(128), FR_ONE), a = FR_ONE, e = FR_ONE, sum = FR_ZERO;
    for (int i = 0; i < 127; i++) {
        int clsb = fr_bit(ct, (unsigned)(2 * i)), cmsb = fr_bit(ct, (unsigned)(2 * i + 1));
        Fr slsb = in[2 * i], smsb = in[2 * i + 1], ml = fr_mul(smsb, slsb), p;
        if (!cmsb && !clsb)      p = fr_add(fr_add(fr_neg(fr_mul(b, ml)), fr_mul(b, smsb)), fr_mul(b, slsb));
        else if (!cmsb && clsb)  p = fr_add(fr_sub(fr_add(fr_sub(fr_mul(a, ml), fr_mul(a, slsb)), fr_mul(b, smsb)), fr_mul(a, smsb)), a);
        else if (cmsb && !clsb)  p = fr_add(fr_sub(fr_mul(b, ml), fr_mul(a, smsb)), a);
        else                     p = fr_add(fr_neg(fr_mul(a, ml)), a);
        parts[i] = p; sum = fr_add(sum, p);
        b = fr_sub(b, e); a = fr_add(a, e); e = fr_add(e, e);
    }
    s[255 + 127] = sum;                                  /* sout */
    size_t nb = T_Num2Bits(c, 135, sum);
    s[0] = WW[nb + 127];
    return o;
}
/* aliascheck.circom AliasCheck :24-32 */
static size_t T_AliasCheck(Ctx *c, const Fr *in) {
    size_t o = alloc_sig(c, 254); cpy(c, o, in, 254);
    Fr m1 = fr_sub(FR_P, FR_ONE);
    size_t cc = T_CompConstant(c, m1, in);
    check(c, fr_is_zero(WW[cc]), o);
    return o;
}
/* Num2Bits_strict :41-53 */
static size_t T_Num2Bits_strict(Ctx *c, Fr in) {
    size_t o = alloc_sig(c, 255); WW[o + 254] = in;
    if (c->hcreate) {                 /* aliasCheck created first */
        Fr *bits = tmp_alloc(254);
        for (int i = 0; i < 254; i++) bits[i] = fr_u64((uint64_t)fr_bit(in, (unsigned)i));
        T_AliasCheck(c, bits);
        size_t nb = T_Num2Bits(c, 254, in);
        cpy(c, o, WW + nb, 254); free(bits);
    } else {                          /* n2b's single input is assigned first => it completes first */
        size_t nb = T_Num2Bits(c, 254, in);
        cpy(c, o, WW + nb, 254);
        T_AliasCheck(c, WW + nb);
    }
    return o;
}

/* ============================ circomlib/circuits/comparators.circom ============================ */
/* IsZero :24-35 */
static inline size_t T_IsZero(Ctx *c, Fr in) {
    size_t o = alloc_sig(c, 3); Fr *s = WW + o; s[1] = in;
    if (fr_is_zero(in)) { s[2] = FR_ZERO; s[0] = FR_ONE; }
    else { s[2] = fr_inv(in); s[0] = fr_add(fr_neg(fr_mul(in, s[2])), FR_ONE); }
    check(c, fr_is_zero(fr_mul(in, s[0])), o);
    return o;
}
/* IsEqual :37-46  isz.in <== in[1] - in[0] */
static inline size_t T_IsEqual(Ctx *c, Fr in0, Fr in1) {
    size_t o = alloc_sig(c, 3); WW[o + 1] = in0; WW[o + 2] = in1;
    size_t z = T_IsZero(c, fr_sub(in1, in0)); WW[o] = WW[z]; return o;
}
/* LessThan(n) :89-100 */
static size_t T_LessThan(Ctx *c, int n, Fr in0, Fr in1) {
    size_t o = alloc_sig(c, 3); WW[o + 1] = in0; WW[o + 2] = in1;
    size_t nb = T_Num2Bits(c, n + 1, fr_sub(fr_add(in0, fr_pow2((unsigned)n)), in1));
    WW[o] = fr_sub(FR_ONE, WW[nb + n]); return o;
}
/* LessEqThan(n) :105-115 */
static size_t T_LessEqThan(Ctx *c, int n, Fr in0, Fr in1) {
    size_t o = alloc_sig(c, 3); WW[o + 1] = in0; WW[o + 2] = in1;
    size_t lt = T_LessThan(c, n, in0, fr_add(in1, FR_ONE)); WW[o] = WW[lt]; return o;
}
/* GreaterEqThan(n) :131-141 */
static size_t T_GreaterEqThan(Ctx *c, int n, Fr in0, Fr in1) {
    size_t o = alloc_sig(c, 3); WW[o + 1] = in0; WW[o + 2] = in1;
    size_t lt = T_LessThan(c, n, in1, fr_add(in0, FR_ONE)); WW[o] = WW[lt]; return o;
}

/* ============================ circomlib/circuits/mux1.circom ============================ */
/* Mux1 :34-48 wrapping MultiMux1(1) :21-32 */
static size_t T_Mux1(Ctx *c, Fr c0, Fr c1, Fr s) {
    size_t o = alloc_sig(c, 4); WW[o + 1] = c0; WW[o + 2] = c1; WW[o + 3] = s;
    size_t m = alloc_sig(c, 4);                          /* MultiMux1(1): out[1], c[1][2], s */
    WW[m + 1] = c0; WW[m + 2] = c1; WW[m + 3] = s;
    WW[m] = fr_add(fr_mul(fr_sub(c1, c0), s), c0);
    WW[o] = WW[m]; return o;
}

/* ============================ circomlib/circuits/poseidon.circom ============================ */
typedef struct { int t, rp; const uint64_t (*C)[4], (*S)[4], (*M)[4], (*P)[4]; } PoseidonK;
static PoseidonK poseidon_k(int t) {
    PoseidonK k; k.t = t;
    if (t == 3) { k.rp = 57; k.C = POSEIDON_C_T3; k.S = POSEIDON_S_T3; k.M = POSEIDON_M_T3; k.P = POSEIDON_P_T3; }
    else if (t == 4) { k.rp = 56; k.C = POSEIDON_C_T4; k.S = POSEIDON_S_T4; k.M = POSEIDON_M_T4; k.P = POSEIDON_P_T4; }
    else if (t == 5) { k.rp = 60; k.C = POSEIDON_C_T5; k.S = POSEIDON_S_T5; k.M = POSEIDON_M_T5; k.P = POSEIDON_P_T5; }
    else { fprintf(stderr, "pob_oracle: Poseidon t=%d not in this circuit's closure\n", t); abort(); }
    return k;
}
static inline Fr K(const uint64_t (*tab)[4], int i) { Fr r; memcpy(&r, tab[i], 32); return r; }
/* Sigma :5-16 */
static size_t T_Sigma(Ctx *c, Fr in) {
    size_t o = alloc_sig(c, 4); Fr *s = WW + o; s[1] = in; s[2] = fr_mul(in, in); s[3] = fr_mul(s[2], s[2]);
    s[0] = fr_mul(s[3], in); return o;
}
/* Ark(t,C,r) :18-25 */
static size_t T_Ark(Ctx *c, PoseidonK k, int r, const Fr *in) {
    int t = k.t; size_t o = alloc_sig(c, 2 * (size_t)t); cpy(c, o + (size_t)t, in, (size_t)t);
    for (int i = 0; i < t; i++) WW[o + (size_t)i] = fr_add(in[i], K(k.C, i + r));
    return o;
}
/* Mix(t,M) :27-39  out[i] = sum_j M[j][i]*in[j] */
static size_t T_Mix(Ctx *c, int t, const uint64_t (*M)[4], const Fr *in) {
    size_t o = alloc_sig(c, 2 * (size_t)t); cpy(c, o + (size_t)t, in, (size_t)t);
    for (int i = 0; i < t; i++) {
        Fr lc = FR_ZERO; for (int j = 0; j < t; j++) lc = fr_add(lc, fr_mul(K(M, j * t + i), in[j]));
        WW[o + (size_t)i] = lc;
    }
    return o;
}
/* MixLast(t,M,s) :41-50 */
static size_t T_MixLast(Ctx *c, int t, const uint64_t (*M)[4], int s, const Fr *in) {
    size_t o = alloc_sig(c, 1 + (size_t)t); cpy(c, o + 1, in, (size_t)t);
    Fr lc = FR_ZERO; for (int j = 0; j < t; j++) lc = fr_add(lc, fr_mul(K(M, j * t + s), in[j]));
    WW[o] = lc; return o;
}
/* MixS(t,S,r) :52-65 */
static size_t T_MixS(Ctx *c, PoseidonK k, int r, const Fr *in) {
    int t = k.t; size_t o = alloc_sig(c, 2 * (size_t)t); cpy(c, o + (size_t)t, in, (size_t)t);
    Fr lc = FR_ZERO; for (int i = 0; i < t; i++) lc = fr_add(lc, fr_mul(K(k.S, (t * 2 - 1) * r + i), in[i]));
    WW[o] = lc;
    for (int i = 1; i < t; i++) WW[o + (size_t)i] = fr_add(in[i], fr_mul(in[0], K(k.S, (t * 2 - 1) * r + t + i - 1)));
    return o;
}
/* PoseidonEx(nInputs, 1) :67-196 */
static size_t T_PoseidonEx(Ctx *c, int nInputs, const Fr *inputs, Fr initialState) {
    int t = nInputs + 1; PoseidonK k = poseidon_k(t);
    size_t o = alloc_sig(c, 1 + (size_t)nInputs + 1);        /* out[1], inputs[n], initialState */
    cpy(c, o + 1, inputs, (size_t)nInputs); WW[o + 1 + (size_t)nInputs] = initialState;
    Fr st[8], nx[8];
    st[0] = initialState; for (int j = 1; j < t; j++) st[j] = inputs[j - 1];
    size_t a = T_Ark(c, k, 0, st); for (int j = 0; j < t; j++) st[j] = WW[a + (size_t)j];     /* ark[0] :92-99 */
    for (int r = 0; r < 3; r++) {                                                            /* :101-121 */
        for (int j = 0; j < t; j++) { size_t s = T_Sigma(c, st[j]); nx[j] = WW[s]; }
        a = T_Ark(c, k, (r + 1) * t, nx); for (int j = 0; j < t; j++) nx[j] = WW[a + (size_t)j];
        size_t m = T_Mix(c, t, k.M, nx); for (int j = 0; j < t; j++) st[j] = WW[m + (size_t)j];
    }
    for (int j = 0; j < t; j++) { size_t s = T_Sigma(c, st[j]); nx[j] = WW[s]; }             /* :123-126 */
    a = T_Ark(c, k, 4 * t, nx); for (int j = 0; j < t; j++) nx[j] = WW[a + (size_t)j];      /* :128-131 */
    { size_t m = T_Mix(c, t, k.P, nx); for (int j = 0; j < t; j++) st[j] = WW[m + (size_t)j]; } /* :133-136 */
    for (int r = 0; r < k.rp; r++) {                                                         /* :138-160 */
        size_t s = T_Sigma(c, st[0]);
        nx[0] = fr_add(WW[s], K(k.C, 5 * t + r)); for (int j = 1; j < t; j++) nx[j] = st[j];
        size_t m = T_MixS(c, k, r, nx); for (int j = 0; j < t; j++) st[j] = WW[m + (size_t)j];
    }
    for (int r = 0; r < 3; r++) {                                                            /* :162-182 */
        for (int j = 0; j < t; j++) { size_t s = T_Sigma(c, st[j]); nx[j] = WW[s]; }
        a = T_Ark(c, k, 5 * t + k.rp + r * t, nx); for (int j = 0; j < t; j++) nx[j] = WW[a + (size_t)j];
        size_t m = T_Mix(c, t, k.M, nx); for (int j = 0; j < t; j++) st[j] = WW[m + (size_t)j];
    }
    for (int j = 0; j < t; j++) { size_t s = T_Sigma(c, st[j]); nx[j] = WW[s]; }             /* :184-187 */
    size_t ml = T_MixLast(c, t, k.M, 0, nx);                                                 /* :189-195 */
    WW[o] = WW[ml]; return o;
}
/* Poseidon(nInputs) :198-208 */
static size_t T_Poseidon(Ctx *c, int n, const Fr *inputs) {
    size_t o = alloc_sig(c, 1 + (size_t)n); cpy(c, o + 1, inputs, (size_t)n);
    size_t e = T_PoseidonEx(c, n, inputs, FR_ZERO); WW[o] = WW[e]; return o;
}

/* ============================ circuits/utils/assert.circom ============================ */
/* AssertBits(B) :13-18 */
static size_t T_AssertBits(Ctx *c, int B, Fr in) {
    size_t o = alloc_sig(c, 1 + (size_t)B); WW[o] = in;
    size_t nb = T_Num2Bits(c, B, in); cpy(c, o + 1, WW + nb, (size_t)B); return o;
}
/* AssertByteString(N) :26-31 */
static size_t T_AssertByteString(Ctx *c, int N, const Fr *in) {
    size_t o = alloc_sig(c, (size_t)N); cpy(c, o, in, (size_t)N);
    for (int i = 0; i < N; i++) T_AssertBits(c, 8, in[i]);
    return o;
}
/* AssertLessThan(B) :40-47, AssertLessEqThan(B) :56-63, AssertGreaterEqThan(B) :72-79 */
static size_t T_AssertCmp(Ctx *c, int kind, int B, Fr a, Fr b) {
    size_t o = alloc_sig(c, 3); WW[o] = a; WW[o + 1] = b;
    T_AssertBits(c, B, a); T_AssertBits(c, B, b);
    size_t r = kind == 0 ? T_LessThan(c, B, a, b) : kind == 1 ? T_LessEqThan(c, B, a, b) : T_GreaterEqThan(c, B, a, b);
    WW[o + 2] = WW[r];
    check(c, fr_eq(WW[r], FR_ONE), o);
    return o;
}
#define T_AssertLessThan(c, B, a, b) T_AssertCmp(c, 0, B, a, b)
#define T_AssertLessEqThan(c, B, a, b) T_AssertCmp(c, 1, B, a, b)
#define T_AssertGreaterEqThan(c, B, a, b) T_AssertCmp(c, 2, B, a, b)

/* ============================ circuits/utils/array.circom ============================ */
/* Filter(N) :26-40 */
static size_t T_Filter(Ctx *c, int N, Fr in) {
    size_t o = alloc_sig(c, 2 * (size_t)N + 1); WW[o + (size_t)N] = in;
    for (int i = 0; i < N; i++) {
        size_t e = T_IsEqual(c, fr_u64((uint64_t)i), in);
        WW[o + (size_t)N + 1 + (size_t)i] = WW[e];
        Fr nf = fr_sub(FR_ONE, WW[e]);
        WW[o + (size_t)i] = i > 0 ? fr_mul(WW[o + (size_t)i - 1], nf) : nf;
    }
    return o;
}
/* Fit(M,N) :47-57 */
static size_t T_Fit(Ctx *c, int M, int N, const Fr *in) {
    size_t o = alloc_sig(c, (size_t)N + (size_t)M); cpy(c, o + (size_t)N, in, (size_t)M);
    for (int i = 0; i < N; i++) WW[o + (size_t)i] = i < M ? in[i] : FR_ZERO;
    return o;
}
/* Flatten(M,N) :64-72 and Reshape(M,N) :79-87 are both the identity on row-major data */
static size_t T_CopyArray(Ctx *c, size_t n, const Fr *in) {
    size_t o = alloc_sig(c, 2 * n); cpy(c, o, in, n); cpy(c, o + n, in, n); return o;
}
/* Reverse(N) :94-99 */
static size_t T_Reverse(Ctx *c, int N, const Fr *in) {
    size_t o = alloc_sig(c, 2 * (size_t)N); cpy(c, o + (size_t)N, in, (size_t)N);
    for (int i = 0; i < N; i++) WW[o + (size_t)i] = in[N - 1 - i];
    return o;
}

/* ============================ circuits/utils/convert.circom ============================ */
/* LittleEndianBytes2Num(N) :12-26 */
static size_t T_LittleEndianBytes2Num(Ctx *c, int N, const Fr *in) {
    size_t o = alloc_sig(c, 1 + (size_t)N); cpy(c, o + 1, in, (size_t)N);
    T_AssertByteString(c, N, in);
    Fr lc = FR_ZERO; for (int i = 0; i < N; i++) lc = fr_add(lc, fr_mul(fr_pow2((unsigned)(8 * i)), in[i]));
    WW[o] = lc; return o;
}
/* BigEndianBytes2Num(N) :33-39 */
static size_t T_BigEndianBytes2Num(Ctx *c, int N, const Fr *in) {
    size_t o = alloc_sig(c, 1 + 2 * (size_t)N); cpy(c, o + 1, in, (size_t)N);
    size_t r = T_Reverse(c, N, in); cpy(c, o + 1 + (size_t)N, WW + r, (size_t)N);
    size_t l = T_LittleEndianBytes2Num(c, N, WW + r); WW[o] = WW[l]; return o;
}
/* Num2BitsSafe(N) :46-56 */
static size_t T_Num2BitsSafe(Ctx *c, int N, Fr in) {
    if (N >= 254) {
        size_t o = alloc_sig(c, (size_t)N + 1 + 254); WW[o + (size_t)N] = in;
        size_t st = T_Num2Bits_strict(c, in); cpy(c, o + (size_t)N + 1, WW + st, 254);
        size_t f = T_Fit(c, 254, N, WW + st); cpy(c, o, WW + f, (size_t)N); return o;
    }
    size_t o = alloc_sig(c, (size_t)N + 1); WW[o + (size_t)N] = in;
    size_t nb = T_Num2Bits(c, N, in); cpy(c, o, WW + nb, (size_t)N); return o;
}
/* Num2LittleEndianBytes(N) :69-82 */
static size_t T_Num2LittleEndianBytes(Ctx *c, int N, Fr in) {
    size_t n = (size_t)N, o = alloc_sig(c, n + 1 + 8 * n + 8 * n); WW[o + n] = in;
    size_t b = T_Num2BitsSafe(c, 8 * N, in); cpy(c, o + n + 1, WW + b, 8 * n);
    size_t r = T_CopyArray(c, 8 * n, WW + b); cpy(c, o + n + 1 + 8 * n, WW + r, 8 * n);   /* Reshape(N,8) */
    for (int i = 0; i < N; i++) { size_t bn = T_Bits2Num(c, 8, WW + r + 8 * (size_t)i); WW[o + (size_t)i] = WW[bn]; }
    return o;
}
/* Num2BigEndianBytes(N) :90-96 */
static size_t T_Num2BigEndianBytes(Ctx *c, int N, Fr in) {
    size_t n = (size_t)N, o = alloc_sig(c, 2 * n + 1); WW[o + n] = in;
    size_t le = T_Num2LittleEndianBytes(c, N, in); cpy(c, o + n + 1, WW + le, n);
    size_t rv = T_Reverse(c, N, WW + le); cpy(c, o, WW + rv, n); return o;
}
/* Bytes2Nibbles(N) :103-125 */
static size_t T_Bytes2Nibbles(Ctx *c, int N, const Fr *in) {
    size_t n = (size_t)N, o = alloc_sig(c, 2 * n + n + 8 * n); cpy(c, o + 2 * n, in, n);
    for (int i = 0; i < N; i++) {
        size_t nb = T_Num2Bits(c, 8, in[i]); cpy(c, o + 3 * n + 8 * (size_t)i, WW + nb, 8);
        Fr lo = FR_ZERO, hi = FR_ZERO;
        for (int j = 0; j < 4; j++) {
            lo = fr_add(lo, fr_mul(WW[nb + (size_t)j], fr_pow2((unsigned)j)));
            hi = fr_add(hi, fr_mul(WW[nb + (size_t)j + 4], fr_pow2((unsigned)j)));
        }
        WW[o + 2 * (size_t)i] = hi; WW[o + 2 * (size_t)i + 1] = lo;
    }
    return o;
}
/* Nibbles2Bytes(n) :132-141 */
static size_t T_Nibbles2Bytes(Ctx *c, int n, const Fr *nib) {
    size_t o = alloc_sig(c, 3 * (size_t)n); cpy(c, o + (size_t)n, nib, 2 * (size_t)n);
    for (int i = 0; i < n; i++) {
        T_AssertBits(c, 4, nib[2 * i]); T_AssertBits(c, 4, nib[2 * i + 1]);
        WW[o + (size_t)i] = fr_add(fr_mul(nib[2 * i], fr_u64(16)), nib[2 * i + 1]);
    }
    return o;
}

/* ============================ circuits/utils/divide.circom ============================ */
/* Divide(N) :17-33 : out <-- a \ b; rem <-- a % b */
static size_t T_Divide(Ctx *c, int N, Fr a, Fr b) {
    size_t o = alloc_sig(c, 4); Fr q = FR_ZERO, r = FR_ZERO;
    if (fr_is_zero(b)) fail_at(c, o); else fr_divmod(a, b, &q, &r);
    WW[o] = q; WW[o + 1] = r; WW[o + 2] = a; WW[o + 3] = b;
    T_AssertLessThan(c, N, r, b);
    T_AssertLessEqThan(c, N, q, a);
    check(c, fr_eq(fr_add(fr_mul(q, b), r), a), o);
    return o;
}

/* ============================ circuits/utils/selector.circom ============================ */
/* Selector(n) :21-46 */
static size_t T_Selector(Ctx *c, int n, const Fr *vals, Fr select) {
    size_t N = (size_t)n, o = alloc_sig(c, 1 + N + 1 + N + N + 1);
    cpy(c, o + 1, vals, N); WW[o + 1 + N] = select;
    size_t isEq = o + 2 + N, sum = isEq + N;
    Fr cnt = FR_ZERO; WW[sum] = FR_ZERO;
    for (int i = 0; i < n; i++) {
        size_t e = T_IsEqual(c, select, fr_u64((uint64_t)i));
        WW[isEq + (size_t)i] = WW[e]; cnt = fr_add(cnt, WW[e]);
        WW[sum + (size_t)i + 1] = fr_add(WW[sum + (size_t)i], fr_mul(WW[e], vals[i]));
    }
    check(c, fr_eq(cnt, FR_ONE), o);
    WW[o] = WW[sum + N]; return o;
}
/* SelectorArray1D(n,p) :62-77 ; SelectorArray2D(n,p,q) :91-110 (same layout with p*q columns) */
static size_t T_SelectorArray(Ctx *c, int n, size_t cols, const Fr *arrays, Fr select) {
    size_t N = (size_t)n, o = alloc_sig(c, cols + N * cols + 1 + cols * N);
    cpy(c, o + cols, arrays, N * cols); WW[o + cols + N * cols] = select;
    size_t T = o + cols + N * cols + 1;
    for (size_t i = 0; i < N; i++) for (size_t j = 0; j < cols; j++) WW[T + j * N + i] = arrays[i * cols + j];
    for (size_t j = 0; j < cols; j++) { size_t s = T_Selector(c, n, WW + T + j * N, select); WW[o + j] = WW[s]; }
    return o;
}

/* ============================ circuits/utils/shift.circom ============================ */
/* ShiftLeft(n) :17-36 */
static size_t T_ShiftLeft(Ctx *c, int n, const Fr *in, Fr count) {
    size_t N = (size_t)n, o = alloc_sig(c, 2 * N + 1 + 2 * N * N);
    cpy(c, o + N, in, N); WW[o + 2 * N] = count;
    size_t isEq = o + 2 * N + 1, temp = isEq + N * N;
    T_AssertLessEqThan(c, 16, count, fr_u64((uint64_t)n));
    for (int i = 0; i < n; i++) {
        Fr acc = FR_ZERO;
        for (int j = 0; j < n; j++) {
            size_t e = T_IsEqual(c, fr_u64((uint64_t)i), fr_sub(fr_u64((uint64_t)j), count));
            WW[isEq + (size_t)i * N + (size_t)j] = WW[e];
            Fr tv = fr_mul(WW[e], in[j]); WW[temp + (size_t)i * N + (size_t)j] = tv; acc = fr_add(acc, tv);
        }
        WW[o + (size_t)i] = acc;
    }
    return o;
}
/* ShiftRight(n, maxShift) :51-75 */
static size_t T_ShiftRight(Ctx *c, int n, int maxShift, const Fr *in, Fr count) {
    size_t N = (size_t)n, MS = (size_t)maxShift, o = alloc_sig(c, N + MS + N + 1 + MS + 1 + (MS + 1) * N);
    cpy(c, o + N + MS, in, N); WW[o + 2 * N + MS] = count;
    size_t isEq = o + 2 * N + MS + 1, temps = isEq + MS + 1;
    T_AssertLessEqThan(c, 16, count, fr_u64((uint64_t)maxShift));
    Fr *acc = tmp_alloc(N + MS);
    for (int i = 0; i <= maxShift; i++) {
        size_t e = T_IsEqual(c, fr_u64((uint64_t)i), count); WW[isEq + (size_t)i] = WW[e];
        for (int j = 0; j < n; j++) {
            Fr tv = fr_mul(WW[e], in[j]); WW[temps + (size_t)i * N + (size_t)j] = tv;
            acc[i + j] = fr_add(acc[i + j], tv);
        }
    }
    cpy(c, o, acc, N + MS); free(acc); return o;
}

/* ============================ circuits/utils/concat.circom ============================ */
/* Mask(n) :18-30 */
static size_t T_Mask(Ctx *c, int n, const Fr *in, Fr count) {
    size_t N = (size_t)n, o = alloc_sig(c, 3 * N + 1); cpy(c, o + N, in, N); WW[o + 2 * N] = count;
    size_t f = T_Filter(c, n, count); cpy(c, o + 2 * N + 1, WW + f, N);
    for (size_t i = 0; i < N; i++) WW[o + i] = fr_mul(in[i], WW[f + i]);
    return o;
}
/* Concat(maxLenA, maxLenB) :47-84 */
static size_t T_Concat(Ctx *c, int A, int B, const Fr *a, Fr aLen, const Fr *b, Fr bLen) {
    size_t NA = (size_t)A, NB = (size_t)B, T = NA + NB;
    size_t o = alloc_sig(c, T + 1 + NA + 1 + NB + 1 + NA + NB + T);
    size_t ia = o + T + 1, iaL = ia + NA, ib = iaL + 1, ibL = ib + NB, mA = ibL + 1, mB = mA + NA, sB = mB + NB;
    cpy(c, ia, a, NA); WW[iaL] = aLen; cpy(c, ib, b, NB); WW[ibL] = bLen;
    T_AssertLessEqThan(c, 16, aLen, fr_u64((uint64_t)A));
    T_AssertLessEqThan(c, 16, bLen, fr_u64((uint64_t)B));
    size_t ma = T_Mask(c, A, a, aLen); cpy(c, mA, WW + ma, NA);
    size_t mb = T_Mask(c, B, b, bLen); cpy(c, mB, WW + mb, NB);
    size_t sh = T_ShiftRight(c, B, A, WW + mb, aLen); cpy(c, sB, WW + sh, T);
    for (size_t i = 0; i < T; i++) WW[o + i] = i < NA ? fr_add(WW[ma + i], WW[sh + i]) : WW[sh + i];
    WW[o + T] = fr_add(aLen, bLen);
    return o;
}

/* ============================ circuits/utils/substring_check.circom ============================ */
/* SubstringCheck(maxMainLen, subLen) :24-100 */
static size_t T_SubstringCheck(Ctx *c, int maxMainLen, int subLen, const Fr *mainInput, Fr mainLen, const Fr *subInput) {
    size_t MM = (size_t)maxMainLen, SL = (size_t)subLen, Kn = MM - SL + 1;
    size_t o = alloc_sig(c, 1 + MM + 1 + SL + 1 + (MM + 1) + Kn + Kn + (Kn + 1) + (Kn + 1) + 1);
    size_t iMain = o + 1, iLen = iMain + MM, iSub = iLen + 1, subNum = iSub + SL, Mo = subNum + 1,
           exists = Mo + MM + 1, isLast = exists + Kn, allowed = isLast + Kn, sums = allowed + Kn + 1, dne = sums + Kn + 1;
    cpy(c, iMain, mainInput, MM); WW[iLen] = mainLen; cpy(c, iSub, subInput, SL);
    T_AssertByteString(c, subLen, subInput);
    T_AssertByteString(c, maxMainLen, mainInput);
    T_AssertLessEqThan(c, 16, mainLen, fr_u64((uint64_t)maxMainLen));
    T_AssertLessEqThan(c, 16, fr_u64((uint64_t)subLen), mainLen);
    size_t sn = T_LittleEndianBytes2Num(c, subLen, subInput); WW[subNum] = WW[sn];
    WW[Mo] = FR_ZERO;
    Fr pw = FR_ONE, c256 = fr_u64(256);
    for (size_t i = 0; i < MM; i++) { WW[Mo + i + 1] = fr_add(fr_mul(mainInput[i], pw), WW[Mo + i]); pw = fr_mul(pw, c256); }
    WW[allowed] = FR_ONE; WW[sums] = FR_ZERO;
    pw = FR_ONE;
    Fr lastIdx = fr_add(fr_sub(mainLen, fr_u64((uint64_t)subLen)), FR_ONE);
    for (size_t i = 0; i < Kn; i++) {
        size_t e1 = T_IsEqual(c, fr_u64((uint64_t)i), lastIdx); WW[isLast + i] = WW[e1];
        WW[allowed + i + 1] = fr_mul(WW[allowed + i], fr_sub(FR_ONE, WW[e1]));
        size_t e2 = T_IsEqual(c, fr_mul(WW[subNum], pw), fr_sub(WW[Mo + i + SL], WW[Mo + i])); WW[exists + i] = WW[e2];
        WW[sums + i + 1] = fr_add(WW[sums + i], fr_mul(WW[allowed + i + 1], WW[e2]));
        pw = fr_mul(pw, c256);
    }
    size_t z = T_IsZero(c, WW[sums + Kn]); WW[dne] = WW[z];
    WW[o] = fr_sub(FR_ONE, WW[z]);
    return o;
}

/* ============================ circuits/utils/keccak.circom ============================ */
/* ShR(n,r) :19-31 / ShL(n,r) :40-51 */
static size_t T_ShR(Ctx *c, int n, int r, const Fr *in) {
    size_t o = alloc_sig(c, 2 * (size_t)n); cpy(c, o + (size_t)n, in, (size_t)n);
    for (int i = 0; i < n; i++) WW[o + (size_t)i] = (i + r >= n) ? FR_ZERO : in[i + r];
    return o;
}
static size_t T_ShL(Ctx *c, int n, int r, const Fr *in) {
    size_t o = alloc_sig(c, 2 * (size_t)n); cpy(c, o + (size_t)n, in, (size_t)n);
    for (int i = 0; i < n; i++) WW[o + (size_t)i] = (i < r) ? FR_ZERO : in[i - r];
    return o;
}
/* XorArray(n) :77-85, OrArray(n) :105-113, AndArray(n) :120-128 */
static size_t T_GateArray(Ctx *c, int gate, int n, const Fr *a, const Fr *b) {
    size_t N = (size_t)n, o = alloc_sig(c, 3 * N); cpy(c, o + N, a, N); cpy(c, o + 2 * N, b, N);
    for (size_t i = 0; i < N; i++) {
        size_t g = gate == 0 ? T_XOR(c, a[i], b[i]) : gate == 1 ? T_OR(c, a[i], b[i]) : T_AND(c, a[i], b[i]);
        WW[o + i] = WW[g];
    }
    return o;
}
#define T_XorArray(c, n, a, b) T_GateArray(c, 0, n, a, b)
#define T_OrArray(c, n, a, b) T_GateArray(c, 1, n, a, b)
#define T_AndArray(c, n, a, b) T_GateArray(c, 2, n, a, b)
/* NotArray(n) :92-98 */
static size_t T_NotArray(Ctx *c, int n, const Fr *a) {
    size_t N = (size_t)n, o = alloc_sig(c, 2 * N); cpy(c, o + N, a, N);
    for (size_t i = 0; i < N; i++) WW[o + i] = fr_sub(FR_ONE, a[i]);
    return o;
}
/* Xor5(n) :58-70 */
static size_t T_Xor5(Ctx *c, int n, const Fr *a, const Fr *b, const Fr *cc, const Fr *d, const Fr *e) {
    size_t N = (size_t)n, o = alloc_sig(c, 9 * N);
    cpy(c, o + N, a, N); cpy(c, o + 2 * N, b, N); cpy(c, o + 3 * N, cc, N); cpy(c, o + 4 * N, d, N); cpy(c, o + 5 * N, e, N);
    size_t x = T_XorArray(c, n, a, b); cpy(c, o + 6 * N, WW + x, N);
    x = T_XorArray(c, n, WW + o + 6 * N, cc); cpy(c, o + 7 * N, WW + x, N);
    x = T_XorArray(c, n, WW + o + 7 * N, d); cpy(c, o + 8 * N, WW + x, N);
    x = T_XorArray(c, n, WW + o + 8 * N, e); cpy(c, o, WW + x, N);
    return o;
}
/* D :135-144 */
static size_t T_D(Ctx *c, const Fr *a, const Fr *b) {
    size_t o = alloc_sig(c, 6 * 64); cpy(c, o + 64, a, 64); cpy(c, o + 128, b, 64);
    size_t s0 = T_ShL(c, 64, 1, a); cpy(c, o + 192, WW + s0, 64);
    size_t s1 = T_ShR(c, 64, 63, a); cpy(c, o + 256, WW + s1, 64);
    size_t r = T_OrArray(c, 64, WW + o + 192, WW + o + 256); cpy(c, o + 320, WW + r, 64);
    size_t x = T_XorArray(c, 64, b, WW + o + 320); cpy(c, o, WW + x, 64);
    return o;
}
/* Theta :151-170 */
static size_t T_Theta(Ctx *c, const Fr *in) {
    size_t o = alloc_sig(c, 1600 + 1600 + 320 + 320); cpy(c, o + 1600, in, 1600);
    size_t cs = o + 3200, ds = o + 3520;
    for (int i = 0; i < 5; i++) {
        size_t x = T_Xor5(c, 64, in + 64 * i, in + 64 * (5 + i), in + 64 * (10 + i), in + 64 * (15 + i), in + 64 * (20 + i));
        cpy(c, cs + 64 * (size_t)i, WW + x, 64);
    }
    for (int i = 0; i < 5; i++) {
        size_t d = T_D(c, WW + cs + 64 * (size_t)((i + 1) % 5), WW + cs + 64 * (size_t)((i + 4) % 5));
        cpy(c, ds + 64 * (size_t)i, WW + d, 64);
    }
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
        size_t x = T_XorArray(c, 64, in + 64 * (i + j * 5), WW + ds + 64 * (size_t)i);
        cpy(c, o + 64 * (size_t)(i + j * 5), WW + x, 64);
    }
    return o;
}
/* stepRhoPi(shl, shr) :177-184 */
static size_t T_stepRhoPi(Ctx *c, int shl, int shr, const Fr *a) {
    size_t o = alloc_sig(c, 4 * 64); cpy(c, o + 64, a, 64);
    size_t s0 = T_ShR(c, 64, shr, a); cpy(c, o + 128, WW + s0, 64);
    size_t s1 = T_ShL(c, 64, shl, a); cpy(c, o + 192, WW + s1, 64);
    size_t r = T_OrArray(c, 64, WW + o + 128, WW + o + 192); cpy(c, o, WW + r, 64);
    return o;
}
/* RhoPi :191-204 */
static size_t T_RhoPi(Ctx *c, const Fr *in) {
    static const int rot[25] = {1, 10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    size_t o = alloc_sig(c, 3200); cpy(c, o + 1600, in, 1600);
    cpy(c, o, in, 64);
    for (int i = 0; i < 24; i++) {
        int shl = ((i + 1) * (i + 2) / 2) % 64;
        size_t s = T_stepRhoPi(c, shl, 64 - shl, in + 64 * rot[i]);
        cpy(c, o + 64 * (size_t)rot[i + 1], WW + s, 64);
    }
    return o;
}
/* stepChi :212-221 */
static size_t T_stepChi(Ctx *c, const Fr *a, const Fr *b, const Fr *cc) {
    size_t o = alloc_sig(c, 6 * 64); cpy(c, o + 64, a, 64); cpy(c, o + 128, b, 64); cpy(c, o + 192, cc, 64);
    size_t n = T_NotArray(c, 64, b); cpy(c, o + 256, WW + n, 64);
    size_t an = T_AndArray(c, 64, WW + o + 256, cc); cpy(c, o + 320, WW + an, 64);
    size_t x = T_XorArray(c, 64, a, WW + o + 320); cpy(c, o, WW + x, 64);
    return o;
}
/* Chi :228-241 */
static size_t T_Chi(Ctx *c, const Fr *in) {
    size_t o = alloc_sig(c, 3200); cpy(c, o + 1600, in, 1600);
    for (int i = 0; i < 25; i++) {
        size_t s;
        if (i % 5 == 3) s = T_stepChi(c, in + 64 * i, in + 64 * (i + 1), in + 64 * (i - 3));
        else if (i % 5 == 4) s = T_stepChi(c, in + 64 * i, in + 64 * (i - 4), in + 64 * (i - 3));
        else s = T_stepChi(c, in + 64 * i, in + 64 * (i + 1), in + 64 * (i + 2));
        cpy(c, o + 64 * (size_t)i, WW + s, 64);
    }
    return o;
}
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
/* RoundConstants(r) :248-266 */
static size_t T_RoundConstants(Ctx *c, int r) {
    size_t o = alloc_sig(c, 64);
    for (int i = 0; i < 64; i++) WW[o + (size_t)i] = fr_u64((KECCAK_RC[r] >> i) & 1);
    return o;
}
/* Iota(r) :273-283 */
static size_t T_Iota(Ctx *c, int r, const Fr *in) {
    size_t o = alloc_sig(c, 3200 + 64); cpy(c, o + 1600, in, 1600);
    size_t rc = T_RoundConstants(c, r); cpy(c, o + 3200, WW + rc, 64);
    size_t x = T_XorArray(c, 64, in, WW + o + 3200);
    cpy(c, o, WW + x, 64); cpy(c, o + 64, in + 64, 1600 - 64);
    return o;
}
/* KeccakfRound(r) :290-297 */
static size_t T_KeccakfRound(Ctx *c, int r, const Fr *in) {
    size_t o = alloc_sig(c, 5 * 1600); cpy(c, o + 1600, in, 1600);
    size_t t = T_Theta(c, in); cpy(c, o + 3200, WW + t, 1600);
    size_t p = T_RhoPi(c, WW + o + 3200); cpy(c, o + 4800, WW + p, 1600);
    size_t x = T_Chi(c, WW + o + 4800); cpy(c, o + 6400, WW + x, 1600);
    size_t i = T_Iota(c, r, WW + o + 6400); cpy(c, o, WW + i, 1600);
    return o;
}
/* Keccakf :356-367 */
static size_t T_Keccakf(Ctx *c, const Fr *in) {
    size_t o = alloc_sig(c, 1600 + 1600 + 25 * 1600); cpy(c, o + 1600, in, 1600);
    size_t mid = o + 3200; cpy(c, mid, in, 1600);
    for (int i = 0; i < 24; i++) {
        size_t r = T_KeccakfRound(c, i, WW + mid + 1600 * (size_t)i);
        cpy(c, mid + 1600 * (size_t)(i + 1), WW + r, 1600);
    }
    cpy(c, o, WW + mid + 1600 * 24, 1600);
    return o;
}
/* Absorb :304-323 */
static size_t T_Absorb(Ctx *c, const Fr *s, const Fr *block) {
    size_t o = alloc_sig(c, 1600 + 1600 + 1088 + 1600); cpy(c, o + 1600, s, 1600); cpy(c, o + 3200, block, 1088);
    size_t aux = o + 4288;
    for (int i = 0; i < 25; i++) {
        if (i < 17) { size_t x = T_XorArray(c, 64, s + 64 * i, block + 64 * i); cpy(c, aux + 64 * (size_t)i, WW + x, 64); }
        else cpy(c, aux + 64 * (size_t)i, s + 64 * i, 64);
    }
    size_t f = T_Keccakf(c, WW + aux); cpy(c, o, WW + f, 1600);
    return o;
}
/* Final(nBlocksIn) :330-349 */
static size_t T_Final(Ctx *c, int n, const Fr *in, Fr blocks) {
    size_t N = (size_t)n, o = alloc_sig(c, 1600 + N * 1088 + 1 + (N + 1) * 1600);
    cpy(c, o + 1600, in, N * 1088); WW[o + 1600 + N * 1088] = blocks;
    size_t s = o + 1600 + N * 1088 + 1;                       /* s[0] is all zero already */
    for (size_t b = 0; b < N; b++) {
        size_t a = T_Absorb(c, WW + s + 1600 * b, in + 1088 * b); cpy(c, s + 1600 * (b + 1), WW + a, 1600);
    }
    size_t sel = T_SelectorArray(c, n + 1, 1600, WW + s, blocks); cpy(c, o, WW + sel, 1600);
    return o;
}
/* Keccak(nBlocksIn) :374-385 */
static size_t T_Keccak(Ctx *c, int n, const Fr *in, Fr blocks) {
    size_t N = (size_t)n, o = alloc_sig(c, 256 + N * 1088 + 1 + 1600);
    cpy(c, o + 256, in, N * 1088); WW[o + 256 + N * 1088] = blocks;
    size_t f = T_Final(c, n, in, blocks); cpy(c, o + 256 + N * 1088 + 1, WW + f, 1600);
    cpy(c, o, WW + f, 256);
    return o;
}
/* Pad(maxBlocks, blockSize) :412-446 */
static size_t T_Pad(Ctx *c, int maxBlocks, int blockSize, const Fr *in, Fr inLen) {
    size_t B = (size_t)maxBlocks * (size_t)blockSize, o = alloc_sig(c, B + 1 + B + 1 + 2 + (B + 1) + B + B);
    size_t numBlocks = o + B, iIn = numBlocks + 1, iLen = iIn + B, div = iLen + 1, rem = div + 1,
           filter = rem + 1, isEq = filter + B + 1, isLast = isEq + B;
    cpy(c, iIn, in, B); WW[iLen] = inLen;
    size_t d = T_Divide(c, 16, inLen, fr_u64((uint64_t)blockSize)); WW[div] = WW[d]; WW[rem] = WW[d + 1];
    WW[numBlocks] = fr_add(WW[div], FR_ONE);
    T_AssertLessEqThan(c, 16, WW[numBlocks], fr_u64((uint64_t)maxBlocks));
    WW[filter] = FR_ONE;
    for (size_t i = 0; i < B; i++) {
        size_t e = T_IsEqual(c, fr_u64(i), inLen); WW[isEq + i] = WW[e];
        WW[filter + i + 1] = fr_mul(WW[filter + i], fr_sub(FR_ONE, WW[e]));
    }
    Fr lastPos = fr_sub(fr_mul(WW[numBlocks], fr_u64((uint64_t)blockSize)), FR_ONE);
    for (size_t i = 0; i < B; i++) {
        size_t e = T_IsEqual(c, fr_u64(i), lastPos); WW[isLast + i] = WW[e];
        WW[o + i] = fr_add(fr_add(fr_mul(in[i], WW[filter + i + 1]), WW[isEq + i]), fr_mul(fr_u64(0x80), WW[e]));
    }
    return o;
}
/* KeccakBytes(maxBlocks) :454-489 */
static size_t T_KeccakBytes(Ctx *c, int maxBlocks, const Fr *in, Fr inLen) {
    size_t B = (size_t)maxBlocks * 136;
    size_t o = alloc_sig(c, 32 + B + 1 + B + 1 + 8 * B + 8 * B + 8 * B + 256 + 256);
    size_t iIn = o + 32, iLen = iIn + B, padded = iLen + 1, numBlocks = padded + B, inBitsArray = numBlocks + 1,
           inBits = inBitsArray + 8 * B, inBlocks = inBits + 8 * B, outBits = inBlocks + 8 * B, outBytes = outBits + 256;
    cpy(c, iIn, in, B); WW[iLen] = inLen;
    T_AssertLessThan(c, 16, inLen, fr_u64(B));
    size_t p = T_Pad(c, maxBlocks, 136, in, inLen); cpy(c, padded, WW + p, B); WW[numBlocks] = WW[p + B];
    for (size_t i = 0; i < B; i++) { size_t nb = T_Num2Bits(c, 8, WW[padded + i]); cpy(c, inBitsArray + 8 * i, WW + nb, 8); }
    size_t fl = T_CopyArray(c, 8 * B, WW + inBitsArray); cpy(c, inBits, WW + fl, 8 * B);    /* Flatten(B,8) */
    cpy(c, inBlocks, WW + inBits, 8 * B);                                                    /* same row-major order */
    size_t k = T_Keccak(c, maxBlocks, WW + inBlocks, WW[numBlocks]); cpy(c, outBits, WW + k, 256);
    size_t rs = T_CopyArray(c, 256, WW + outBits); cpy(c, outBytes, WW + rs, 256);          /* Reshape(32,8) */
    for (size_t i = 0; i < 32; i++) { size_t bn = T_Bits2Num(c, 8, WW + outBytes + 8 * i); WW[o + i] = WW[bn]; }
    return o;
}

/* ============================ circuits/utils/public_commitment.circom ============================ */
/* PublicCommitment(N) :18-42 */
static size_t T_PublicCommitment(Ctx *c, int N, const Fr *in) {
    size_t n32 = (size_t)N * 32; int nb = (N * 32) / 136 + ((N * 32) % 136 != 0); size_t blk = (size_t)nb * 136;
    size_t o = alloc_sig(c, 1 + n32 + n32 + blk + 32 + 31);
    size_t iIn = o + 1, flat = iIn + n32, block = flat + n32, hash = block + blk, red = hash + 32;
    cpy(c, iIn, in, n32);
    for (int i = 0; i < N; i++) T_AssertByteString(c, 32, in + 32 * i);
    size_t f = T_CopyArray(c, n32, in); cpy(c, flat, WW + f, n32);                             /* Flatten(N,32) */
    size_t ft = T_Fit(c, (int)n32, (int)blk, WW + flat); cpy(c, block, WW + ft, blk);
    size_t k = T_KeccakBytes(c, nb, WW + block, fr_u64(n32)); cpy(c, hash, WW + k, 32);
    size_t f2 = T_Fit(c, 32, 31, WW + hash); cpy(c, red, WW + f2, 31);
    size_t be = T_BigEndianBytes2Num(c, 31, WW + red); WW[o] = WW[be];
    return o;
}

/* ============================ circuits/utils/constants.circom ============================ */
static Fr POSEIDON_PREFIX(int add) {   /* :3-15 keccak("EIP-7503") mod p, +0 address / +1 nullifier / +2 coin */
    /* 5265656504298861414514317065875120428884240036965045859626767452974705356670 */
    Fr r = {{0xf0363f983d892f7eULL, 0xd115b780980a6b46ULL, 0x007d2482cd46cec2ULL, 0x0ba44186ee7876b8ULL}};
    return fr_add(r, fr_u64((uint64_t)add));
}

/* ============================ circuits/utils/burn_address.circom ============================ */
/* BurnAddress :47-58 */
static size_t T_BurnAddress(Ctx *c, Fr burnKey, Fr revealAmount, Fr burnExtraCommitment) {
    size_t o = alloc_sig(c, 20 + 3 + 1 + 32);
    WW[o + 20] = burnKey; WW[o + 21] = revealAmount; WW[o + 22] = burnExtraCommitment;
    Fr ins[4] = {POSEIDON_PREFIX(0), burnKey, revealAmount, burnExtraCommitment};
    size_t p = T_Poseidon(c, 4, ins); WW[o + 23] = WW[p];
    size_t b = T_Num2BigEndianBytes(c, 32, WW[p]); cpy(c, o + 24, WW + b, 32);
    size_t f = T_Fit(c, 32, 20, WW + o + 24); cpy(c, o, WW + f, 20);
    return o;
}
/* BurnAddressHash :67-83 */
static size_t T_BurnAddressHash(Ctx *c, Fr burnKey, Fr revealAmount, Fr burnExtraCommitment) {
    size_t o = alloc_sig(c, 64 + 3 + 20 + 136 + 32);
    WW[o + 64] = burnKey; WW[o + 65] = revealAmount; WW[o + 66] = burnExtraCommitment;
    size_t a = T_BurnAddress(c, burnKey, revealAmount, burnExtraCommitment); cpy(c, o + 67, WW + a, 20);
    size_t f = T_Fit(c, 20, 136, WW + o + 67); cpy(c, o + 87, WW + f, 136);
    size_t k = T_KeccakBytes(c, 1, WW + o + 87, fr_u64(20)); cpy(c, o + 223, WW + k, 32);
    size_t nb = T_Bytes2Nibbles(c, 32, WW + o + 223); cpy(c, o, WW + nb, 64);
    return o;
}

/* ============================ circuits/utils/proof_of_work.circom ============================ */
/* EIP7503 :11-21 */
static size_t T_EIP7503(Ctx *c) {
    static const uint8_t s[8] = {69, 73, 80, 45, 55, 53, 48, 51};
    size_t o = alloc_sig(c, 8); for (int i = 0; i < 8; i++) WW[o + (size_t)i] = fr_u64(s[i]); return o;
}
/* ConcatFixed4(A,B,C,D) :28-48 */
static size_t T_ConcatFixed4(Ctx *c, int A, int B, int C, int D, const Fr *a, const Fr *b, const Fr *cc, const Fr *d) {
    size_t T = (size_t)(A + B + C + D), o = alloc_sig(c, 2 * T);
    cpy(c, o, a, (size_t)A); cpy(c, o + (size_t)A, b, (size_t)B); cpy(c, o + (size_t)(A + B), cc, (size_t)C); cpy(c, o + (size_t)(A + B + C), d, (size_t)D);
    cpy(c, o + T, WW + o, T);      /* inputs a,b,c,d in declaration order == the same concatenation */
    return o;
}
/* ProofOfWorkChecker :54-81 */
static size_t T_ProofOfWorkChecker(Ctx *c, Fr burnKey, Fr revealAmount, Fr burnExtraCommitment, Fr minimumZeroBytes) {
    size_t o = alloc_sig(c, 4 + 32 * 3 + 8 + 104 + 136 + 32 + 32);
    size_t bk = o + 4, ra = bk + 32, be = ra + 32, eip = be + 32, hin = eip + 8, blk = hin + 104, kec = blk + 136, sbz = kec + 32;
    WW[o] = burnKey; WW[o + 1] = revealAmount; WW[o + 2] = burnExtraCommitment; WW[o + 3] = minimumZeroBytes;
    size_t x = T_Num2BigEndianBytes(c, 32, burnKey); cpy(c, bk, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, revealAmount); cpy(c, ra, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, burnExtraCommitment); cpy(c, be, WW + x, 32);
    x = T_EIP7503(c); cpy(c, eip, WW + x, 8);
    x = T_ConcatFixed4(c, 32, 32, 32, 8, WW + bk, WW + ra, WW + be, WW + eip); cpy(c, hin, WW + x, 104);
    x = T_Fit(c, 104, 136, WW + hin); cpy(c, blk, WW + x, 136);
    x = T_KeccakBytes(c, 1, WW + blk, fr_u64(104)); cpy(c, kec, WW + x, 32);
    x = T_Filter(c, 32, minimumZeroBytes); cpy(c, sbz, WW + x, 32);
    for (size_t i = 0; i < 32; i++) check(c, fr_is_zero(fr_mul(WW[kec + i], WW[sbz + i])), o);
    return o;
}

/* ============================ circuits/utils/rlp/integer.circom ============================ */
/* CountBytes(N) :16-49 */
static size_t T_CountBytes(Ctx *c, int N, const Fr *bytes) {
    size_t n = (size_t)N, o = alloc_sig(c, 1 + 3 * n); cpy(c, o + 1, bytes, n);
    for (size_t i = 0; i < n; i++) { size_t z = T_IsZero(c, bytes[i]); WW[o + 1 + n + i] = WW[z]; }
    Fr lead = FR_ZERO;
    for (size_t i = 0; i < n; i++) {
        WW[o + 1 + 2 * n + i] = i == 0 ? WW[o + 1 + n] : fr_mul(WW[o + 1 + n + i], WW[o + 1 + 2 * n + i - 1]);
        lead = fr_add(lead, WW[o + 1 + 2 * n + i]);
    }
    WW[o] = fr_sub(fr_u64(n), lead); return o;
}
/* RlpInteger(N) :67-110 */
static size_t T_RlpInteger(Ctx *c, int N, Fr in) {
    size_t n = (size_t)N, o = alloc_sig(c, n + 1 + 1 + 1 + n + 1 + n + 3);
    size_t outLen = o + n + 1, iIn = outLen + 1, bytes = iIn + 1, length = bytes + n, bigEndian = length + 1,
           isSingle = bigEndian + n, isZero = isSingle + 1, first = isZero + 1;
    WW[iIn] = in;
    size_t x = T_Num2BigEndianBytes(c, N, in); cpy(c, bytes, WW + x, n);
    x = T_CountBytes(c, N, WW + bytes); WW[length] = WW[x];
    x = T_ShiftLeft(c, N, WW + bytes, fr_sub(fr_u64(n), WW[length])); cpy(c, bigEndian, WW + x, n);
    x = T_LessThan(c, N * 8, in, fr_u64(128)); WW[isSingle] = WW[x];
    x = T_IsZero(c, in); WW[isZero] = WW[x];
    x = T_Mux1(c, fr_add(fr_u64(0x80), WW[length]), in, WW[isSingle]); WW[first] = WW[x];
    WW[o] = fr_add(WW[first], fr_mul(WW[isZero], fr_u64(0x80)));
    Fr ns = fr_sub(FR_ONE, WW[isSingle]);
    for (size_t i = 1; i < n + 1; i++) WW[o + i] = fr_mul(ns, WW[bigEndian + i - 1]);
    WW[outLen] = fr_add(fr_add(ns, WW[length]), WW[isZero]);
    return o;
}
/* ============================ circuits/utils/rlp/empty_account.circom ============================ */
/* RlpEmptyAccount(maxBalanceBytes) :20-134 ; hash literals :55-120 = keccak(rlp("")) and keccak("") */
static const uint8_t STORAGE_CODE_RLP[66] = {
    160, 86, 232, 31, 23, 27, 204, 85, 166, 255, 131, 69, 230, 146, 192, 248, 110, 91, 72, 224, 27, 153, 108, 173, 192, 1, 98, 47, 181, 227, 99, 180, 33,
    160, 197, 210, 70, 1, 134, 247, 35, 60, 146, 126, 125, 178, 220, 199, 3, 192, 229, 0, 182, 83, 202, 130, 39, 59, 123, 250, 216, 4, 93, 133, 164, 112};
static size_t T_RlpEmptyAccount(Ctx *c, int mbb, Fr balance) {
    size_t m = (size_t)mbb, OL = 4 + m + 66, o = alloc_sig(c, OL + 1 + 1 + (4 + m) + 1 + (m + 1) + 1 + 1 + 66);
    size_t outLen = o + OL, iBal = outLen + 1, pre = iBal + 1, preLen = pre + 4 + m, balRlp = preLen + 1,
           balRlpLen = balRlp + m + 1, nabLen = balRlpLen + 1, sc = nabLen + 1;
    WW[iBal] = balance;
    WW[pre + 2] = fr_u64(0x80);
    size_t r = T_RlpInteger(c, mbb, balance); cpy(c, balRlp, WW + r, m + 1); WW[balRlpLen] = WW[r + m + 1];
    for (size_t i = 0; i < m + 1; i++) WW[pre + 3 + i] = WW[balRlp + i];
    WW[nabLen] = fr_add(FR_ONE, WW[balRlpLen]);
    WW[preLen] = fr_add(fr_u64(2), WW[nabLen]);
    for (size_t i = 0; i < 66; i++) WW[sc + i] = fr_u64(STORAGE_CODE_RLP[i]);
    WW[pre] = fr_u64(0xf8);
    WW[pre + 1] = fr_add(WW[nabLen], fr_u64(66));
    size_t cc = T_Concat(c, 4 + mbb, 66, WW + pre, WW[preLen], WW + sc, fr_u64(66));
    cpy(c, o, WW + cc, OL); WW[outLen] = WW[cc + OL];
    return o;
}
/* ============================ circuits/utils/rlp/merkle_patricia_trie_leaf.circom ============================ */
/* TruncatedAddressHash(addressHashBytes) :50-90 ; `temp` (:76) is declared and never assigned => stays 0 */
static size_t T_TruncatedAddressHash(Ctx *c, int ahb, const Fr *nibbles, Fr nibLen) {
    size_t a = (size_t)ahb, o = alloc_sig(c, (a + 1) + 1 + 2 * a + 1 + 2 + 2 * a + (2 * a + 2) + (2 * a - 1));
    size_t outLen = o + a + 1, iNib = outLen + 1, iLen = iNib + 2 * a, div = iLen + 1, rem = div + 1, shifted = rem + 1,
           outNib = shifted + 2 * a;
    cpy(c, iNib, nibbles, 2 * a); WW[iLen] = nibLen;
    T_AssertLessEqThan(c, 7, nibLen, fr_u64(2 * a));
    size_t d = T_Divide(c, 7, nibLen, fr_u64(2)); WW[div] = WW[d]; WW[rem] = WW[d + 1];
    size_t s = T_ShiftLeft(c, 2 * ahb, nibbles, fr_sub(fr_u64(2 * a), nibLen)); cpy(c, shifted, WW + s, 2 * a);
    WW[outNib] = fr_add(fr_u64(2), WW[rem]);
    WW[outNib + 1] = fr_mul(WW[rem], WW[shifted]);
    for (size_t i = 0; i < 2 * a; i++) {
        if (i < 2 * a - 1) { size_t m = T_Mux1(c, WW[shifted + i], WW[shifted + i + 1], WW[rem]); WW[outNib + i + 2] = WW[m]; }
        else WW[outNib + i + 2] = fr_mul(fr_sub(FR_ONE, WW[rem]), WW[shifted + i]);
    }
    size_t nb = T_Nibbles2Bytes(c, ahb + 1, WW + outNib); cpy(c, o, WW + nb, a + 1);
    WW[outLen] = fr_add(FR_ONE, WW[div]);
    return o;
}
/* RlpMerklePatriciaTrieLeaf(maxAddressHashBytes, maxBalanceBytes) :102-189 */
static size_t T_RlpMerklePatriciaTrieLeaf(Ctx *c, int mahb, int mbb, const Fr *nibbles, Fr nibLen, Fr balance) {
    size_t mrea = 4 + (size_t)mbb + 66, mvr = 2 + mrea, mkl = 1 + (size_t)mahb, mkr = 1 + mkl, mpk = 2 + mkr, MO = mpk + mvr;
    size_t o = alloc_sig(c, MO + 1 + 2 * (size_t)mahb + 1 + 1 + mkl + 1 + mrea + 1 + mpk + 1 + mvr + 1);
    size_t outLen = o + MO, iNib = outLen + 1, iLen = iNib + 2 * (size_t)mahb, iBal = iLen + 1, key = iBal + 1, keyLen = key + mkl,
           rea = keyLen + 1, reaLen = rea + mrea, pk = reaLen + 1, pkLen = pk + mpk, vr = pkLen + 1, vrLen = vr + mvr;
    cpy(c, iNib, nibbles, 2 * (size_t)mahb); WW[iLen] = nibLen; WW[iBal] = balance;
    size_t t = T_TruncatedAddressHash(c, mahb, nibbles, nibLen); cpy(c, key, WW + t, mkl); WW[keyLen] = WW[t + mkl];
    T_AssertGreaterEqThan(c, 16, WW[keyLen], fr_u64(2));
    size_t e = T_RlpEmptyAccount(c, mbb, balance); cpy(c, rea, WW + e, mrea); WW[reaLen] = WW[e + mrea];
    WW[vr] = fr_u64(0xb8); WW[vr + 1] = WW[reaLen];
    for (size_t i = 0; i < mrea; i++) WW[vr + i + 2] = WW[rea + i];
    WW[vrLen] = fr_add(fr_u64(2), WW[reaLen]);
    WW[pk] = fr_u64(0xf8);
    WW[pk + 1] = fr_add(fr_add(WW[keyLen], FR_ONE), WW[vrLen]);
    WW[pk + 2] = fr_add(fr_u64(0x80), WW[keyLen]);
    for (size_t i = 0; i < mkl; i++) WW[pk + i + 3] = WW[key + i];
    WW[pkLen] = fr_add(fr_u64(3), WW[keyLen]);
    size_t cc = T_Concat(c, (int)mpk, (int)mvr, WW + pk, WW[pkLen], WW + vr, WW[vrLen]);
    cpy(c, o, WW + cc, MO); WW[outLen] = WW[cc + MO];
    return o;
}
/* IsInRange(B) :196-207 */
static size_t T_IsInRange(Ctx *c, int B, Fr lower, Fr value, Fr upper) {
    size_t o = alloc_sig(c, 6); WW[o + 1] = lower; WW[o + 2] = value; WW[o + 3] = upper;
    T_AssertBits(c, B, lower); T_AssertBits(c, B, value); T_AssertBits(c, B, upper);
    size_t a = T_LessEqThan(c, B, lower, value); WW[o + 4] = WW[a];
    size_t b = T_LessEqThan(c, B, value, upper); WW[o + 5] = WW[b];
    WW[o] = fr_mul(WW[a], WW[b]); return o;
}
/* LeafDetector(N) :247-294 */
static size_t T_LeafDetector(Ctx *c, int N, const Fr *layer, Fr layerLen) {
    size_t n = (size_t)N, o = alloc_sig(c, 1 + n + 1 + 16);
    cpy(c, o + 1, layer, n); WW[o + 1 + n] = layerLen;
    Fr *v = WW + o + 2 + n;   /* leafPrefixIsF8, totalLength, isConsistentWithLayerLen, keyPrefix, keyPrefixIsValid, keyIsMultiByte,
                                 keyExtraLen, keyLen, valueWrapperPrefix, valueWrapperPrefixIsB8, valueWrapperLen, valuePrefix,
                                 valuePrefixIsF8, valueLen, isValueWrapperLenConsistent, isKeyValueLenEqualWithLayerLen */
    T_AssertLessEqThan(c, 16, layerLen, fr_u64(n));
    size_t x = T_IsEqual(c, layer[0], fr_u64(0xf8)); v[0] = WW[x];
    v[1] = layer[1];
    x = T_IsEqual(c, fr_add(v[1], fr_u64(2)), layerLen); v[2] = WW[x];
    v[3] = layer[2];
    x = T_LessEqThan(c, 16, v[3], fr_u64(0xb7)); v[4] = WW[x];
    x = T_IsInRange(c, 16, fr_u64(0x81), v[3], fr_u64(0xb7)); v[5] = WW[x];
    v[6] = fr_mul(v[5], fr_sub(v[3], fr_u64(0x80)));
    v[7] = fr_add(FR_ONE, v[6]);
    Fr base = fr_add(fr_u64(2), v[7]);
    x = T_Selector(c, N, layer, base); v[8] = WW[x];
    x = T_IsEqual(c, v[8], fr_u64(0xb8)); v[9] = WW[x];
    x = T_Selector(c, N, layer, fr_add(base, FR_ONE)); v[10] = WW[x];
    x = T_Selector(c, N, layer, fr_add(base, fr_u64(2))); v[11] = WW[x];
    x = T_IsEqual(c, v[11], fr_u64(0xf8)); v[12] = WW[x];
    x = T_Selector(c, N, layer, fr_add(base, fr_u64(3))); v[13] = WW[x];
    x = T_IsEqual(c, v[10], fr_add(v[13], fr_u64(2))); v[14] = WW[x];
    x = T_IsEqual(c, fr_add(fr_add(v[7], v[13]), fr_u64(6)), layerLen); v[15] = WW[x];
    Fr ands[7] = {v[0], v[2], v[4], v[9], v[14], v[12], v[15]};
    x = T_MultiAND(c, 7, ands); WW[o] = WW[x];
    return o;
}

/* ============================ circuits/spend.circom ============================ */
/* Spend(maxAmountBytes) :32-53 */
static size_t T_Spend(Ctx *c, int mab, Fr burnKey, Fr balance, Fr withdrawnBalance, Fr extraCommitment) {
    size_t o = alloc_sig(c, 1 + 4 + 2 + 4 * 32);
    WW[o + 1] = burnKey; WW[o + 2] = balance; WW[o + 3] = withdrawnBalance; WW[o + 4] = extraCommitment;
    size_t coin = o + 5, rem = o + 6, by = o + 7;   /* coinBytes, withdrawnBalanceBytes, remainingCoinBytes, extraCommmitmentBytes */
    T_AssertGreaterEqThan(c, mab * 8, balance, withdrawnBalance);
    Fr i1[3] = {POSEIDON_PREFIX(2), burnKey, balance};
    size_t x = T_Poseidon(c, 3, i1); WW[coin] = WW[x];
    Fr i2[3] = {POSEIDON_PREFIX(2), burnKey, fr_sub(balance, withdrawnBalance)};
    x = T_Poseidon(c, 3, i2); WW[rem] = WW[x];
    x = T_Num2BigEndianBytes(c, 32, WW[coin]); cpy(c, by, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, withdrawnBalance); cpy(c, by + 32, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, WW[rem]); cpy(c, by + 64, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, extraCommitment); cpy(c, by + 96, WW + x, 32);
    x = T_PublicCommitment(c, 4, WW + by); WW[o] = WW[x];
    return o;
}

/* ============================ circuits/proof_of_burn.circom ============================ */
typedef struct { int maxNumLayers, maxNodeBlocks, maxHeaderBlocks, minLeafAddressNibbles, amountBytes, powMinimumZeroBytes; Fr maxIntendedBalance, maxActualBalance; } PobParams;
/* ProofOfBurn(...) :34-212 ; inputs in declaration order :43-72 */
static size_t T_ProofOfBurn(Ctx *c, PobParams P, const Fr *in) {
    size_t L = (size_t)P.maxNumLayers, NB = (size_t)P.maxNodeBlocks * 136, HB = (size_t)P.maxHeaderBlocks * 136;
    size_t nIn = 6 + L * NB + L + 1 + HB + 3;
    size_t nMid = 2 + 64 + 32 + 32 + 5 * 32 + NB + 1 + L + (L - 1) + L * 32 + L * 31 + L + 1 + 139 + 1;
    size_t o = alloc_sig(c, 1 + nIn + nMid);
    cpy(c, o + 1, in, nIn);
    /* input views */
    const Fr *I = WW + o + 1;
    Fr burnKey = I[0], actualBalance = I[1], intendedBalance = I[2], revealAmount = I[3], burnExtraCommitment = I[4], numLeafAddressNibbles = I[5];
    const Fr *layers = I + 6, *layerLens = layers + L * NB;
    Fr numLayers = layerLens[L];
    const Fr *blockHeader = layerLens + L + 1;
    Fr blockHeaderLen = blockHeader[HB], byteSecurityRelax = blockHeader[HB + 1], proofExtra = blockHeader[HB + 2];
    /* intermediates */
    size_t remainingCoin = o + 1 + nIn, nullifier = remainingCoin + 1, addrNib = nullifier + 1, blockRoot = addrNib + 64, stateRoot = blockRoot + 32,
           nullB = stateRoot + 32, remB = nullB + 32, revB = remB + 32, becB = revB + 32, ecB = becB + 32, lastLayer = ecB + 32,
           lastLayerLen = lastLayer + NB, layerExists = lastLayerLen + 1, subChk = layerExists + L, layerKec = subChk + (L - 1),
           redKec = layerKec + L * 32, isLeaf = redKec + L * 31, isLastLeaf = isLeaf + L, leaf = isLastLeaf + 1, leafLen = leaf + 139;
    int ab8 = P.amountBytes * 8;
    T_AssertLessEqThan(c, ab8, intendedBalance, P.maxIntendedBalance);                       /* :84 */
    T_AssertLessEqThan(c, ab8, actualBalance, P.maxActualBalance);                           /* :85 */
    T_AssertLessEqThan(c, ab8, intendedBalance, actualBalance);                              /* :86 */
    Fr relax2 = fr_mul(byteSecurityRelax, fr_u64(2)), minNib = fr_u64((uint64_t)P.minLeafAddressNibbles);
    T_AssertLessEqThan(c, 16, relax2, minNib);                                               /* :90 */
    T_AssertGreaterEqThan(c, 16, numLeafAddressNibbles, fr_sub(minNib, relax2));             /* :91 */
    T_AssertBits(c, ab8, revealAmount);                                                      /* :96 */
    T_AssertLessEqThan(c, ab8, revealAmount, intendedBalance);                               /* :97 */
    for (size_t i = 0; i < L; i++) {                                                         /* :99-103 */
        T_AssertLessThan(c, 16, layerLens[i], fr_u64(NB * 8));
        T_AssertByteString(c, (int)NB, layers + i * NB);
    }
    T_AssertLessThan(c, 16, blockHeaderLen, fr_u64(HB * 8));                                 /* :105 */
    T_AssertByteString(c, (int)HB, blockHeader);                                             /* :106 */
    Fr p3[3] = {POSEIDON_PREFIX(2), burnKey, fr_sub(intendedBalance, revealAmount)};
    size_t x = T_Poseidon(c, 3, p3); WW[remainingCoin] = WW[x];                              /* :113 */
    Fr p2[2] = {POSEIDON_PREFIX(1), burnKey};
    x = T_Poseidon(c, 2, p2); WW[nullifier] = WW[x];                                         /* :116 */
    x = T_BurnAddressHash(c, burnKey, revealAmount, burnExtraCommitment); cpy(c, addrNib, WW + x, 64);  /* :119 */
    x = T_KeccakBytes(c, P.maxHeaderBlocks, blockHeader, blockHeaderLen); cpy(c, blockRoot, WW + x, 32); /* :122 */
    for (size_t i = 0; i < 32; i++) WW[stateRoot + i] = blockHeader[91 + i];                 /* :125-129 */
    x = T_Num2BigEndianBytes(c, 32, WW[nullifier]); cpy(c, nullB, WW + x, 32);               /* :132-136 */
    x = T_Num2BigEndianBytes(c, 32, WW[remainingCoin]); cpy(c, remB, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, revealAmount); cpy(c, revB, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, burnExtraCommitment); cpy(c, becB, WW + x, 32);
    x = T_Num2BigEndianBytes(c, 32, proofExtra); cpy(c, ecB, WW + x, 32);
    {                                                                                        /* :137-139 */
        Fr *six = tmp_alloc(192);
        memcpy(six, WW + blockRoot, 32 * sizeof(Fr)); memcpy(six + 32, WW + nullB, 32 * sizeof(Fr)); memcpy(six + 64, WW + remB, 32 * sizeof(Fr));
        memcpy(six + 96, WW + revB, 32 * sizeof(Fr)); memcpy(six + 128, WW + becB, 32 * sizeof(Fr)); memcpy(six + 160, WW + ecB, 32 * sizeof(Fr));
        x = T_PublicCommitment(c, 6, six); WW[o] = WW[x]; free(six);
    }
    Fr selLast = fr_sub(numLayers, FR_ONE);
    x = T_SelectorArray(c, P.maxNumLayers, NB, layers, selLast); cpy(c, lastLayer, WW + x, NB);       /* :142-143 */
    x = T_Selector(c, P.maxNumLayers, layerLens, selLast); WW[lastLayerLen] = WW[x];                  /* :146 */
    x = T_Filter(c, P.maxNumLayers, numLayers); cpy(c, layerExists, WW + x, L);                       /* :150 */
    Fr numLeaves = FR_ZERO;
    for (size_t i = 0; i < L; i++) {                                                                  /* :157-181 */
        x = T_LeafDetector(c, (int)NB, layers + i * NB, layerLens[i]); WW[isLeaf + i] = WW[x];
        numLeaves = fr_add(numLeaves, WW[x]);
        x = T_KeccakBytes(c, P.maxNodeBlocks, layers + i * NB, layerLens[i]); cpy(c, layerKec + 32 * i, WW + x, 32);
        x = T_Fit(c, 32, 31, WW + layerKec + 32 * i); cpy(c, redKec + 31 * i, WW + x, 31);
        if (i > 0) {
            x = T_SubstringCheck(c, (int)NB, 31, layers + (i - 1) * NB, layerLens[i - 1], WW + redKec + 31 * i);
            WW[subChk + i - 1] = WW[x];
            check(c, fr_is_zero(fr_mul(fr_sub(FR_ONE, WW[x]), WW[layerExists + i])), o);             /* :179 */
        }
    }
    check(c, fr_eq(numLeaves, FR_ONE), o);                                                            /* :186 */
    x = T_LeafDetector(c, (int)NB, WW + lastLayer, WW[lastLayerLen]); WW[isLastLeaf] = WW[x];         /* :187 */
    check(c, fr_eq(WW[isLastLeaf], FR_ONE), o);                                                       /* :188 */
    for (size_t i = 0; i < 32; i++) check(c, fr_eq(WW[layerKec + i], WW[stateRoot + i]), o);          /* :191-193 */
    x = T_RlpMerklePatriciaTrieLeaf(c, 32, P.amountBytes, WW + addrNib, numLeafAddressNibbles, actualBalance); /* :198-200 */
    cpy(c, leaf, WW + x, 139); WW[leafLen] = WW[x + 139];
    for (size_t i = 0; i < 139; i++) check(c, fr_eq(WW[leaf + i], WW[lastLayer + i]), o);             /* :203-205 */
    check(c, fr_eq(WW[leafLen], WW[lastLayerLen]), o);                                                /* :206 */
    T_ProofOfWorkChecker(c, burnKey, revealAmount, burnExtraCommitment,
                         fr_add(fr_u64((uint64_t)P.powMinimumZeroBytes), byteSecurityRelax));         /* :211 */
    return o;
}

/* ============================ main-component dispatcher ============================ */
/* One entry per `component main = ...` expression the reference builds: the two product mains
 * (main_proof_of_burn.circom:27, main_spend.circom:6) and every gadget suite of tests/test.py:146-201. */
typedef struct { const char *name; int nparams; const char *schema; } MainInfo;
/* schema: comma separated input names in declaration order; dimensions are expressions over p0..p7 */
static const MainInfo MAINS[] = {
    {"Spend", 1, "burnKey,balance,withdrawnBalance,extraCommitment"},
    {"ProofOfBurn", 8, "burnKey,actualBalance,intendedBalance,revealAmount,burnExtraCommitment,numLeafAddressNibbles,layers[p0][p1*136],layerLens[p0],numLayers,blockHeader[p2*136],blockHeaderLen,byteSecurityRelax,_proofExtraCommitment"},
    {"EIP7503", 0, ""},
    {"ConcatFixed4", 4, "a[p0],b[p1],c[p2],d[p3]"},
    {"ProofOfWorkChecker", 0, "burnKey,revealAmount,burnExtraCommitment,minimumZeroBytes"},
    {"PublicCommitment", 1, "in[p0][32]"},
    {"Poseidon", 1, "inputs[p0]"},
    {"Divide", 1, "a,b"},
    {"SubstringCheck", 2, "mainInput[p0],mainLen,subInput[p1]"},
    {"ShiftLeft", 1, "in[p0],count"},
    {"ShiftRight", 2, "in[p0],count"},
    {"Mask", 1, "in[p0],count"},
    {"Concat", 2, "a[p0],aLen,b[p1],bLen"},
    {"Selector", 1, "vals[p0],select"},
    {"SelectorArray1D", 2, "arrays[p0][p1],select"},
    {"SelectorArray2D", 3, "arrays[p0][p1][p2],select"},
    {"BigEndianBytes2Num", 1, "in[p0]"},
    {"LittleEndianBytes2Num", 1, "in[p0]"},
    {"Bytes2Nibbles", 1, "in[p0]"},
    {"Num2BigEndianBytes", 1, "in"},
    {"Num2LittleEndianBytes", 1, "in"},
    {"Nibbles2Bytes", 1, "nibbles[2*p0]"},
    {"Num2BitsSafe", 1, "in"},
    {"Pad", 2, "in[p0*p1],inLen"},
    {"KeccakBytes", 1, "in[p0*136],inLen"},
    {"BurnAddress", 0, "burnKey,revealAmount,burnExtraCommitment"},
    {"BurnAddressHash", 0, "burnKey,revealAmount,burnExtraCommitment"},
    {"AssertBits", 1, "in"},
    {"AssertByteString", 1, "in[p0]"},
    {"AssertLessThan", 1, "a,b"},
    {"AssertLessEqThan", 1, "a,b"},
    {"AssertGreaterEqThan", 1, "a,b"},
    {"Filter", 1, "in"},
    {"Fit", 2, "in[p0]"},
    {"Reverse", 1, "in[p0]"},
    {"Flatten", 2, "in[p0][p1]"},
    {"Reshape", 2, "in[p0*p1]"},
    {"RlpInteger", 1, "in"},
    {"CountBytes", 1, "bytes[p0]"},
    {"RlpEmptyAccount", 1, "balance"},
    {"TruncatedAddressHash", 1, "addressHashNibbles[2*p0],addressHashNibblesLen"},
    {"IsInRange", 1, "lower,value,upper"},
    {"LeafDetector", 1, "layer[p0],layerLen"},
    {"RlpMerklePatriciaTrieLeaf", 2, "addressHashNibbles[2*p0],addressHashNibblesLen,balance"},
    {NULL, 0, NULL}};

const char *pob_oracle_schema(const char *name, int *nparams) {
    for (const MainInfo *m = MAINS; m->name; m++) if (!strcmp(m->name, name)) { if (nparams) *nparams = m->nparams; return m->schema; }
    return NULL;
}

static int PI(const Fr *p, int i) { return (int)p[i].l[0]; }

/* returns number of main output signals, or -1 on unknown main */
static int run_main(Ctx *c, const char *name, const Fr *p, const Fr *in) {
#define IS(s) (!strcmp(name, s))
    if (IS("Spend")) { T_Spend(c, PI(p, 0), in[0], in[1], in[2], in[3]); return 1; }
    if (IS("ProofOfBurn")) {
        PobParams P = {PI(p, 0), PI(p, 1), PI(p, 2), PI(p, 3), PI(p, 4), PI(p, 5), p[6], p[7]};
        T_ProofOfBurn(c, P, in); return 1;
    }
    if (IS("EIP7503")) { T_EIP7503(c); return 8; }
    if (IS("ConcatFixed4")) { int A = PI(p, 0), B = PI(p, 1), C = PI(p, 2), D = PI(p, 3); T_ConcatFixed4(c, A, B, C, D, in, in + A, in + A + B, in + A + B + C); return A + B + C + D; }
    if (IS("ProofOfWorkChecker")) { T_ProofOfWorkChecker(c, in[0], in[1], in[2], in[3]); return 0; }
    if (IS("PublicCommitment")) { T_PublicCommitment(c, PI(p, 0), in); return 1; }
    if (IS("Poseidon")) { T_Poseidon(c, PI(p, 0), in); return 1; }
    if (IS("Divide")) { T_Divide(c, PI(p, 0), in[0], in[1]); return 2; }
    if (IS("SubstringCheck")) { int mm = PI(p, 0); T_SubstringCheck(c, mm, PI(p, 1), in, in[mm], in + mm + 1); return 1; }
    if (IS("ShiftLeft")) { int n = PI(p, 0); T_ShiftLeft(c, n, in, in[n]); return n; }
    if (IS("ShiftRight")) { int n = PI(p, 0), ms = PI(p, 1); T_ShiftRight(c, n, ms, in, in[n]); return n + ms; }
    if (IS("Mask")) { int n = PI(p, 0); T_Mask(c, n, in, in[n]); return n; }
    if (IS("Concat")) { int A = PI(p, 0), B = PI(p, 1); T_Concat(c, A, B, in, in[A], in + A + 1, in[A + 1 + B]); return A + B + 1; }
    if (IS("Selector")) { int n = PI(p, 0); T_Selector(c, n, in, in[n]); return 1; }
    if (IS("SelectorArray1D")) { int n = PI(p, 0), q = PI(p, 1); T_SelectorArray(c, n, (size_t)q, in, in[n * q]); return q; }
    if (IS("SelectorArray2D")) { int n = PI(p, 0), q = PI(p, 1) * PI(p, 2); T_SelectorArray(c, n, (size_t)q, in, in[n * q]); return q; }
    if (IS("BigEndianBytes2Num")) { T_BigEndianBytes2Num(c, PI(p, 0), in); return 1; }
    if (IS("LittleEndianBytes2Num")) { T_LittleEndianBytes2Num(c, PI(p, 0), in); return 1; }
    if (IS("Bytes2Nibbles")) { T_Bytes2Nibbles(c, PI(p, 0), in); return 2 * PI(p, 0); }
    if (IS("Num2BigEndianBytes")) { T_Num2BigEndianBytes(c, PI(p, 0), in[0]); return PI(p, 0); }
    if (IS("Num2LittleEndianBytes")) { T_Num2LittleEndianBytes(c, PI(p, 0), in[0]); return PI(p, 0); }
    if (IS("Nibbles2Bytes")) { T_Nibbles2Bytes(c, PI(p, 0), in); return PI(p, 0); }
    if (IS("Num2BitsSafe")) { T_Num2BitsSafe(c, PI(p, 0), in[0]); return PI(p, 0); }
    if (IS("Pad")) { int B = PI(p, 0) * PI(p, 1); T_Pad(c, PI(p, 0), PI(p, 1), in, in[B]); return B + 1; }
    if (IS("KeccakBytes")) { int B = PI(p, 0) * 136; T_KeccakBytes(c, PI(p, 0), in, in[B]); return 32; }
    if (IS("BurnAddress")) { T_BurnAddress(c, in[0], in[1], in[2]); return 20; }
    if (IS("BurnAddressHash")) { T_BurnAddressHash(c, in[0], in[1], in[2]); return 64; }
    if (IS("AssertBits")) { T_AssertBits(c, PI(p, 0), in[0]); return 0; }
    if (IS("AssertByteString")) { T_AssertByteString(c, PI(p, 0), in); return 0; }
    if (IS("AssertLessThan")) { T_AssertLessThan(c, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("AssertLessEqThan")) { T_AssertLessEqThan(c, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("AssertGreaterEqThan")) { T_AssertGreaterEqThan(c, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("Filter")) { T_Filter(c, PI(p, 0), in[0]); return PI(p, 0); }
    if (IS("Fit")) { T_Fit(c, PI(p, 0), PI(p, 1), in); return PI(p, 1); }
    if (IS("Reverse")) { T_Reverse(c, PI(p, 0), in); return PI(p, 0); }
    if (IS("Flatten") || IS("Reshape")) { size_t n = (size_t)PI(p, 0) * (size_t)PI(p, 1); T_CopyArray(c, n, in); return (int)n; }
    if (IS("RlpInteger")) { T_RlpInteger(c, PI(p, 0), in[0]); return PI(p, 0) + 2; }
    if (IS("CountBytes")) { T_CountBytes(c, PI(p, 0), in); return 1; }
    if (IS("RlpEmptyAccount")) { T_RlpEmptyAccount(c, PI(p, 0), in[0]); return 4 + PI(p, 0) + 66 + 1; }
    if (IS("TruncatedAddressHash")) { int a = PI(p, 0); T_TruncatedAddressHash(c, a, in, in[2 * a]); return a + 2; }
    if (IS("IsInRange")) { T_IsInRange(c, PI(p, 0), in[0], in[1], in[2]); return 1; }
    if (IS("LeafDetector")) { int n = PI(p, 0); T_LeafDetector(c, n, in, in[n]); return 1; }
    if (IS("RlpMerklePatriciaTrieLeaf")) {
        int a = PI(p, 0), b = PI(p, 1); T_RlpMerklePatriciaTrieLeaf(c, a, b, in, in[2 * a], in[2 * a + 1]);
        return (2 + 1 + 1 + a) + (2 + 4 + b + 66) + 1;
    }
    return -1;
#undef IS
}

/* ---- public API (ctypes) ----
 * params: nparams x 4 limbs; inputs: n_inputs x 4 limbs, canonical, flattened in declaration order.
 * On success *witness points at n_signals x 32 bytes (owned by the oracle; free with pob_oracle_free). */
typedef struct { Fr *w; size_t cap_bytes; } OracleBuf;
#define ORACLE_MAX_SIGNALS ((size_t)300 * 1000 * 1000)

int pob_oracle_run(const char *main_name, const uint64_t *params, int nparams, const uint64_t *inputs, size_t n_inputs,
                   int hcreate, uint64_t **witness, uint64_t *n_signals, uint32_t *n_outputs, uint64_t *status) {
    fr_init();
    (void)nparams; (void)n_inputs;
    Ctx c; memset(&c, 0, sizeof c);
    size_t bytes = ORACLE_MAX_SIGNALS * sizeof(Fr);
    c.w = (Fr *)mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (c.w == MAP_FAILED) return -2;
    c.cap = ORACLE_MAX_SIGNALS; c.hcreate = hcreate;
    c.w[0] = FR_ONE; c.pos = 1;
    int nout = run_main(&c, main_name, (const Fr *)params, (const Fr *)inputs);
    if (nout < 0) { munmap(c.w, bytes); return -1; }
    *witness = (uint64_t *)c.w; *n_signals = c.pos; *n_outputs = (uint32_t)nout; *status = c.status;
    return 0;
}
void pob_oracle_free(uint64_t *witness) { if (witness) munmap(witness, ORACLE_MAX_SIGNALS * sizeof(Fr)); }

/* .wtns writer (the iden3 binary witness format the circom runtime's writeBinWitness emits) */
int pob_oracle_write_wtns(const char *path, const uint64_t *witness, uint64_t n_signals) {
    FILE *f = fopen(path, "wb"); if (!f) return -1;
    uint32_t u32; uint64_t u64;
    fwrite("wtns", 1, 4, f);
    u32 = 2; fwrite(&u32, 4, 1, f); u32 = 2; fwrite(&u32, 4, 1, f);
    u32 = 1; fwrite(&u32, 4, 1, f); u64 = 40; fwrite(&u64, 8, 1, f);
    u32 = 32; fwrite(&u32, 4, 1, f); fwrite(FR_P.l, 8, 4, f); u32 = (uint32_t)n_signals; fwrite(&u32, 4, 1, f);
    u32 = 2; fwrite(&u32, 4, 1, f); u64 = 32 * n_signals; fwrite(&u64, 8, 1, f);
    size_t wr = fwrite(witness, 32, n_signals, f);
    fclose(f);
    return wr == n_signals ? 0 : -1;
}

/* order-independent digest used by the full-size parity tests: sum_i mix(i, limbs) mod 2^64 */
uint64_t pob_oracle_digest(const uint64_t *witness, uint64_t n_signals) {
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_signals; i++) {
        const uint64_t *l = witness + 4 * i;
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ULL;
        h ^= l[0] * 0xBF58476D1CE4E5B9ULL + l[1] * 0x94D049BB133111EBULL + l[2] * 0xD6E8FEB86659FD93ULL + l[3] * 0xA0761D6478BD642FULL;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
        acc += h;
    }
    return acc;
}
