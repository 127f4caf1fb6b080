/* oracle/fr.h -- BN254 scalar field (Fr) arithmetic for the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from the
 * product path (proof-of-burn_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * Restates the arithmetic the circom-generated calculator performs under every signal
 * (reference: the implicit field ops in every `<==`/`<--`, e.g. circomlib/circuits/gates.circom:26,
 * comparators.circom:30-33; prime from tests/poseidon.py:1-3).  The reference's own fr.cpp/fr.asm is
 * emitted by the external circom toolchain (iden3/circom, unpinned master per Dockerfile:5) and is not
 * in the tree, so this is a from-scratch restatement of ordinary Montgomery arithmetic.
 * Values are kept CANONICAL (non-Montgomery) in [0,p), little-endian 4x64-bit limbs -- exactly the
 * 32-byte form the .wtns file stores.
 */
#ifndef POB_ORACLE_FR_H
#define POB_ORACLE_FR_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[4]; } Fr;
typedef unsigned __int128 u128;

static const Fr FR_P = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const Fr FR_ZERO = {{0, 0, 0, 0}};
static const Fr FR_ONE = {{1, 0, 0, 0}};

static uint64_t FR_N0;  /* -p^-1 mod 2^64 */
static Fr FR_R2;        /* 2^512 mod p   */

static inline Fr fr_u64(uint64_t v) { Fr r = {{v, 0, 0, 0}}; return r; }
static inline int fr_is_zero(Fr a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int fr_eq(Fr a, Fr b) { return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3]; }
static inline int fr_fits64(Fr a) { return (a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int fr_cmp(Fr a, Fr b) {
    for (int i = 3; i >= 0; i--) { if (a.l[i] < b.l[i]) return -1; if (a.l[i] > b.l[i]) return 1; }
    return 0;
}
static inline int fr_bit(Fr a, unsigned i) { return i < 256 ? (int)((a.l[i >> 6] >> (i & 63)) & 1) : 0; }

static inline uint64_t raw_add(Fr *r, Fr a, Fr b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t raw_sub(Fr *r, Fr a, Fr b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - br; r->l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}
static inline Fr fr_add(Fr a, Fr b) {
    Fr r; uint64_t c = raw_add(&r, a, b);
    if (c || fr_cmp(r, FR_P) >= 0) { Fr t; raw_sub(&t, r, FR_P); return t; }
    return r;
}
static inline Fr fr_sub(Fr a, Fr b) {
    Fr r; if (raw_sub(&r, a, b)) { Fr t; raw_add(&t, r, FR_P); return t; }
    return r;
}
static inline Fr fr_neg(Fr a) { return fr_is_zero(a) ? a : fr_sub(FR_P, a); }

/* Montgomery product a*b*2^-256 mod p (CIOS, 64-bit limbs). */
static inline Fr fr_mont(Fr a, Fr b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_N0;
        c = (u128)m * FR_P.l[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * FR_P.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fr_cmp(r, FR_P) >= 0) { Fr s; raw_sub(&s, r, FR_P); return s; }
    return r;
}
/* canonical product */
static inline Fr fr_mul(Fr a, Fr b) {
    if (fr_fits64(a) && fr_fits64(b)) {            /* product < 2^128 < p: no reduction */
        u128 m = (u128)a.l[0] * b.l[0]; Fr r = {{(uint64_t)m, (uint64_t)(m >> 64), 0, 0}}; return r;
    }
    return fr_mont(fr_mont(a, b), FR_R2);
}
static inline Fr fr_pow2(unsigned n) {            /* 2^n mod p */
    if (n < 254) { Fr r = FR_ZERO; r.l[n >> 6] = 1ULL << (n & 63); return r; }
    Fr r = fr_pow2(253); for (unsigned i = 253; i < n; i++) r = fr_add(r, r); return r;
}
static Fr fr_pow(Fr a, Fr e) {
    Fr r = FR_ONE;
    for (int i = 255; i >= 0; i--) { r = fr_mul(r, r); if (fr_bit(e, (unsigned)i)) r = fr_mul(r, a); }
    return r;
}
static Fr fr_inv_slow(Fr a) { Fr e; raw_sub(&e, FR_P, fr_u64(2)); return fr_pow(a, e); }

/* small-inverse cache: most IsZero inputs in these circuits are differences |x| < 2^16 */
#define FR_INV_CACHE 65537
static Fr fr_inv_cache[FR_INV_CACHE];
static uint8_t fr_inv_have[FR_INV_CACHE];
static Fr fr_inv(Fr a) {            /* a != 0 */
    if (fr_fits64(a) && a.l[0] < FR_INV_CACHE) {
        uint64_t k = a.l[0];
        if (!fr_inv_have[k]) { fr_inv_cache[k] = fr_inv_slow(a); fr_inv_have[k] = 1; }
        return fr_inv_cache[k];
    }
    Fr n = fr_sub(FR_P, a);
    if (fr_fits64(n) && n.l[0] < FR_INV_CACHE) return fr_neg(fr_inv(n));
    return fr_inv_slow(a);
}

static void fr_init(void) {
    static int done = 0; if (done) return; done = 1;
    uint64_t x = 1;                                  /* Newton: x = p^-1 mod 2^64 */
    for (int i = 0; i < 7; i++) x *= 2 - FR_P.l[0] * x;
    FR_N0 = (uint64_t)0 - x;
    Fr r = FR_ONE; for (int i = 0; i < 512; i++) r = fr_add(r, r);
    FR_R2 = r;
}

/* integer quotient / remainder on canonical representatives (circom `\` and `%`) */
static void fr_divmod(Fr a, Fr b, Fr *q, Fr *r) {
    Fr quo = FR_ZERO, rem = FR_ZERO;
    for (int i = 255; i >= 0; i--) {
        /* rem = rem*2 + bit */
        uint64_t c = 0;
        for (int k = 0; k < 4; k++) { uint64_t n = (rem.l[k] << 1) | c; c = rem.l[k] >> 63; rem.l[k] = n; }
        rem.l[0] |= (uint64_t)fr_bit(a, (unsigned)i);
        if (fr_cmp(rem, b) >= 0) { Fr t; raw_sub(&t, rem, b); rem = t; quo.l[i >> 6] |= 1ULL << (i & 63); }
    }
    *q = quo; *r = rem;
}
#endif
