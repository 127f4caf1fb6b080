// vm_exec.h -- semantics of one witness-VM thread op and of operand/witness code decoding.
//
// __host__ __device__: the device instantiation is what pob_b200.cu runs (the product); the host instantiation
// exists only for tests/emu/ (a test-only emulator that checks compiled programs against the oracle on a box
// without a GPU, together with the scalar references of the warp ops in tests/emu/host_ref.h).  Nothing in the
// shipped library executes these on the host.
#pragma once
#include "program.h"

namespace pob {

static const uint32_t INV_TABLE_N = 1u << 16;     // inverses of 1 .. 65535 (most IsZero inputs are small differences)
static const uint32_t STATUS_OK = 0xffffffffu;

struct VmCtx {
    uint64_t *U;            // instance store
    uint32_t val_base;
    const Fr *konst;
    const Code *aux;
    const Fr *invtab;
    uint32_t *status;       // min over failing constraints of (component base + 1); STATUS_OK if none
};

// A value slot is 32 bytes, 32-byte aligned (val_base is a multiple of 4 words, the store stride of 32, witness entries are 32 bytes):
// on the device one 256-bit access per value.  As four 64-bit accesses a warp's 32 values (a contiguous 1 KB run) cost four
// instructions of 32 half-used sectors each, and the thread-op levels of k_eval were bound by exactly that.
POB_HD Fr vm_load_val(const uint64_t *p) {
    Fr r;
#ifdef __CUDA_ARCH__
    uint64_t v[4];
    asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p) : "memory");
#pragma unroll
    for (int i = 0; i < 4; i++) { r.l[2 * i] = (uint32_t)v[i]; r.l[2 * i + 1] = (uint32_t)(v[i] >> 32); }
#else
#pragma unroll
    for (int i = 0; i < 4; i++) { uint64_t v = p[i]; r.l[2 * i] = (uint32_t)v; r.l[2 * i + 1] = (uint32_t)(v >> 32); }
#endif
    return r;
}
POB_HD void vm_store_val(uint64_t *p, const Fr &v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" :: "l"(p), "l"((uint64_t)v.l[0] | ((uint64_t)v.l[1] << 32)), "l"((uint64_t)v.l[2] | ((uint64_t)v.l[3] << 32)),
                 "l"((uint64_t)v.l[4] | ((uint64_t)v.l[5] << 32)), "l"((uint64_t)v.l[6] | ((uint64_t)v.l[7] << 32)) : "memory");
#else
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = (uint64_t)v.l[2 * i] | ((uint64_t)v.l[2 * i + 1] << 32);
#endif
}
POB_HD Fr vm_load(const VmCtx &x, Code c) {
    uint32_t k = code_kind(c), p = code_payload(c);
    if (k == K_CONST) return fr_from_u64(p);
    if (k == K_BIT) return fr_from_u64((x.U[p >> 6] >> (p & 63)) & 1ull);
    if (k == K_VAL) return vm_load_val(x.U + x.val_base + 4ull * p);
    return x.konst[p];
}
POB_HD void vm_fail(const VmCtx &x, uint32_t base) {
#if defined(__CUDA_ARCH__)
    atomicMin(x.status, base + 1u);
#else
    if (base + 1u < *x.status) *x.status = base + 1u;
#endif
}
POB_HD Fr vm_inverse(const VmCtx &x, const Fr &a) {
    if (fr_is_zero(a)) return a;
    if (fr_fits64(a) && fr_lo64(a) < INV_TABLE_N) return x.invtab[fr_lo64(a)];
    Fr n; fr_raw_sub(n, fr_p(), a);
    if (fr_fits64(n) && fr_lo64(n) < INV_TABLE_N) return fr_neg(x.invtab[fr_lo64(n)]);
    return fr_inv_eea(a);
}

POB_HD void vm_exec_op(const VmCtx &x, const Op &op) {
    uint32_t opc = op_opc(op), dst = op_dst(op);
    uint64_t *vd = x.U + x.val_base + 4ull * dst;
    switch (opc) {
    case OP_FMA: {
        // the two commonest multipliers never need a field multiplication: b == 1 (a + c) and b == -1 (c - a;
        // konst[0] is p-1 by construction -- Builder::MINUS1)
        Fr a = vm_load(x, op.a), c = vm_load(x, op.c), r;
        if (op.b == c_const(1)) r = fr_add(a, c);
        else if (op.b == c_konst(0)) r = fr_sub(c, a);
        else { Fr b = vm_load(x, op.b); r = fr_add(fr_mul(a, b), c); }
        vm_store_val(vd, r);
        break; }
    case OP_ISZ: { Fr a = vm_load(x, op.a); vm_store_val(vd, fr_from_u64(fr_is_zero(a) ? 1 : 0)); break; }
    case OP_INV: { Fr a = vm_load(x, op.a); vm_store_val(vd, vm_inverse(x, a)); break; }
    case OP_DIV:
    case OP_MOD: {
        Fr a = vm_load(x, op.a), b = vm_load(x, op.b), q = fr_zero(), r = fr_zero();
        if (fr_is_zero(b)) vm_fail(x, op.c); else fr_divmod(a, b, q, r);
        vm_store_val(vd, opc == OP_DIV ? q : r);
        break; }
    case OP_PACK8: {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) { Fr v = vm_load(x, x.aux[op.a + k]); w |= (uint64_t)(v.l[0] & 0xffu) << (8 * k); }
        x.U[dst] = w;
        break; }
    case OP_GTC: {
        Fr a = vm_load(x, op.a);
        bool gt = !fr_fits64(a) || fr_lo64(a) > (uint64_t)op.b;
        vm_store_val(vd, fr_from_u64(gt ? 1 : 0));
        break; }
    case OP_SELSUM: {
        Fr sel = vm_load(x, op.a), r = fr_zero();
        if (fr_fits64(sel) && fr_lo64(sel) <= (uint64_t)op.c) r = vm_load(x, x.aux[op.b + (uint32_t)fr_lo64(sel)]);
        vm_store_val(vd, r);
        break; }
    case OP_CHK_EQ: { Fr a = vm_load(x, op.a), b = vm_load(x, op.b); if (!fr_eq(a, b)) vm_fail(x, op.c); break; }
    case OP_CHK_RANGE: { Fr a = vm_load(x, op.a); if (!fr_lt_pow2(a, op.b)) vm_fail(x, op.c); break; }
    default: break;
    }
}

// Deferred IsZero inverses (comparators.circom:30).  Thread `tid` of `nthr` owns ops begin+tid, +nthr, ...: zero and
// table-sized inputs are answered directly, the rest share ONE field inversion per thread (Montgomery's trick:
// prefix products parked in the destination slots, then unwound in reverse).
POB_HD int vm_inv_class(const VmCtx &x, const Fr &a, Fr &direct) {
    if (fr_is_zero(a)) { direct = a; return 0; }
    if (fr_fits64(a) && fr_lo64(a) < INV_TABLE_N) { direct = x.invtab[fr_lo64(a)]; return 0; }
    Fr n; fr_raw_sub(n, fr_p(), a);
    if (fr_fits64(n) && fr_lo64(n) < INV_TABLE_N) { direct = fr_neg(x.invtab[fr_lo64(n)]); return 0; }
    return 1;
}
POB_HD void vm_inv_batch(const VmCtx &x, const Op *ops, uint32_t begin, uint32_t end, uint32_t tid, uint32_t nthr) {
    if (begin + tid >= end) return;
    Fr acc = fr_from_u64(1); bool any = false;
    uint32_t last = begin + tid;
    for (uint32_t i = begin + tid; i < end; i += nthr) {
        last = i;
        Fr a = vm_load(x, ops[i].a), d;
        uint64_t *vd = x.U + x.val_base + 4ull * op_dst(ops[i]);
        if (vm_inv_class(x, a, d) == 0) vm_store_val(vd, d);
        else { vm_store_val(vd, acc); acc = fr_mul(acc, a); any = true; }
    }
    if (!any) return;
    Fr inv = fr_inv_eea(acc);
    for (uint32_t i = last;; i -= nthr) {
        Fr a = vm_load(x, ops[i].a), d;
        if (vm_inv_class(x, a, d) != 0) {
            uint64_t *vd = x.U + x.val_base + 4ull * op_dst(ops[i]);
            Fr pre = vm_load_val(vd);
            vm_store_val(vd, fr_mul(inv, pre));
            inv = fr_mul(inv, a);
        }
        if (i < begin + tid + nthr) break;
    }
}

// ---- the deferred inverses as a state machine spread over the levels (k_eval; emulated 1:1 by tests/emu) ------------------------
// Worker w of nw owns the deferred ops begin + w, + nw, ...  START (the first level at which all inputs are ready): zero /
// table-sized inputs are answered at once, the others get their prefix product parked in the destination slot (Montgomery's
// trick) and the worker's total product becomes an inversion in progress.  STEP (every later level): a bounded number of
// iterations of the inversion, state parked by the caller.  FINISH (as soon as the inversion is through, at the latest after the last level): the products are unwound.
// The inversion itself is Kaliski's "almost inverse" (phase 1 of the Montgomery inverse): per iteration one of four cheap cases
// on (u, v, r, s) -- halve u, halve v, or subtract-and-halve -- with no modular halving and no inner loops (in SIMT the
// per-lane `while (even) halve` loops of the textbook binary Euclid run for the maximum over the warp's lanes: measured 5x
// slower).  After k iterations (254 <= k <= 508) r = -a^-1 2^k mod p; two Montgomery products remove the 2^k.
struct InvChain { Fr u, v, r, s; uint32_t k; };
POB_HD uint32_t bn_sub(Fr &d, const Fr &a, const Fr &b) { return fr_raw_sub(d, a, b); }   // d = a - b mod 2^256, returns the borrow (one carry chain on the device)
POB_HD void bn_add(Fr &d, const Fr &a, const Fr &b) { fr_raw_add(d, a, b); }                // d = a + b (the caller knows it fits)
POB_HD void bn_shl1(Fr &a) {
#pragma unroll
    for (int i = 7; i > 0; i--) a.l[i] = (a.l[i] << 1) | (a.l[i - 1] >> 31);
    a.l[0] <<= 1;
}
POB_HD uint32_t bn_sel(uint32_t m, uint32_t a, uint32_t b) { return b ^ ((a ^ b) & m); }      // m all-ones: a, zero: b
POB_HD void inv_chain_init(InvChain &c, const Fr &a) { c.u = fr_p(); c.v = a; c.r = fr_zero(); c.s = fr_from_u64(1); c.k = 0; }
// One iteration is written WITHOUT data-dependent branches (both differences, the sum and the doublings are computed, the case
// picks among them): in a warp the four cases would run one after the other, each a dependent carry chain -- measured 625 cycles per
// iteration with branches.
POB_HD bool inv_chain_steps(InvChain &c, uint32_t steps) {                         // true: finished (v == 0)
    for (uint32_t n = 0; n < steps; n++) {
        if (fr_is_zero(c.v)) return true;
        const bool ue = !(c.u.l[0] & 1u), ve = !(c.v.l[0] & 1u);
        Fr d, e, sum;
        const uint32_t lt = bn_sub(d, c.u, c.v); bn_sub(e, c.v, c.u); bn_add(sum, c.r, c.s);
        const bool both_odd = !ue && !ve, gt = both_odd && !lt && !fr_is_zero(d);   // u > v
        const bool side_u = ue || gt;                                               // the case halves u (else v)
        const bool sub = both_odd;                                                  // ... after subtracting the other one
        const uint32_t mu = 0u - (uint32_t)side_u, ms = 0u - (uint32_t)sub;        // masks: bitwise selects keep the compiler from branching
        Fr t, r2 = c.r, s2 = c.s;
#pragma unroll
        for (int i = 0; i < 8; i++) t.l[i] = bn_sel(mu, bn_sel(ms, d.l[i], c.u.l[i]), bn_sel(ms, e.l[i], c.v.l[i]));
        fr_shr1(t); bn_shl1(r2); bn_shl1(s2);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c.u.l[i] = bn_sel(mu, t.l[i], c.u.l[i]);
            c.v.l[i] = bn_sel(mu, c.v.l[i], t.l[i]);
            const uint32_t rs = bn_sel(ms, sum.l[i], bn_sel(mu, c.r.l[i], c.s.l[i]));   // the side that is not doubled: r (+ s) resp. s (+ r)
            c.r.l[i] = bn_sel(mu, rs, r2.l[i]);
            c.s.l[i] = bn_sel(mu, s2.l[i], rs);
        }
        c.k++;
    }
    return fr_is_zero(c.v);
}
POB_HD Fr inv_chain_result(const InvChain &c) {                                     // a^-1 (canonical) once inv_chain_steps returned true
    Fr r = c.r;
    if (fr_geq_p(r)) { Fr t; fr_raw_sub(t, r, fr_p()); r = t; }                      // r < 2p
    { Fr t; fr_raw_sub(t, fr_p(), r); r = t; }                                       // r = a^-1 2^k mod p, in [1, p-1] (r == 0 cannot happen for a != 0)
    uint32_t k = c.k;                                                                // 254 <= k <= 508
    if (k > 256) { r = fr_mont(r, fr_from_u64(1)); k -= 256; }                      // r * 2^-256; now 0 < k <= 256
    const uint32_t e = 256 - k;                                                      // r * 2^e * 2^-256 = r * 2^-k
    Fr pw = fr_zero(); pw.l[e >> 5] = 1u << (e & 31);
    while (fr_geq_p(pw)) { Fr t; fr_raw_sub(t, pw, fr_p()); pw = t; }               // e = 254, 255 only
    return fr_mont(r, pw);
}
// Products are taken with the bare Montgomery product (x * y * R^-1, R = 2^256), never converted: acc_i = acc_{i-1} (*) a_i is
// a_1..a_i * R^-i, its field inverse I_n = (a_1..a_n)^-1 * R^n, and then I_n (*) acc_{n-1} = a_n^-1 exactly and I_n (*) a_n = I_{n-1}:
// one product per element on the way up, two on the way down (fr_mul would be twice that).
POB_HD bool vm_ginv_start(const VmCtx &x, const Op *ops, uint32_t begin, uint32_t end, uint32_t w, uint32_t nw, InvChain &c) {
    Fr acc = fr_from_u64(1); bool any = false;
    for (uint32_t i = begin + w; i < end; i += nw) {
        Fr a = vm_load(x, ops[i].a), d;
        uint64_t *vd = x.U + x.val_base + 4ull * op_dst(ops[i]);
        if (vm_inv_class(x, a, d) == 0) vm_store_val(vd, d);
        else { vm_store_val(vd, acc); acc = fr_mont(acc, a); any = true; }
    }
    if (any) inv_chain_init(c, acc);
    return any;                                   // false: nothing left to do for this worker
}
POB_HD void vm_ginv_finish(const VmCtx &x, const Op *ops, uint32_t begin, uint32_t end, uint32_t w, uint32_t nw, Fr inv) {
    uint32_t last = begin + w;
    for (uint32_t i = last; i < end; i += nw) last = i;
    for (uint32_t i = last;; i -= nw) {
        Fr a = vm_load(x, ops[i].a), d;
        if (vm_inv_class(x, a, d) != 0) {
            uint64_t *vd = x.U + x.val_base + 4ull * op_dst(ops[i]);
            Fr pre = vm_load_val(vd);
            vm_store_val(vd, fr_mont(inv, pre));
            inv = fr_mont(inv, a);
        }
        if (i < begin + w + nw) break;
    }
}

// Keccak-f round constants and rotation tables (utils/keccak.circom:195, 200, 253-262)
POB_HD uint64_t keccak_rc(int r) {
    constexpr uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    return RC[r];
}
POB_HD int keccak_rot(int i) {      // RhoPi lane walk, keccak.circom:195
    constexpr int ROT[25] = {1, 10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    return ROT[i];
}
POB_HD int keccak_shl(int i) { return ((i + 1) * (i + 2) / 2) % 64; }   // keccak.circom:200
// (b, c) operand lanes of stepChi for output lane l, keccak.circom:233-239
POB_HD int chi_b(int l) { return (l % 5 == 4) ? l - 4 : l + 1; }
POB_HD int chi_c(int l) { return (l % 5 >= 3) ? l - 3 : l + 2; }

// ---- expand: one witness code -> 32-byte little-endian field element (4 x u64) --------------------------------
POB_HD void vm_expand(Code c, const uint64_t *U, uint32_t ubase, uint32_t val_base, const Fr *konst, uint64_t out[4]) {
    uint32_t k = code_kind(c), p = code_payload(c);
    out[1] = out[2] = out[3] = 0;
    if (k == K_BIT) { out[0] = (U[ubase + (p >> 6)] >> (p & 63)) & 1ull; }
    else if (k == K_CONST) { out[0] = p; }
    else if (k == K_VAL) { const uint64_t *s = U + val_base + 4ull * p; out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; out[3] = s[3]; }
    else { const uint32_t *l = konst[p].l; for (int i = 0; i < 4; i++) out[i] = (uint64_t)l[2 * i] | ((uint64_t)l[2 * i + 1] << 32); }
}

}  // namespace pob
