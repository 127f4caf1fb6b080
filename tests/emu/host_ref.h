// tests/emu/host_ref.h -- TEST-ONLY scalar references of the witness VM's warp ops (Keccak absorb, Poseidon permutation,
// prefix sum).  The product implements these as warp-cooperative CUDA device functions in
// proof-of-burn_b200/csrc/pob_b200.cu (absorb_warp, poseidon_warp, psum_warp); the versions here exist so that the
// host emulator can run complete compiled programs on a machine without a GPU.  Both must fill the store layouts
// documented in csrc/program.h -- which is exactly what the GPU-vs-oracle and emulator-vs-oracle tests cross-check.
#pragma once
#include "vm_exec.h"

namespace pob {

// Scalar reference of the absorb warp op (HOST ONLY users: tests/emu).  The device implementation is the
// warp-cooperative absorb_warp() in kernels.cu; both must write the word layout documented in program.h.
inline void vm_absorb_scalar(uint64_t *W, const AbsorbOp &op) {
    uint64_t st[25];
    for (int l = 0; l < 25; l++) {
        uint64_t s = op.s_idx == NONE_IDX ? 0 : W[op.s_idx + l];
        st[l] = l < 17 ? (s ^ W[op.blk_idx + l]) : s;
        W[op.out_idx + l] = st[l];
    }
    for (int r = 0; r < 24; r++) {
        uint64_t *B = W + op.out_idx + RW * r;
        uint64_t c[5], d[5], th[25], rp[25], ch[25];
        for (int i = 0; i < 5; i++) {
            uint64_t v = st[i] ^ st[5 + i]; B[rw_x5(i, 0)] = v;
            v ^= st[10 + i]; B[rw_x5(i, 1)] = v;
            v ^= st[15 + i]; B[rw_x5(i, 2)] = v;
            v ^= st[20 + i]; B[rw_x5(i, 3)] = v; c[i] = v;
        }
        for (int i = 0; i < 5; i++) {
            uint64_t a = c[(i + 1) % 5], b = c[(i + 4) % 5], s0 = a << 1, s1 = a >> 63, o = s0 | s1;
            d[i] = b ^ o;
            B[rw_dd(i, 0)] = s0; B[rw_dd(i, 1)] = s1; B[rw_dd(i, 2)] = o; B[rw_dd(i, 3)] = d[i];
        }
        for (int l = 0; l < 25; l++) { th[l] = st[l] ^ d[l % 5]; B[rw_th(l)] = th[l]; }
        rp[0] = th[0];
        for (int i = 0; i < 24; i++) {
            int shl = keccak_shl(i); uint64_t a = th[keccak_rot(i)], a0 = a >> (64 - shl), a1 = a << shl, o = a0 | a1;
            B[rw_rp(i, 0)] = a0; B[rw_rp(i, 1)] = a1; B[rw_rp(i, 2)] = o; rp[keccak_rot(i + 1)] = o;
        }
        for (int l = 0; l < 25; l++) {
            uint64_t nb = ~rp[chi_b(l)], bc = nb & rp[chi_c(l)]; ch[l] = rp[l] ^ bc;
            B[rw_ch(l, 0)] = nb; B[rw_ch(l, 1)] = bc; B[rw_ch(l, 2)] = ch[l];
        }
        B[RW_RC] = keccak_rc(r);
        ch[0] ^= keccak_rc(r);
        for (int l = 0; l < 25; l++) { B[rw_out(l)] = ch[l]; st[l] = ch[l]; }
    }
}

// Scalar reference of the prefix-sum warp op (HOST ONLY users: tests/emu); device: psum_warp() in pob_b200.cu.
inline void vm_psum_scalar(const VmCtx &x, const PsumOp &op) {
    Fr acc = vm_load(x, op.x0);
    for (uint32_t k = 0; k < op.n; k++) { acc = fr_add(acc, vm_load(x, x.aux[op.aux0 + k])); vm_store_val(x.U + x.val_base + 4ull * (op.dst + k), acc); }
}

// Scalar reference of the Poseidon warp op (HOST ONLY users: tests/emu).  The device implementation is the
// warp-cooperative poseidon_warp() in pob_b200.cu; both must fill the slot layout documented in program.h.
inline void vm_poseidon_scalar(const VmCtx &x, const PoseidonOp &op, const Fr *pk) {
    if (op.q0 != 0) return;      // the product runs a permutation as POS_SEGMENTS warp ops in consecutive levels; only the last segment's
                                 // result is consumed, so the scalar reference may do the whole permutation at the first one
    const PosLayout L = pos_layout(op.t); const uint32_t t = op.t;
    const Fr *K = pk + op.koff;
    auto C = [&](uint32_t i) { return fr_from_mont(K[L.kC + i]); };
    auto S = [&](uint32_t i) { return fr_from_mont(K[L.kS + i]); };
    auto put = [&](uint32_t off, const Fr &v) { vm_store_val(x.U + x.val_base + 4ull * (op.base + off), v); };
    Fr st[5], y[5];
    for (uint32_t j = 0; j < t; j++) { st[j] = fr_add(vm_load(x, x.aux[op.in_aux + j]), C(j)); put(j, st[j]); }
    auto full = [&](uint32_t F, uint32_t coff, uint32_t moff) {
        for (uint32_t j = 0; j < t; j++) {
            Fr x2 = fr_mul(st[j], st[j]), x4 = fr_mul(x2, x2), x5 = fr_mul(x4, st[j]);
            put(F + 3 * j, x2); put(F + 3 * j + 1, x4); put(F + 3 * j + 2, x5);
            y[j] = fr_add(x5, C(coff + j)); put(F + 3 * t + j, y[j]);
        }
        for (uint32_t i = 0; i < t; i++) {
            Fr acc = fr_zero(); for (uint32_t j = 0; j < t; j++) acc = fr_add(acc, fr_mul(fr_from_mont(K[moff + j * t + i]), y[j]));
            put(F + 4 * t + i, acc);
        }
        for (uint32_t i = 0; i < t; i++) st[i] = vm_load_val(x.U + x.val_base + 4ull * (op.base + F + 4 * t + i));
    };
    for (uint32_t f = 0; f < 4; f++) full(L.F1 + 5 * t * f, (f + 1) * t, f == 3 ? L.kP : L.kM);
    for (uint32_t r = 0; r < L.rp; r++) {
        const uint32_t B = L.PB + r * (4 + t);
        Fr x2 = fr_mul(st[0], st[0]), x4 = fr_mul(x2, x2), x5 = fr_mul(x4, st[0]);
        put(B, x2); put(B + 1, x4); put(B + 2, x5);
        Fr z0 = fr_add(x5, C(5 * t + r)); put(B + 3, z0);
        Fr o0 = fr_mul(S((2 * t - 1) * r), z0);
        for (uint32_t i = 1; i < t; i++) o0 = fr_add(o0, fr_mul(S((2 * t - 1) * r + i), st[i]));
        for (uint32_t i = 1; i < t; i++) { st[i] = fr_add(st[i], fr_mul(z0, S((2 * t - 1) * r + t + i - 1))); put(B + 4 + i, st[i]); }
        st[0] = o0; put(B + 4, o0);
    }
    for (uint32_t f = 0; f < 3; f++) full(L.SB + 5 * t * f, 5 * t + L.rp + f * t, L.kM);
    Fr out = fr_zero();
    for (uint32_t j = 0; j < t; j++) {
        Fr x2 = fr_mul(st[j], st[j]), x4 = fr_mul(x2, x2), x5 = fr_mul(x4, st[j]);
        put(L.LB + 3 * j, x2); put(L.LB + 3 * j + 1, x4); put(L.LB + 3 * j + 2, x5);
        out = fr_add(out, fr_mul(fr_from_mont(K[L.kM + j * t]), x5));
    }
    put(L.LB + 3 * t, out);
}

}  // namespace pob
