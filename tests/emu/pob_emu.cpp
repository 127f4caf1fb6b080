// tests/emu/pob_emu.cpp -- TEST-ONLY host emulator of the witness VM.
//
// Runs a compiled Program (the product's layout compiler output) with the host instantiation of
// vm_exec.h so that program correctness can be checked against the oracle on a machine without a GPU.
// It is NOT part of the product: libpob_b200.so never contains or calls this file, and on a GPU box the
// parity tests go through the CUDA kernels via the C-ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "compiler.h"
#include "vm_exec.h"
#include "host_ref.h"
#include "cons_check.h"

using namespace pob;

struct EmuProgram { Program P; std::vector<Fr> invtab; std::string err; };

extern "C" {

void *pob_emu_compile(const char *main_name, const uint64_t *params, int nparams, int hcreate, char *err, int errlen) {
    try {
        std::vector<Fr> ps((size_t)nparams);
        for (int i = 0; i < nparams; i++) memcpy(ps[(size_t)i].l, params + 4 * i, 32);
        EmuProgram *e = new EmuProgram();
        e->P = compile_circuit(main_name, ps, (hcreate & 1) != 0, false, (hcreate & 0x100) ? 1 : 0);
        e->invtab = build_inverse_table();
        return e;
    } catch (const std::exception &ex) { if (err) snprintf(err, (size_t)errlen, "%s", ex.what()); return nullptr; }
}
void pob_emu_free(void *h) { delete (EmuProgram *)h; }

// stats: n_signals, n_outputs, n_inputs, n_words, n_vals, n_ops, n_absorbs, n_levels, n_tiles, n_codes, n_konst, n_round_blocks
void pob_emu_stats(void *h, uint64_t *out) {
    const Program &P = ((EmuProgram *)h)->P;
    out[0] = P.n_signals; out[1] = P.n_outputs; out[2] = P.n_inputs; out[3] = P.n_words; out[4] = P.n_vals;
    out[5] = P.ops.size(); out[6] = P.absorbs.size(); out[7] = P.levels.size(); out[8] = P.tiles.size();
    out[9] = P.codes.size(); out[10] = P.konst.size(); out[11] = P.n_round_blocks;
}
// reduced program: map[k] = --O0 index of reduced entry k; returns the --O0 signal count
uint64_t pob_emu_witness_map(void *h, uint32_t *map) {
    const Program &P = ((EmuProgram *)h)->P;
    if (map) memcpy(map, P.witness_map.data(), P.witness_map.size() * 4);
    return P.n_signals_o0;
}
const char *pob_emu_schema(void *h) { return ((EmuProgram *)h)->P.input_schema.c_str(); }

// inputs: n_inputs x 4 u64; witness: n_signals x 4 u64 (caller allocated) or NULL to skip expansion;
// outputs: n_outputs x 4 u64.  Returns status (0 = accepted, else 1 + failing component base).
uint64_t pob_emu_run(void *h, const uint64_t *inputs, uint64_t *witness, uint64_t *outputs) {
    EmuProgram *e = (EmuProgram *)h; const Program &P = e->P;
    std::vector<uint64_t> U(P.store_u64() + 4, 0);
    for (uint32_t i = 0; i < P.n_inputs; i++) {         // same input reduction as k_eval
        Fr v = vm_load_val(inputs + 4ull * i);
        while (fr_geq_p(v)) { Fr t; fr_raw_sub(t, v, fr_p()); v = t; }
        vm_store_val(U.data() + P.val_base + 4ull * i, v);
    }
    uint32_t status = STATUS_OK;
    VmCtx x{U.data(), P.val_base, P.konst.data(), P.aux.data(), e->invtab.data(), &status};
    // deferred inverses: the same start / step / finish schedule as k_eval, with 8 workers and 64 steps per level
    const uint32_t NW = 8, STEPS = 64;
    std::vector<InvChain> chain(NW); std::vector<uint32_t> phase(NW, 0);
    for (uint32_t li = 0; li < P.levels.size(); li++) {
        const Level &lv = P.levels[li];
        for (uint32_t i = lv.t_begin; i < lv.t_end; i++) vm_exec_op(x, P.ops[i]);
        for (uint32_t i = lv.w_begin; i < lv.w_end; i++) vm_absorb_scalar(U.data(), P.absorbs[i]);
        for (uint32_t i = lv.p_begin; i < lv.p_end; i++) vm_poseidon_scalar(x, P.poseidons[i], P.pos_konst.data());
        for (uint32_t i = lv.s_begin; i < lv.s_end; i++) vm_psum_scalar(x, P.psums[i]);
        for (uint32_t w = 0; w < NW; w++) {
            if (li == P.ginv_level) phase[w] = vm_ginv_start(x, P.ops.data(), P.ginv_begin, P.inv_end, w, NW, chain[w]) ? 1 : 0;
            else if (phase[w] == 1 && inv_chain_steps(chain[w], STEPS)) {
                vm_ginv_finish(x, P.ops.data(), P.ginv_begin, P.inv_end, w, NW, inv_chain_result(chain[w])); phase[w] = 0;
            }
        }
    }
    for (uint32_t w = 0; w < NW; w++) {
        if (P.ginv_level >= P.levels.size()) vm_inv_batch(x, P.ops.data(), P.ginv_begin, P.inv_end, w, NW);
        else if (phase[w] == 1) {
            while (!inv_chain_steps(chain[w], 64)) { }
            vm_ginv_finish(x, P.ops.data(), P.ginv_begin, P.inv_end, w, NW, inv_chain_result(chain[w]));
        }
    }
    if (witness)
        for (const Tile &t : P.tiles)
            for (uint32_t k = 0; k < t.n; k++) {
                uint64_t *o = witness + 4 * (t.dst + k);
                if (t.pad) {            // KeccakfRound tile: 64-signal group descriptors (same decode as k_expand_round)
                    uint64_t d = P.round_desc[(t.code_off >> 6) + (k >> 6)];
                    uint32_t tt = k & 63, mode = (uint32_t)(d >> 48), w = (uint32_t)(d & 0xffff), b = tt;
                    if (mode) { uint32_t sidx = (mode - 1) * 64 + tt, g = sidx / 3, m = sidx - 3 * g; b = g; w = (uint32_t)((d >> (16 * m)) & 0xffff); }
                    o[0] = (U[t.ubase + w] >> b) & 1ull; o[1] = o[2] = o[3] = 0;
                } else vm_expand(P.codes[t.code_off + k], U.data(), t.ubase, P.val_base, P.konst.data(), o);
            }
    if (outputs)
        for (uint32_t i = 0; i < P.n_outputs; i++)
            vm_expand(P.codes[P.out_code_off + i], U.data(), 0, P.val_base, P.konst.data(), outputs + 4 * i);
    return status == STATUS_OK ? 0 : status;
}
}

// per-level detail for tuning: out[24*i + opc] = thread ops of that opcode (opc < 16), [16] absorbs, [17] poseidon segments,
// [18] prefix sums, [19] longest prefix sum, [20] sum of prefix-sum lengths, [21] sum of poseidon t
extern "C" uint32_t pob_emu_level_detail(void *h, uint32_t *out, uint32_t max_levels) {
    const Program &P = ((EmuProgram *)h)->P;
    uint32_t n = (uint32_t)P.levels.size();
    for (uint32_t i = 0; i < n && i < max_levels; i++) {
        const Level &L = P.levels[i];
        uint32_t *o = out + 24 * i;
        for (int k = 0; k < 24; k++) o[k] = 0;
        for (uint32_t k = L.t_begin; k < L.t_end; k++) { uint32_t opc = op_opc(P.ops[k]); if (opc < 16) o[opc]++; }
        o[16] = L.w_end - L.w_begin; o[17] = L.p_end - L.p_begin; o[18] = L.s_end - L.s_begin;
        for (uint32_t k = L.s_begin; k < L.s_end; k++) { o[19] = std::max(o[19], P.psums[k].n); o[20] += P.psums[k].n; }
        for (uint32_t k = L.p_begin; k < L.p_end; k++) o[21] += P.poseidons[k].t;
    }
    return n;
}

// level histogram for tuning: out[3*i+0..2] = thread ops, absorb ops, INV ops of level i (up to max_levels)
extern "C" uint32_t pob_emu_level_hist(void *h, uint32_t *out, uint32_t max_levels) {
    const Program &P = ((EmuProgram *)h)->P;
    uint32_t n = (uint32_t)P.levels.size();
    for (uint32_t i = 0; i < n && i < max_levels; i++) {
        const Level &L = P.levels[i];
        uint32_t inv = 0;
        for (uint32_t k = L.t_begin; k < L.t_end; k++) if (op_opc(P.ops[k]) == OP_INV) inv++;
        out[3 * i] = L.t_end - L.t_begin; out[3 * i + 1] = L.w_end - L.w_begin; out[3 * i + 2] = inv;
    }
    return n;
}

// self-test of the binary-EEA inversion against the Fermat ladder on n pseudo-random field elements; returns mismatches
extern "C" uint32_t pob_emu_inv_selftest(uint32_t n) {
    uint32_t bad = 0; uint64_t st = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < n; i++) {
        Fr a;
        for (int k = 0; k < 8; k++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; a.l[k] = (uint32_t)(st >> 16); }
        a.l[7] &= 0x0fffffffu;                                       // < 2^252 < p
        if (i % 7 == 0) { a = fr_from_u64(i + 1); }                  // small values too
        if (i % 11 == 0) { Fr t; fr_raw_sub(t, fr_p(), fr_from_u64(i + 1)); a = t; }   // p - small
        if (fr_is_zero(a)) continue;
        Fr x = fr_inv_eea(a), y = fr_inv(a);
        if (!fr_eq(x, y) || !fr_eq(fr_mul(x, a), fr_from_u64(1))) bad++;
        // the deferred inverses' inversion (Kaliski almost-inverse, vm_exec.h) cut into slices: any slicing gives the Fermat result
        InvChain c; inv_chain_init(c, a);
        uint32_t slices = 0;
        while (!inv_chain_steps(c, 1 + (i % 37))) slices++;
        if (c.k < 254 || c.k > 508 || !fr_eq(inv_chain_result(c), y)) bad++;
    }
    return bad;
}

// histogram of code kinds in the code tiles (tuning aid): out[0..3] = CONST, BIT, VAL, KONST entries
extern "C" void pob_emu_code_hist(void *h, uint64_t *out) {
    const Program &P = ((EmuProgram *)h)->P;
    out[0] = out[1] = out[2] = out[3] = 0;
    for (const Tile &t : P.tiles) if (!t.pad) for (uint32_t k = 0; k < t.n; k++) out[code_kind(P.codes[t.code_off + k])]++;
}

// self-test of fr_mont / fr_mul against 512-bit schoolbook arithmetic done with unsigned __int128 on the host
extern "C" uint32_t pob_emu_mul_selftest(uint32_t n) {
    typedef unsigned __int128 u128;
    uint32_t bad = 0; uint64_t st = 0x2545F4914F6CDD1Dull;
    auto rnd = [&](Fr &a, uint32_t i) {
        for (int k = 0; k < 8; k++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; a.l[k] = (uint32_t)(st >> 11); }
        a.l[7] &= 0x0fffffffu;
        if (i % 5 == 0) { Fr t; fr_raw_sub(t, fr_p(), fr_from_u64(i + 1)); a = t; }      // p - small
        if (i % 9 == 0) a = fr_from_u64(st);
    };
    // reference: (a*b) mod p by shift-and-add
    auto mulmod = [&](const Fr &a, const Fr &b) { Fr r = fr_zero(); for (int i = 255; i >= 0; i--) { r = fr_add(r, r); if (fr_bit(b, (unsigned)i)) r = fr_add(r, a); } return r; };
    for (uint32_t i = 0; i < n; i++) {
        Fr a, b; rnd(a, i); rnd(b, i + 3);
        Fr want = mulmod(a, b);
        if (!fr_eq(fr_mul(a, b), want)) bad++;
        if (!fr_eq(fr_from_mont(fr_mont(fr_to_mont(a), fr_to_mont(b))), want)) bad++;
    }
    (void)sizeof(u128);
    return bad;
}

// ---- constraint system (csrc/cons_check.h, host instantiation): compile with constraints and evaluate every record against
// a witness supplied by the caller (the oracle's).  out[0] = circuit constraints evaluated, out[1] = hint records evaluated,
// out[2] = failing circuit constraints, out[3] = failing hints, out[4] = id of the first failing record (~0 if none),
// out[5..7] = flat eq / kc / r1 counts, out[8..10] = per-round-block eq / kc / r1 counts, out[11] = round blocks,
// out[12] = distinct witness entries referenced by at least one record.
extern "C" int pob_emu_check_constraints(const char *main_name, const uint64_t *params, int nparams, int hcreate,
                                         const uint64_t *witness, uint64_t n_signals, uint64_t *out, char *err, int errlen) {
    try {
        std::vector<Fr> ps((size_t)nparams);
        for (int i = 0; i < nparams; i++) memcpy(ps[(size_t)i].l, params + 4 * i, 32);
        Program P = compile_circuit(main_name, ps, hcreate != 0, true);
        if (P.n_signals != n_signals) throw std::runtime_error("witness size mismatch");
        uint64_t n_cons = 0, n_hint = 0, bad = 0, hbad = 0, first = ~0ull, id = 0;
        std::vector<uint8_t> seen(n_signals, 0);
        auto run = [&](const ConsSet &S, uint64_t base, uint64_t rc) {
            auto mark = [&](uint32_t idx) { if (idx != CONS_ONE) seen[base + idx] = 1; };
            for (size_t i = 0; i + 1 < S.eq.size(); i += 2, id++) { n_cons++; mark(S.eq[i]); mark(S.eq[i + 1]); if (!cons_eq_ok(witness, base, S.eq[i], S.eq[i + 1])) { bad++; if (first == ~0ull) first = id; } }
            for (const ConsTerm &t : S.kc) { n_cons++; mark(t.idx); if (!cons_kc_ok(witness, base, t, P.cons_konst.data(), rc)) { bad++; if (first == ~0ull) first = id; } id++; }
            for (const ConsR1 &r : S.r1) {
                for (uint32_t k = 0; k < (uint32_t)r.na + r.nb + r1_nc(r); k++) mark(S.terms[r.off + k].idx);
                const bool ok = cons_r1_ok(witness, base, r, S.terms.data(), P.cons_konst.data());
                if (r1_hint(r)) { n_hint++; if (!ok) { hbad++; if (first == ~0ull) first = id; } } else { n_cons++; if (!ok) { bad++; if (first == ~0ull) first = id; } }
                id++;
            }
        };
        run(P.cons_flat, 0, 0);
        for (size_t b = 0; b < P.round_block_sig.size(); b++) run(P.cons_round, P.round_block_sig[b], keccak_rc((int)(b % 24)));
        out[0] = n_cons; out[1] = n_hint; out[2] = bad; out[3] = hbad; out[4] = first;
        out[5] = P.cons_flat.eq.size() / 2; out[6] = P.cons_flat.kc.size(); out[7] = P.cons_flat.r1.size();
        out[8] = P.cons_round.eq.size() / 2; out[9] = P.cons_round.kc.size(); out[10] = P.cons_round.r1.size(); out[11] = P.round_block_sig.size();
        uint64_t ns = 0; for (uint8_t v : seen) ns += v; out[12] = ns;
        return 0;
    } catch (const std::exception &ex) { if (err) snprintf(err, (size_t)errlen, "%s", ex.what()); return -1; }
}

// tuning aid: how many IsZero inverses of the "table" group miss the small-value table on this input (they pay a real inversion)
extern "C" void pob_emu_inv_miss_count(void *h, const uint64_t *inputs, uint64_t *out) {
    EmuProgram *e = (EmuProgram *)h; const Program &P = e->P;
    std::vector<uint64_t> U(P.store_u64() + 4, 0);
    for (uint32_t i = 0; i < P.n_inputs; i++) vm_store_val(U.data() + P.val_base + 4ull * i, vm_load_val(inputs + 4ull * i));
    uint32_t status = STATUS_OK;
    VmCtx x{U.data(), P.val_base, P.konst.data(), P.aux.data(), e->invtab.data(), &status};
    for (const Level &lv : P.levels) {
        for (uint32_t i = lv.t_begin; i < lv.t_end; i++) vm_exec_op(x, P.ops[i]);
        for (uint32_t i = lv.w_begin; i < lv.w_end; i++) vm_absorb_scalar(U.data(), P.absorbs[i]);
        for (uint32_t i = lv.p_begin; i < lv.p_end; i++) vm_poseidon_scalar(x, P.poseidons[i], P.pos_konst.data());
        for (uint32_t i = lv.s_begin; i < lv.s_end; i++) vm_psum_scalar(x, P.psums[i]);
    }
    // out: 0 leveled INV ops, 1 deferred INV ops, 2 leveled ones that miss the table, 3 deferred ones that would have hit it,
    //      4 first level with every deferred input ready, 5 number of levels
    out[0] = out[2] = out[3] = 0; out[1] = P.inv_end - P.ginv_begin;
    for (uint32_t i = 0; i < P.inv_end; i++) {
        if (op_opc(P.ops[i]) != OP_INV) continue;
        Fr a = vm_load(x, P.ops[i].a), d;
        const bool miss = vm_inv_class(x, a, d) != 0;
        if (i < P.ginv_begin) { out[0]++; if (miss) out[2]++; } else if (!miss) out[3]++;
    }
    out[4] = P.ginv_level; out[5] = P.levels.size();
}
