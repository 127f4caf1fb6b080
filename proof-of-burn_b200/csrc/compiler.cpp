// compiler.cpp -- layout compiler for the proof-of-burn circuits (see compiler.h / program.h).
//
// One C++ function per circom template of the include closure of circuits/main_proof_of_burn.circom:27 and
// circuits/main_spend.circom:6.  Each function (i) reserves the component's own signals in circom's --O0
// order [outputs; inputs; intermediates], (ii) instantiates its sub-components in the order circom numbers
// them, (iii) records for every signal a 32-bit CODE saying where its value comes from, emitting VM ops for
// the values that actually have to be computed.  Copies (`a <== b`) cost nothing at run time: both signals
// get the same code.  Keccak lanes are 64-bit store words whose 64 bit-signals are BIT codes.
//
// Numbering rules: SURVEY.md Appendix C.  R3 default = completion order (circom >= 2.1 instantiates a
// sub-component when its last input is assigned); `hcreate` flips Num2Bits_strict and MultiAND(n>=3).
#include "compiler.h"
#include <sys/mman.h>
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include "poseidon_constants_data.h"
#include "vm_exec.h"

namespace pob {
namespace {

struct Blk { uint64_t sig; size_t pos; };          // first own signal: witness index and flat-code position
struct Lane { uint32_t w; };                       // lane word index in the store; NONE_IDX = constant zero lane

static const Code ZERO = 0, ONE = 1;

struct KeyHash {
    size_t operator()(const std::array<uint32_t, 4> &k) const {
        uint64_t h = 0x9E3779B97F4A7C15ULL;
        for (uint32_t v : k) { h ^= v; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 29; }
        return (size_t)h;
    }
};
struct FrHash {
    size_t operator()(const std::array<uint32_t, 8> &k) const {
        uint64_t h = 0x9E3779B97F4A7C15ULL;
        for (uint32_t v : k) { h ^= v; h *= 0x94D049BB133111EBULL; h ^= h >> 31; }
        return (size_t)h;
    }
};

class Builder {
  public:
    bool hcreate, dry;                 // dry: first pass, only counts lane words
    uint32_t val_base;
    // flat codes (pointer-stable arena)
    Code *flat = nullptr; size_t flat_n = 0, flat_cap = (size_t)1 << 30;
    uint64_t nsig = 0;
    struct Seg { uint64_t dst; size_t pos; uint64_t n; bool round; uint32_t ubase; };
    std::vector<Seg> segs;
    // store
    uint32_t n_words = 0, n_vals = 0;
    std::vector<uint32_t> lvlW, lvlV;
    // program
    struct OpRec { Op op; uint32_t level; };
    std::vector<OpRec> ops;
    struct PsumRec { PsumOp op; uint32_t level; };
    std::vector<PsumRec> psums;
    std::vector<uint8_t> inv_generic;   // per value slot: 1 = an IsZero inverse whose input is expected to be a large value
    struct PosRec { PoseidonOp op; uint32_t level; };
    std::vector<PosRec> poseidons;
    std::vector<Fr> pos_konst; uint32_t pos_koff[6] = {~0u, ~0u, ~0u, ~0u, ~0u, ~0u};
    struct AbsRec { AbsorbOp op; uint32_t level; };
    std::vector<AbsRec> absorbs;
    std::vector<Code> aux;
    std::vector<Fr> konsts;
    std::unordered_map<std::array<uint32_t, 8>, uint32_t, FrHash> konst_ix;
    std::unordered_map<std::array<uint32_t, 4>, Code, KeyHash> cse;
    std::unordered_map<std::array<uint32_t, 4>, size_t, KeyHash> chk_ix;
    uint64_t n_round_blocks = 0;
    Code MINUS1;

    Builder(bool hc, bool dry_, uint32_t vb) : hcreate(hc), dry(dry_), val_base(vb) {
        flat = (Code *)mmap(nullptr, flat_cap * sizeof(Code), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (flat == MAP_FAILED) throw std::runtime_error("pob: cannot reserve code arena");
        Fr m1; Fr one = fr_from_u64(1); fr_raw_sub(m1, fr_p(), one);
        MINUS1 = konst(m1);                             // must be konst index 0: vm_exec.h keys its negate fast path on it
        if (MINUS1 != c_konst(0)) throw std::runtime_error("pob: internal: MINUS1 must be the first table constant");
        Blk b = alloc(1); flat[b.pos] = ONE;            // witness[0] = 1
    }
    ~Builder() { if (flat && flat != MAP_FAILED) munmap(flat, flat_cap * sizeof(Code)); }

    // ---- signals ----
    Blk alloc(size_t n) {
        if (segs.empty() || segs.back().round) segs.push_back({nsig, flat_n, 0, false, 0});
        if (flat_n + n > flat_cap) throw std::runtime_error("pob: code arena exhausted");
        Blk b{nsig, flat_n};
        segs.back().n += n; flat_n += n; nsig += n;
        return b;
    }
    void round_block(uint32_t ubase) {
        segs.push_back({nsig, 0, ROUND_SIGNALS, true, ubase});
        nsig += ROUND_SIGNALS; n_round_blocks++;
    }
    inline Code &at(size_t pos) { return flat[pos]; }
    void copy(size_t dst, const Code *src, size_t n) { memcpy(flat + dst, src, n * sizeof(Code)); }

    // ---- store ----
    uint32_t new_words(uint32_t n, uint32_t level) {
        uint32_t w = n_words; n_words += n;
        if (!dry && (uint64_t)n_words > val_base) throw std::runtime_error("pob: lane-word region overflow");
        lvlW.resize(n_words, level);
        return w;
    }
    uint32_t new_val(uint32_t level) { lvlV.push_back(level); return n_vals++; }
    uint32_t val_u64(uint32_t slot) const { return val_base + 4u * slot; }
    Code bit_of_val(uint32_t slot, uint32_t bit) const {
        uint32_t idx = val_u64(slot) + bit / 64;
        if (!dry && idx >= MAX_STORE_U64) throw std::runtime_error("pob: instance store exceeds 128 MiB (code range)");
        return c_bit(idx & (MAX_STORE_U64 - 1), bit % 64);
    }
    Code lane_bit(Lane l, uint32_t k) const { return l.w == NONE_IDX ? ZERO : c_bit(l.w, k); }
    void set_lane(size_t pos, Lane l) { for (uint32_t k = 0; k < 64; k++) flat[pos + k] = lane_bit(l, k); }

    uint32_t level_of(Code c) const {
        uint32_t k = code_kind(c), p = code_payload(c);
        if (k == K_VAL) return lvlV[p];
        if (k == K_BIT) { uint32_t idx = p >> 6; if (dry) return 0; return idx < val_base ? lvlW[idx] : lvlV[(idx - val_base) / 4]; }
        return 0;
    }

    // ---- constants ----
    Code konst(const Fr &v) {
        if (fr_fits64(v) && fr_lo64(v) < (1u << 30)) return c_const((uint32_t)fr_lo64(v));
        std::array<uint32_t, 8> k; memcpy(k.data(), v.l, 32);
        auto it = konst_ix.find(k);
        if (it != konst_ix.end()) return c_konst(it->second);
        uint32_t ix = (uint32_t)konsts.size(); konsts.push_back(v); konst_ix.emplace(k, ix);
        return c_konst(ix);
    }
    Code konst_u64(uint64_t v) { return konst(fr_from_u64(v)); }
    Code pow2(unsigned n) {                         // 2^n mod p
        Fr r = fr_from_u64(1); for (unsigned i = 0; i < n; i++) r = fr_add(r, r); return konst(r);
    }
    bool const_val(Code c, Fr &out) const {
        uint32_t k = code_kind(c);
        if (k == K_CONST) { out = fr_from_u64(code_payload(c)); return true; }
        if (k == K_KONST) { out = konsts[code_payload(c)]; return true; }
        return false;
    }
    bool is_zero(Code c) const { return c == ZERO; }

    // ---- value ops ----
    Code emit_val(uint32_t opc, Code a, Code b, Code c, uint32_t level) {
        std::array<uint32_t, 4> key{opc, a, b, c};
        auto it = cse.find(key);
        if (it != cse.end()) return it->second;
        uint32_t slot = new_val(level);
        if (slot >= (1u << 26)) throw std::runtime_error("pob: too many value slots");
        ops.push_back({Op{(opc << 26) | slot, a, b, c}, level});
        Code r = c_val(slot); cse.emplace(key, r); return r;
    }
    Code fma(Code a, Code b, Code c) {              // a*b + c
        Fr fa, fb, fc; bool ka = const_val(a, fa), kb = const_val(b, fb), kc = const_val(c, fc);
        if (ka && kb && kc) return konst(fr_add(fr_mul(fa, fb), fc));
        if ((ka && fr_is_zero(fa)) || (kb && fr_is_zero(fb))) return c;
        if (ka && kb) { a = konst(fr_mul(fa, fb)); b = ONE; fa = fr_mul(fa, fb); fb = fr_from_u64(1); }
        if (kc && fr_is_zero(fc)) {
            if (ka && fr_eq(fa, fr_from_u64(1))) return b;
            if (kb && fr_eq(fb, fr_from_u64(1))) return a;
        }
        if (ka && !kb) { std::swap(a, b); std::swap(ka, kb); std::swap(fa, fb); }   // keep the constant factor in `b` (VM fast paths)
        uint32_t lv = 1 + std::max(level_of(a), std::max(level_of(b), level_of(c)));
        return emit_val(OP_FMA, a, b, c, lv);
    }
    Code add(Code a, Code b) { return fma(a, ONE, b); }
    Code sub(Code a, Code b) { return fma(b, MINUS1, a); }
    Code mul(Code a, Code b) { return fma(a, b, ZERO); }
    Code not1(Code a) { return fma(a, MINUS1, ONE); }           // 1 - a
    Code isz(Code a) {
        Fr fa; if (const_val(a, fa)) return fr_is_zero(fa) ? ONE : ZERO;
        return emit_val(OP_ISZ, a, 0, 0, 1 + level_of(a));
    }
    Code inv(Code a, bool likely_large = false) {
        Fr fa; if (const_val(a, fa)) return fr_is_zero(fa) ? ZERO : konst(fr_inv(fa));
        Code r = emit_val(OP_INV, a, 0, 0, 1 + level_of(a));
        if (likely_large) { inv_generic.resize(n_vals, 0); inv_generic[code_payload(r)] = 1; }
        return r;
    }
    // (a > k) on canonical integers == prod_{j<=k} (1 - IsEqual(j, a))
    Code gtc(Code a, uint32_t k) {
        Fr fa; if (const_val(a, fa)) return (!fr_fits64(fa) || fr_lo64(fa) > k) ? ONE : ZERO;
        return emit_val(OP_GTC, a, k, 0, 1 + level_of(a));
    }
    // sum_{j<=i} IsEqual(sel, j) * vals[j]  for every i < n, as n independent ops over one shared operand list
    void selsum(Code sel, const Code *vals, size_t n, Code *out) {
        uint32_t lv = level_of(sel);
        for (size_t i = 0; i < n; i++) lv = std::max(lv, level_of(vals[i]));
        uint32_t a0 = (uint32_t)aux.size();
        aux.insert(aux.end(), vals, vals + n);
        for (size_t i = 0; i < n; i++) {
            uint32_t slot = new_val(lv + 1);
            ops.push_back({Op{(OP_SELSUM << 26) | slot, sel, a0, (uint32_t)i}, lv + 1});
            out[i] = c_val(slot);
        }
    }
    // V[k] = x0 + sum_{i<=k} terms[i], every partial sum a signal: one warp-level prefix-sum op
    void prefix_sum(Code x0, const Code *terms, size_t n, Code *out) {
        uint32_t lv = level_of(x0);
        for (size_t i = 0; i < n; i++) lv = std::max(lv, level_of(terms[i]));
        uint32_t a0 = (uint32_t)aux.size(); aux.insert(aux.end(), terms, terms + n);
        uint32_t first = n_vals;
        for (size_t i = 0; i < n; i++) { uint32_t slot = new_val(lv + 1); out[i] = c_val(slot); }
        psums.push_back({PsumOp{a0, (uint32_t)n, first, x0}, lv + 1});
    }
    // balanced sum of terms that are `var` accumulations in the circuit (only the total is a signal)
    Code sum_tree(std::vector<Code> v) {
        if (v.empty()) return ZERO;
        while (v.size() > 1) {
            std::vector<Code> nx; nx.reserve((v.size() + 1) / 2);
            for (size_t i = 0; i + 1 < v.size(); i += 2) nx.push_back(add(v[i], v[i + 1]));
            if (v.size() & 1) nx.push_back(v.back());
            v.swap(nx);
        }
        return v[0];
    }
    Code divmod(bool want_mod, Code a, Code b, uint64_t base) {
        uint32_t lv = 1 + std::max(level_of(a), level_of(b));
        return emit_val(want_mod ? OP_MOD : OP_DIV, a, b, (uint32_t)base, lv);
    }
    // ---- constraint checks ----
    void emit_chk(uint32_t opc, Code a, uint32_t b, uint64_t base, uint32_t level) {
        std::array<uint32_t, 4> key{opc, a, b, 0};
        auto it = chk_ix.find(key);
        if (it != chk_ix.end()) { Op &o = ops[it->second].op; if ((uint32_t)base < o.c) o.c = (uint32_t)base; return; }
        chk_ix.emplace(key, ops.size());
        ops.push_back({Op{opc << 26, a, b, (uint32_t)base}, level});
    }
    void chk_eq(Code a, Code b, uint64_t base) {
        Fr fa, fb;
        if (a == b) return;
        if (const_val(a, fa) && const_val(b, fb) && fr_eq(fa, fb)) return;
        emit_chk(OP_CHK_EQ, a, b, base, 1 + std::max(level_of(a), level_of(b)));
    }
    void chk_range(Code a, unsigned nbits, uint64_t base) {
        if (nbits >= 254) return;                    // every canonical value is < 2^254
        Fr fa;
        if (const_val(a, fa)) { if (!fr_lt_pow2(fa, nbits)) emit_chk(OP_CHK_EQ, ONE, ZERO, base, 1); return; }
        if (code_kind(a) == K_BIT && nbits >= 1) return;
        emit_chk(OP_CHK_RANGE, a, nbits, base, 1 + level_of(a));
    }
    // ---- lane ops ----
    uint32_t pack8_words(const Code *bytes, uint32_t nwords) {   // bytes[8*nwords] -> nwords consecutive lane words
        uint32_t lv = 0;
        for (uint32_t i = 0; i < 8 * nwords; i++) lv = std::max(lv, level_of(bytes[i]));
        uint32_t w0 = new_words(nwords, lv + 1);
        for (uint32_t w = 0; w < nwords; w++) {
            uint32_t a0 = (uint32_t)aux.size();
            for (int k = 0; k < 8; k++) aux.push_back(bytes[8 * w + k]);
            uint32_t l = 0; for (int k = 0; k < 8; k++) l = std::max(l, level_of(bytes[8 * w + k]));
            lvlW[w0 + w] = l + 1;
            ops.push_back({Op{(OP_PACK8 << 26) | (w0 + w), a0, 0, 0}, l + 1});
        }
        return w0;
    }
    // one Poseidon permutation as a warp op: returns the first slot of its value block (layout: program.h)
    uint32_t poseidon(uint32_t t, const Code *in) {
        if (t < 3 || t > 5) throw std::runtime_error("pob: Poseidon width outside this circuit's closure (t = 3, 4, 5)");
        const PosLayout L = pos_layout(t);
        if (pos_koff[t] == ~0u) {                       // C, S, M, P of this width, converted to Montgomery form once
            pos_koff[t] = (uint32_t)pos_konst.size();
            const uint64_t (*tabs[4])[4] = {t == 3 ? POSEIDON_C_T3 : t == 4 ? POSEIDON_C_T4 : POSEIDON_C_T5,
                                            t == 3 ? POSEIDON_S_T3 : t == 4 ? POSEIDON_S_T4 : POSEIDON_S_T5,
                                            t == 3 ? POSEIDON_M_T3 : t == 4 ? POSEIDON_M_T4 : POSEIDON_M_T5,
                                            t == 3 ? POSEIDON_P_T3 : t == 4 ? POSEIDON_P_T4 : POSEIDON_P_T5};
            const uint32_t cnt[4] = {t * 8 + L.rp, L.rp * (2 * t - 1), t * t, t * t};
            for (int k = 0; k < 4; k++) for (uint32_t i = 0; i < cnt[k]; i++) { Fr f; memcpy(f.l, tabs[k][i], 32); pos_konst.push_back(fr_to_mont(f)); }
        }
        uint32_t lv = 0; for (uint32_t j = 0; j < t; j++) lv = std::max(lv, level_of(in[j]));
        uint32_t a0 = (uint32_t)aux.size(); aux.insert(aux.end(), in, in + t);
        uint32_t base = n_vals;
        for (uint32_t i = 0; i < L.total; i++) new_val(lv + 1);
        poseidons.push_back({PoseidonOp{t, a0, base, pos_koff[t]}, lv + 1});
        return base;
    }
    // one Absorb: returns base word of its ABSORB_WORDS block
    uint32_t absorb(uint32_t s_idx, uint32_t blk_idx) {
        uint32_t lv = 0;
        if (s_idx != NONE_IDX) for (int l = 0; l < 25; l++) lv = std::max(lv, dry ? 0u : lvlW[s_idx + l]);
        for (int l = 0; l < 17; l++) lv = std::max(lv, dry ? 0u : lvlW[blk_idx + l]);
        uint32_t out = new_words(ABSORB_WORDS, lv + 1);
        absorbs.push_back({AbsorbOp{s_idx, blk_idx, out, 0}, lv + 1});
        return out;
    }
};

// ============================================================================================================
// circomlib/circuits/gates.circom
// ============================================================================================================
// AND :29-35 (own: out, a, b).  Scalar XOR/OR only occur inside the Keccak lane arrays (handled as lanes).
static Blk T_AND(Builder &B, Code a, Code b) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = a; B.at(o.pos + 2) = b; B.at(o.pos) = B.mul(a, b); return o;
}
// MultiAND(n) :68-96
static Blk T_MultiAND(Builder &B, int n, const Code *in) {
    Blk o = B.alloc(1 + (size_t)n); B.copy(o.pos + 1, in, (size_t)n);
    if (n == 1) B.at(o.pos) = in[0];
    else if (n == 2) { Blk a = T_AND(B, in[0], in[1]); B.at(o.pos) = B.at(a.pos); }
    else {
        int n1 = n / 2, n2 = n - n / 2;
        if (B.hcreate) {
            Blk a2 = B.alloc(3);
            Blk x0 = T_MultiAND(B, n1, in), x1 = T_MultiAND(B, n2, in + n1);
            Code u = B.at(x0.pos), v = B.at(x1.pos);
            B.at(a2.pos + 1) = u; B.at(a2.pos + 2) = v; B.at(a2.pos) = B.mul(u, v); B.at(o.pos) = B.at(a2.pos);
        } else {
            Blk x0 = T_MultiAND(B, n1, in), x1 = T_MultiAND(B, n2, in + n1);
            Blk a2 = T_AND(B, B.at(x0.pos), B.at(x1.pos)); B.at(o.pos) = B.at(a2.pos);
        }
    }
    return o;
}

// ============================================================================================================
// circomlib/circuits/bitify.circom, aliascheck.circom, compconstant.circom
// ============================================================================================================
// bit i of a value code
static Code bit_of(Builder &B, Code in, unsigned i) {
    Fr f;
    if (B.const_val(in, f)) return i < 256 ? c_const((uint32_t)fr_bit(f, i)) : ZERO;
    if (code_kind(in) == K_BIT) return i == 0 ? in : ZERO;
    return B.bit_of_val(code_payload(in), i);
}
// Num2Bits(n) :25-39  own: out[n], in
static Blk T_Num2Bits(Builder &B, int n, Code in) {
    Blk o = B.alloc((size_t)n + 1);
    for (int i = 0; i < n; i++) B.at(o.pos + (size_t)i) = bit_of(B, in, (unsigned)i);
    B.at(o.pos + (size_t)n) = in;
    B.chk_range(in, (unsigned)n, o.sig);
    return o;
}
// Bits2Num(n) :55-67  own: out, in[n]
static Blk T_Bits2Num(Builder &B, int n, const Code *in) {
    Blk o = B.alloc((size_t)n + 1); B.copy(o.pos + 1, in, (size_t)n);
    std::vector<Code> terms; for (int i = 0; i < n; i++) terms.push_back(B.mul(in[i], B.pow2((unsigned)i)));
    B.at(o.pos) = B.sum_tree(terms); return o;
}
// CompConstant(ct) :25-73 with ct = p-1  own: out, in[254], parts[127], sout ; child Num2Bits(135)
static Blk T_CompConstant(Builder &B, const Fr &ct, const Code *in) {
    Blk o = B.alloc(1 + 254 + 127 + 1); B.copy(o.pos + 1, in, 254);
    size_t parts = o.pos + 255;
    Fr one = fr_from_u64(1);
    Fr b; { Fr t = fr_from_u64(1); for (int i = 0; i < 128; i++) t = fr_add(t, t); b = fr_sub(t, one); }
    Fr a = one, e = one;
    std::vector<Code> terms;
    for (int i = 0; i < 127; i++) {
        int clsb = fr_bit(ct, (unsigned)(2 * i)), cmsb = fr_bit(ct, (unsigned)(2 * i + 1));
        Code slsb = in[2 * i], smsb = in[2 * i + 1], ml = B.mul(smsb, slsb), p;
        Code kb = B.konst(b), ka = B.konst(a), nkb = B.konst(fr_neg(b)), nka = B.konst(fr_neg(a));
        if (!cmsb && !clsb)      p = B.fma(slsb, kb, B.fma(smsb, kb, B.fma(ml, nkb, ZERO)));
        else if (!cmsb && clsb)  p = B.fma(smsb, nka, B.fma(smsb, kb, B.fma(slsb, nka, B.fma(ml, ka, ka))));
        else if (cmsb && !clsb)  p = B.fma(smsb, nka, B.fma(ml, kb, ka));
        else                     p = B.fma(ml, nka, ka);
        B.at(parts + (size_t)i) = p; terms.push_back(p);
        b = fr_sub(b, e); a = fr_add(a, e); e = fr_add(e, e);
    }
    Code sum = B.sum_tree(terms);
    B.at(o.pos + 255 + 127) = sum;
    Blk nb = T_Num2Bits(B, 135, sum);
    B.at(o.pos) = B.at(nb.pos + 127);
    return o;
}
// AliasCheck :24-32  own: in[254]
static Blk T_AliasCheck(Builder &B, const Code *in) {
    Blk o = B.alloc(254); B.copy(o.pos, in, 254);
    Fr m1; Fr one = fr_from_u64(1); fr_raw_sub(m1, fr_p(), one);
    Blk cc = T_CompConstant(B, m1, in);
    B.chk_eq(B.at(cc.pos), ZERO, o.sig);
    return o;
}
// Num2Bits_strict :41-53  own: out[254], in
static Blk T_Num2Bits_strict(Builder &B, Code in) {
    Blk o = B.alloc(255); B.at(o.pos + 254) = in;
    if (B.hcreate) {
        std::vector<Code> bits(254); for (int i = 0; i < 254; i++) bits[i] = bit_of(B, in, (unsigned)i);
        T_AliasCheck(B, bits.data());
        Blk nb = T_Num2Bits(B, 254, in); B.copy(o.pos, &B.at(nb.pos), 254);
    } else {
        Blk nb = T_Num2Bits(B, 254, in); B.copy(o.pos, &B.at(nb.pos), 254);
        T_AliasCheck(B, &B.at(nb.pos));
    }
    return o;
}

// ============================================================================================================
// circomlib/circuits/comparators.circom, mux1.circom
// ============================================================================================================
// IsZero :24-35  own: out, in, inv
static Blk T_IsZero(Builder &B, Code in, bool likely_large = false) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = in; B.at(o.pos + 2) = B.inv(in, likely_large); B.at(o.pos) = B.isz(in); return o;
}
// IsEqual :37-46  own: out, in[2] ; isz.in = in[1] - in[0]
static Blk T_IsEqual(Builder &B, Code in0, Code in1, bool likely_large = false) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = in0; B.at(o.pos + 2) = in1;
    Blk z = T_IsZero(B, B.sub(in1, in0), likely_large); B.at(o.pos) = B.at(z.pos); return o;
}
// LessThan(n) :89-100
static Blk T_LessThan(Builder &B, int n, Code in0, Code in1) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = in0; B.at(o.pos + 2) = in1;
    Blk nb = T_Num2Bits(B, n + 1, B.sub(B.add(in0, B.pow2((unsigned)n)), in1));
    B.at(o.pos) = B.not1(B.at(nb.pos + (size_t)n)); return o;
}
// LessEqThan(n) :105-115
static Blk T_LessEqThan(Builder &B, int n, Code in0, Code in1) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = in0; B.at(o.pos + 2) = in1;
    Blk lt = T_LessThan(B, n, in0, B.add(in1, ONE)); B.at(o.pos) = B.at(lt.pos); return o;
}
// GreaterEqThan(n) :131-141
static Blk T_GreaterEqThan(Builder &B, int n, Code in0, Code in1) {
    Blk o = B.alloc(3); B.at(o.pos + 1) = in0; B.at(o.pos + 2) = in1;
    Blk lt = T_LessThan(B, n, in1, B.add(in0, ONE)); B.at(o.pos) = B.at(lt.pos); return o;
}
// Mux1 :34-48 + MultiMux1(1) :21-32
static Blk T_Mux1(Builder &B, Code c0, Code c1, Code s) {
    Blk o = B.alloc(4); B.at(o.pos + 1) = c0; B.at(o.pos + 2) = c1; B.at(o.pos + 3) = s;
    Blk m = B.alloc(4); B.at(m.pos + 1) = c0; B.at(m.pos + 2) = c1; B.at(m.pos + 3) = s;
    B.at(m.pos) = B.fma(B.sub(c1, c0), s, c0);
    B.at(o.pos) = B.at(m.pos); return o;
}

// ============================================================================================================
// circomlib/circuits/poseidon.circom
// ============================================================================================================
// PoseidonEx(nInputs,1) :67-196  own: out[1], inputs[n], initialState.  The arithmetic is one warp op (Builder::poseidon);
// here only the circom numbering of its ~1100 signals is laid out over the op's value block.
static Blk T_PoseidonEx(Builder &B, int nInputs, const Code *inputs, Code initialState) {
    const uint32_t t = (uint32_t)nInputs + 1; const PosLayout L = pos_layout(t);
    Blk o = B.alloc(2 + (size_t)nInputs); B.copy(o.pos + 1, inputs, (size_t)nInputs); B.at(o.pos + 1 + (size_t)nInputs) = initialState;
    Code st[8], cur[8];
    st[0] = initialState; for (uint32_t j = 1; j < t; j++) st[j] = inputs[j - 1];
    const uint32_t base = B.poseidon(t, st);
    auto V = [&](uint32_t off) { return c_val(base + off); };
    { Blk a = B.alloc(2 * t); for (uint32_t j = 0; j < t; j++) { B.at(a.pos + j) = V(j); B.at(a.pos + t + j) = st[j]; cur[j] = V(j); } }   // ark[0]
    auto sigma = [&](Code in, uint32_t off) {            // Sigma :5-16  own: out, in, in2, in4
        Blk s = B.alloc(4); B.at(s.pos) = V(off + 2); B.at(s.pos + 1) = in; B.at(s.pos + 2) = V(off); B.at(s.pos + 3) = V(off + 1);
    };
    auto full = [&](uint32_t F) {                        // t x Sigma, Ark :18-25, Mix :27-39
        for (uint32_t j = 0; j < t; j++) sigma(cur[j], F + 3 * j);
        Blk a = B.alloc(2 * t); for (uint32_t j = 0; j < t; j++) { B.at(a.pos + j) = V(F + 3 * t + j); B.at(a.pos + t + j) = V(F + 3 * j + 2); }
        Blk m = B.alloc(2 * t); for (uint32_t j = 0; j < t; j++) { B.at(m.pos + j) = V(F + 4 * t + j); B.at(m.pos + t + j) = V(F + 3 * t + j); cur[j] = V(F + 4 * t + j); }
    };
    for (uint32_t f = 0; f < 4; f++) full(L.F1 + 5 * t * f);                        // :101-136 (the 4th mixes with P)
    for (uint32_t r = 0; r < L.rp; r++) {                                            // :138-160
        const uint32_t Bs = L.PB + r * (4 + t);
        sigma(cur[0], Bs);
        Blk m = B.alloc(2 * t);                                                      // MixS :52-65  own: out[t], in[t]
        for (uint32_t j = 0; j < t; j++) { B.at(m.pos + j) = V(Bs + 4 + j); B.at(m.pos + t + j) = j == 0 ? V(Bs + 3) : cur[j]; }
        for (uint32_t j = 0; j < t; j++) cur[j] = V(Bs + 4 + j);
    }
    for (uint32_t f = 0; f < 3; f++) full(L.SB + 5 * t * f);                        // :162-182
    for (uint32_t j = 0; j < t; j++) sigma(cur[j], L.LB + 3 * j);                    // :184-187
    Blk ml = B.alloc(1 + t);                                                         // MixLast :41-50  own: out, in[t]
    B.at(ml.pos) = V(L.LB + 3 * t); for (uint32_t j = 0; j < t; j++) B.at(ml.pos + 1 + j) = V(L.LB + 3 * j + 2);
    B.at(o.pos) = V(L.LB + 3 * t); return o;
}
// Poseidon(n) :198-208
static Blk T_Poseidon(Builder &B, int n, const Code *inputs) {
    Blk o = B.alloc(1 + (size_t)n); B.copy(o.pos + 1, inputs, (size_t)n);
    Blk e = T_PoseidonEx(B, n, inputs, ZERO); B.at(o.pos) = B.at(e.pos); return o;
}

// ============================================================================================================
// circuits/utils/assert.circom
// ============================================================================================================
// AssertBits(B) :13-18  own: in, bits[B]
static Blk T_AssertBits(Builder &B, int nb_, Code in) {
    Blk o = B.alloc(1 + (size_t)nb_); B.at(o.pos) = in;
    Blk nb = T_Num2Bits(B, nb_, in); B.copy(o.pos + 1, &B.at(nb.pos), (size_t)nb_); return o;
}
// AssertByteString(N) :26-31
static Blk T_AssertByteString(Builder &B, int N, const Code *in) {
    Blk o = B.alloc((size_t)N); B.copy(o.pos, in, (size_t)N);
    for (int i = 0; i < N; i++) T_AssertBits(B, 8, in[i]);
    return o;
}
// AssertLessThan :40-47 / AssertLessEqThan :56-63 / AssertGreaterEqThan :72-79  own: a, b, out
static Blk T_AssertCmp(Builder &B, int kind, int nb, Code a, Code b) {
    Blk o = B.alloc(3); B.at(o.pos) = a; B.at(o.pos + 1) = b;
    T_AssertBits(B, nb, a); T_AssertBits(B, nb, b);
    Blk r = kind == 0 ? T_LessThan(B, nb, a, b) : kind == 1 ? T_LessEqThan(B, nb, a, b) : T_GreaterEqThan(B, nb, a, b);
    B.at(o.pos + 2) = B.at(r.pos);
    B.chk_eq(B.at(r.pos), ONE, o.sig);
    return o;
}
static Blk T_AssertLessThan(Builder &B, int nb, Code a, Code b) { return T_AssertCmp(B, 0, nb, a, b); }
static Blk T_AssertLessEqThan(Builder &B, int nb, Code a, Code b) { return T_AssertCmp(B, 1, nb, a, b); }
static Blk T_AssertGreaterEqThan(Builder &B, int nb, Code a, Code b) { return T_AssertCmp(B, 2, nb, a, b); }

// ============================================================================================================
// circuits/utils/array.circom
// ============================================================================================================
// Filter(N) :26-40  own: out[N], in, isEq[N]
static Blk T_Filter(Builder &B, int N, Code in) {
    size_t n = (size_t)N; Blk o = B.alloc(2 * n + 1); B.at(o.pos + n) = in;
    for (size_t i = 0; i < n; i++) {
        Blk e = T_IsEqual(B, c_const((uint32_t)i), in);
        Code eq = B.at(e.pos); B.at(o.pos + n + 1 + i) = eq;
        B.at(o.pos + i) = B.gtc(in, (uint32_t)i);      // prod_{j<=i} (1 - isEq[j]) == (in > i)
    }
    return o;
}
// Fit(M,N) :47-57  own: out[N], in[M]
static Blk T_Fit(Builder &B, int M, int N, const Code *in) {
    Blk o = B.alloc((size_t)N + (size_t)M); B.copy(o.pos + (size_t)N, in, (size_t)M);
    for (int i = 0; i < N; i++) B.at(o.pos + (size_t)i) = i < M ? in[i] : ZERO;
    return o;
}
// Flatten(M,N) :64-72 / Reshape(M,N) :79-87: identity on row-major data  own: out[n], in[n]
static Blk T_CopyArray(Builder &B, size_t n, const Code *in) {
    Blk o = B.alloc(2 * n); B.copy(o.pos, in, n); B.copy(o.pos + n, in, n); return o;
}
// Reverse(N) :94-99
static Blk T_Reverse(Builder &B, int N, const Code *in) {
    size_t n = (size_t)N; Blk o = B.alloc(2 * n); B.copy(o.pos + n, in, n);
    for (size_t i = 0; i < n; i++) B.at(o.pos + i) = in[n - 1 - i];
    return o;
}

// ============================================================================================================
// circuits/utils/convert.circom
// ============================================================================================================
// LittleEndianBytes2Num(N) :12-26  own: out, in[N]
static Blk T_LittleEndianBytes2Num(Builder &B, int N, const Code *in) {
    Blk o = B.alloc(1 + (size_t)N); B.copy(o.pos + 1, in, (size_t)N);
    T_AssertByteString(B, N, in);
    std::vector<Code> terms; for (int i = 0; i < N; i++) terms.push_back(B.mul(in[i], B.pow2((unsigned)(8 * i))));
    B.at(o.pos) = B.sum_tree(terms); return o;
}
// BigEndianBytes2Num(N) :33-39  own: out, in[N], inReversed[N]
static Blk T_BigEndianBytes2Num(Builder &B, int N, const Code *in) {
    size_t n = (size_t)N; Blk o = B.alloc(1 + 2 * n); B.copy(o.pos + 1, in, n);
    Blk r = T_Reverse(B, N, in); B.copy(o.pos + 1 + n, &B.at(r.pos), n);
    Blk l = T_LittleEndianBytes2Num(B, N, &B.at(r.pos)); B.at(o.pos) = B.at(l.pos); return o;
}
// Num2BitsSafe(N) :46-56
static Blk T_Num2BitsSafe(Builder &B, int N, Code in) {
    size_t n = (size_t)N;
    if (N >= 254) {
        Blk o = B.alloc(n + 1 + 254); B.at(o.pos + n) = in;
        Blk st = T_Num2Bits_strict(B, in); B.copy(o.pos + n + 1, &B.at(st.pos), 254);
        Blk f = T_Fit(B, 254, N, &B.at(st.pos)); B.copy(o.pos, &B.at(f.pos), n); return o;
    }
    Blk o = B.alloc(n + 1); B.at(o.pos + n) = in;
    Blk nb = T_Num2Bits(B, N, in); B.copy(o.pos, &B.at(nb.pos), n); return o;
}
// Num2LittleEndianBytes(N) :69-82  own: out[N], in, bits[8N], byteArrays[N][8]
static Blk T_Num2LittleEndianBytes(Builder &B, int N, Code in) {
    size_t n = (size_t)N; Blk o = B.alloc(n + 1 + 16 * n); B.at(o.pos + n) = in;
    Blk b = T_Num2BitsSafe(B, 8 * N, in); B.copy(o.pos + n + 1, &B.at(b.pos), 8 * n);
    Blk r = T_CopyArray(B, 8 * n, &B.at(b.pos)); B.copy(o.pos + n + 1 + 8 * n, &B.at(r.pos), 8 * n);
    for (size_t i = 0; i < n; i++) { Blk bn = T_Bits2Num(B, 8, &B.at(r.pos + 8 * i)); B.at(o.pos + i) = B.at(bn.pos); }
    return o;
}
// Num2BigEndianBytes(N) :90-96  own: out[N], in, littleEndian[N]
static Blk T_Num2BigEndianBytes(Builder &B, int N, Code in) {
    size_t n = (size_t)N; Blk o = B.alloc(2 * n + 1); B.at(o.pos + n) = in;
    Blk le = T_Num2LittleEndianBytes(B, N, in); B.copy(o.pos + n + 1, &B.at(le.pos), n);
    Blk rv = T_Reverse(B, N, &B.at(le.pos)); B.copy(o.pos, &B.at(rv.pos), n); return o;
}
// Bytes2Nibbles(N) :103-125  own: out[2N], in[N], inDecomposed[N][8]
static Blk T_Bytes2Nibbles(Builder &B, int N, const Code *in) {
    size_t n = (size_t)N; Blk o = B.alloc(11 * n); B.copy(o.pos + 2 * n, in, n);
    for (size_t i = 0; i < n; i++) {
        Blk nb = T_Num2Bits(B, 8, in[i]); B.copy(o.pos + 3 * n + 8 * i, &B.at(nb.pos), 8);
        Code lo = ZERO, hi = ZERO;
        for (unsigned j = 0; j < 4; j++) {
            lo = B.fma(B.at(nb.pos + j), B.pow2(j), lo);
            hi = B.fma(B.at(nb.pos + j + 4), B.pow2(j), hi);
        }
        B.at(o.pos + 2 * i) = hi; B.at(o.pos + 2 * i + 1) = lo;
    }
    return o;
}
// Nibbles2Bytes(n) :132-141  own: bytes[n], nibbles[2n]
static Blk T_Nibbles2Bytes(Builder &B, int n_, const Code *nib) {
    size_t n = (size_t)n_; Blk o = B.alloc(3 * n); B.copy(o.pos + n, nib, 2 * n);
    for (size_t i = 0; i < n; i++) {
        T_AssertBits(B, 4, nib[2 * i]); T_AssertBits(B, 4, nib[2 * i + 1]);
        B.at(o.pos + i) = B.fma(nib[2 * i], c_const(16), nib[2 * i + 1]);
    }
    return o;
}

// ============================================================================================================
// circuits/utils/divide.circom
// ============================================================================================================
// Divide(N) :17-33  own: out, rem, a, b
static Blk T_Divide(Builder &B, int N, Code a, Code b) {
    Blk o = B.alloc(4);
    Code q, r; Fr fa, fb;
    if (B.const_val(a, fa) && B.const_val(b, fb) && !fr_is_zero(fb)) { Fr fq, fr_; fr_divmod(fa, fb, fq, fr_); q = B.konst(fq); r = B.konst(fr_); }
    else { q = B.divmod(false, a, b, o.sig); r = B.divmod(true, a, b, o.sig); }
    B.at(o.pos) = q; B.at(o.pos + 1) = r; B.at(o.pos + 2) = a; B.at(o.pos + 3) = b;
    T_AssertLessThan(B, N, r, b);
    T_AssertLessEqThan(B, N, q, a);
    B.chk_eq(B.fma(q, b, r), a, o.sig);
    return o;
}

// ============================================================================================================
// circuits/utils/selector.circom
// ============================================================================================================
// Selector(n) :21-46  own: out, vals[n], select, isEq[n], sum[n+1]
static Blk T_Selector(Builder &B, int n_, const Code *vals, Code select) {
    size_t n = (size_t)n_; Blk o = B.alloc(1 + n + 1 + n + n + 1);
    B.copy(o.pos + 1, vals, n); B.at(o.pos + 1 + n) = select;
    size_t isEq = o.pos + 2 + n, sum = isEq + n;
    B.at(sum) = ZERO;
    for (size_t i = 0; i < n; i++) { Blk e = T_IsEqual(B, select, c_const((uint32_t)i)); B.at(isEq + i) = B.at(e.pos); }
    B.selsum(select, vals, n, &B.at(sum + 1));          // sum[i+1] = sum_{j<=i} isEq[j]*vals[j]
    B.chk_eq(B.gtc(select, (uint32_t)(n - 1)), ZERO, o.sig);   // sumIsEq === 1  <=>  select in [0, n)
    B.at(o.pos) = B.at(sum + n); return o;
}
// SelectorArray1D(n,p) :62-77 / SelectorArray2D(n,p,q) :91-110  own: out[cols], arrays[n][cols], select, arraysT[cols][n]
static Blk T_SelectorArray(Builder &B, int n_, size_t cols, const Code *arrays, Code select) {
    size_t n = (size_t)n_; Blk o = B.alloc(cols + n * cols + 1 + cols * n);
    B.copy(o.pos + cols, arrays, n * cols); B.at(o.pos + cols + n * cols) = select;
    size_t T = o.pos + cols + n * cols + 1;
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < cols; j++) B.at(T + j * n + i) = arrays[i * cols + j];
    for (size_t j = 0; j < cols; j++) { Blk s = T_Selector(B, n_, &B.at(T + j * n), select); B.at(o.pos + j) = B.at(s.pos); }
    return o;
}

// ============================================================================================================
// circuits/utils/shift.circom, concat.circom
// ============================================================================================================
// ShiftLeft(n) :17-36  own: out[n], in[n], count, isEq[n][n], temp[n][n]
static Blk T_ShiftLeft(Builder &B, int n_, const Code *in, Code count) {
    size_t n = (size_t)n_; Blk o = B.alloc(2 * n + 1 + 2 * n * n);
    B.copy(o.pos + n, in, n); B.at(o.pos + 2 * n) = count;
    size_t isEq = o.pos + 2 * n + 1, temp = isEq + n * n;
    T_AssertLessEqThan(B, 16, count, c_const((uint32_t)n));
    for (size_t i = 0; i < n; i++) {
        std::vector<Code> terms;
        for (size_t j = 0; j < n; j++) {
            Blk e = T_IsEqual(B, c_const((uint32_t)i), B.sub(c_const((uint32_t)j), count));
            Code eq = B.at(e.pos); B.at(isEq + i * n + j) = eq;
            Code tv = B.mul(eq, in[j]); B.at(temp + i * n + j) = tv; terms.push_back(tv);
        }
        B.at(o.pos + i) = B.sum_tree(terms);
    }
    return o;
}
// ShiftRight(n,maxShift) :51-75  own: out[n+ms], in[n], count, isEq[ms+1], temps[ms+1][n]
static Blk T_ShiftRight(Builder &B, int n_, int ms_, const Code *in, Code count) {
    size_t n = (size_t)n_, ms = (size_t)ms_; Blk o = B.alloc(n + ms + n + 1 + ms + 1 + (ms + 1) * n);
    B.copy(o.pos + n + ms, in, n); B.at(o.pos + 2 * n + ms) = count;
    size_t isEq = o.pos + 2 * n + ms + 1, temps = isEq + ms + 1;
    T_AssertLessEqThan(B, 16, count, c_const((uint32_t)ms));
    std::vector<std::vector<Code>> acc(n + ms);
    for (size_t i = 0; i <= ms; i++) {
        Blk e = T_IsEqual(B, c_const((uint32_t)i), count); Code eq = B.at(e.pos); B.at(isEq + i) = eq;
        for (size_t j = 0; j < n; j++) { Code tv = B.mul(eq, in[j]); B.at(temps + i * n + j) = tv; acc[i + j].push_back(tv); }
    }
    for (size_t k = 0; k < n + ms; k++) B.at(o.pos + k) = B.sum_tree(acc[k]);
    return o;
}
// Mask(n) :18-30  own: out[n], in[n], count, filter[n]
static Blk T_Mask(Builder &B, int n_, const Code *in, Code count) {
    size_t n = (size_t)n_; Blk o = B.alloc(3 * n + 1); B.copy(o.pos + n, in, n); B.at(o.pos + 2 * n) = count;
    Blk f = T_Filter(B, n_, count); B.copy(o.pos + 2 * n + 1, &B.at(f.pos), n);
    for (size_t i = 0; i < n; i++) B.at(o.pos + i) = B.mul(in[i], B.at(f.pos + i));
    return o;
}
// Concat(A,B) :47-84  own: out[A+B], outLen, a[A], aLen, b[B], bLen, maskedA[A], maskedB[B], shiftedB[A+B]
static Blk T_Concat(Builder &B, int A, int Bn, const Code *a, Code aLen, const Code *b, Code bLen) {
    size_t NA = (size_t)A, NB = (size_t)Bn, T = NA + NB;
    Blk o = B.alloc(T + 1 + NA + 1 + NB + 1 + NA + NB + T);
    size_t ia = o.pos + T + 1, iaL = ia + NA, ib = iaL + 1, ibL = ib + NB, mA = ibL + 1, mB = mA + NA, sB = mB + NB;
    B.copy(ia, a, NA); B.at(iaL) = aLen; B.copy(ib, b, NB); B.at(ibL) = bLen;
    T_AssertLessEqThan(B, 16, aLen, c_const((uint32_t)A));
    T_AssertLessEqThan(B, 16, bLen, c_const((uint32_t)Bn));
    Blk ma = T_Mask(B, A, a, aLen); B.copy(mA, &B.at(ma.pos), NA);
    Blk mb = T_Mask(B, Bn, b, bLen); B.copy(mB, &B.at(mb.pos), NB);
    Blk sh = T_ShiftRight(B, Bn, A, &B.at(mb.pos), aLen); B.copy(sB, &B.at(sh.pos), T);
    for (size_t i = 0; i < T; i++) B.at(o.pos + i) = i < NA ? B.add(B.at(ma.pos + i), B.at(sh.pos + i)) : B.at(sh.pos + i);
    B.at(o.pos + T) = B.add(aLen, bLen);
    return o;
}

// ============================================================================================================
// circuits/utils/substring_check.circom
// ============================================================================================================
// SubstringCheck(maxMainLen, subLen) :24-100
static Blk T_SubstringCheck(Builder &B, int maxMainLen, int subLen, const Code *mainInput, Code mainLen, const Code *subInput) {
    size_t MM = (size_t)maxMainLen, SL = (size_t)subLen, Kn = MM - SL + 1;
    Blk o = B.alloc(1 + MM + 1 + SL + 1 + (MM + 1) + Kn + Kn + (Kn + 1) + (Kn + 1) + 1);
    size_t iMain = o.pos + 1, iLen = iMain + MM, iSub = iLen + 1, subNum = iSub + SL, Mo = subNum + 1,
           exists = Mo + MM + 1, isLast = exists + Kn, allowed = isLast + Kn, sums = allowed + Kn + 1, dne = sums + Kn + 1;
    B.copy(iMain, mainInput, MM); B.at(iLen) = mainLen; B.copy(iSub, subInput, SL);
    T_AssertByteString(B, subLen, subInput);
    T_AssertByteString(B, maxMainLen, mainInput);
    T_AssertLessEqThan(B, 16, mainLen, c_const((uint32_t)MM));
    T_AssertLessEqThan(B, 16, c_const((uint32_t)SL), mainLen);
    Blk sn = T_LittleEndianBytes2Num(B, subLen, subInput); Code subN = B.at(sn.pos); B.at(subNum) = subN;
    B.at(Mo) = ZERO;
    Fr pw = fr_from_u64(1), c256 = fr_from_u64(256);
    {   // M[i+1] = mainInput[i]*256^i + M[i]
        std::vector<Code> terms(MM);
        for (size_t i = 0; i < MM; i++) { terms[i] = B.mul(mainInput[i], B.konst(pw)); pw = fr_mul(pw, c256); }
        B.prefix_sum(ZERO, terms.data(), MM, &B.at(Mo + 1));
    }
    B.at(allowed) = ONE; B.at(sums) = ZERO;
    pw = fr_from_u64(1);
    Code lastIdx = B.add(B.sub(mainLen, c_const((uint32_t)SL)), ONE);
    std::vector<Code> sterms(Kn);
    for (size_t i = 0; i < Kn; i++) {
        Blk e1 = T_IsEqual(B, c_const((uint32_t)i), lastIdx); B.at(isLast + i) = B.at(e1.pos);
        B.at(allowed + i + 1) = B.gtc(lastIdx, (uint32_t)i);       // allowed[i]*(1 - isLastIndex[i]) == (lastIdx > i)
        Blk e2 = T_IsEqual(B, B.mul(subN, B.konst(pw)), B.sub(B.at(Mo + i + SL), B.at(Mo + i)), /*likely_large=*/true);
        Code ex = B.at(e2.pos); B.at(exists + i) = ex;
        sterms[i] = B.mul(B.at(allowed + i + 1), ex);
        pw = fr_mul(pw, c256);
    }
    B.prefix_sum(ZERO, sterms.data(), Kn, &B.at(sums + 1));   // sums[i+1] = sums[i] + allowed[i+1]*exists[i]
    Blk z = T_IsZero(B, B.at(sums + Kn)); B.at(dne) = B.at(z.pos);
    B.at(o.pos) = B.not1(B.at(z.pos));
    return o;
}

// ============================================================================================================
// circuits/utils/keccak.circom -- lane level.  A lane is a store word; its 64 signals are BIT codes.
// The emitters below write codes sequentially through a cursor because inside the Keccak sub-circuit every
// lane word index is known from the layout alone.
// ============================================================================================================
struct LaneSink {
    Code *p;
    uint32_t base;                       // added to every lane word (0 when emitting the shared relative table)
    void lane(Lane l) { for (uint32_t k = 0; k < 64; k++) *p++ = (l.w == NONE_IDX) ? ZERO : c_bit(base + l.w, k); }
    // XorArray/OrArray/AndArray(64) :77-128 -- own out,a,b then 64 gates [out,a,b]   (384 signals)
    void gate_array(Lane out, Lane a, Lane b) {
        lane(out); lane(a); lane(b);
        for (uint32_t k = 0; k < 64; k++) {
            *p++ = (out.w == NONE_IDX) ? ZERO : c_bit(base + out.w, k);
            *p++ = (a.w == NONE_IDX) ? ZERO : c_bit(base + a.w, k);
            *p++ = (b.w == NONE_IDX) ? ZERO : c_bit(base + b.w, k);
        }
    }
    void unary(Lane out, Lane in) { lane(out); lane(in); }   // ShL/ShR :19-51, NotArray :92-98  (128 signals)
};
// KeccakfRound(r) :290-297 relative to the round base (see program.h for the word map)
static void emit_round(LaneSink &S) {
    auto W = [](uint32_t w) { return Lane{w}; };
    Lane in[25], th[25], rp[25], ch[25], out[25];
    for (int l = 0; l < 25; l++) { in[l] = W((uint32_t)l); th[l] = W(rw_th(l)); ch[l] = W(rw_ch(l, 2)); out[l] = W(rw_out(l)); }
    rp[0] = th[0];
    for (int i = 0; i < 24; i++) rp[keccak_rot(i + 1)] = W(rw_rp(i, 2));
    // own: out, in, theta, rhopi, chi
    for (int l = 0; l < 25; l++) S.lane(out[l]);
    for (int l = 0; l < 25; l++) S.lane(in[l]);
    for (int l = 0; l < 25; l++) S.lane(th[l]);
    for (int l = 0; l < 25; l++) S.lane(rp[l]);
    for (int l = 0; l < 25; l++) S.lane(ch[l]);
    // Theta :151-170  own: out, in, c[5], d[5]
    for (int l = 0; l < 25; l++) S.lane(th[l]);
    for (int l = 0; l < 25; l++) S.lane(in[l]);
    for (int i = 0; i < 5; i++) S.lane(W(rw_x5(i, 3)));
    for (int i = 0; i < 5; i++) S.lane(W(rw_dd(i, 3)));
    for (int i = 0; i < 5; i++) {        // Xor5(64) :58-70  own: out,a,b,c,d,e,xor_ab,xor_abc,xor_abcd
        S.lane(W(rw_x5(i, 3)));
        for (int j = 0; j < 5; j++) S.lane(in[5 * j + i]);
        S.lane(W(rw_x5(i, 0))); S.lane(W(rw_x5(i, 1))); S.lane(W(rw_x5(i, 2)));
        S.gate_array(W(rw_x5(i, 0)), in[i], in[5 + i]);
        S.gate_array(W(rw_x5(i, 1)), W(rw_x5(i, 0)), in[10 + i]);
        S.gate_array(W(rw_x5(i, 2)), W(rw_x5(i, 1)), in[15 + i]);
        S.gate_array(W(rw_x5(i, 3)), W(rw_x5(i, 2)), in[20 + i]);
    }
    for (int i = 0; i < 5; i++) {        // D :135-144  own: out,a,b,aux0,aux1,aux2
        Lane a = W(rw_x5((i + 1) % 5, 3)), b = W(rw_x5((i + 4) % 5, 3));
        S.lane(W(rw_dd(i, 3))); S.lane(a); S.lane(b); S.lane(W(rw_dd(i, 0))); S.lane(W(rw_dd(i, 1))); S.lane(W(rw_dd(i, 2)));
        S.unary(W(rw_dd(i, 0)), a);                                   // ShL(64,1)
        S.unary(W(rw_dd(i, 1)), a);                                   // ShR(64,63)
        S.gate_array(W(rw_dd(i, 2)), W(rw_dd(i, 0)), W(rw_dd(i, 1)));   // OrArray
        S.gate_array(W(rw_dd(i, 3)), b, W(rw_dd(i, 2)));                // XorArray(b, aux2)
    }
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) S.gate_array(th[i + 5 * j], in[i + 5 * j], W(rw_dd(i, 3)));
    // RhoPi :191-204  own: out, in
    for (int l = 0; l < 25; l++) S.lane(rp[l]);
    for (int l = 0; l < 25; l++) S.lane(th[l]);
    for (int i = 0; i < 24; i++) {       // stepRhoPi :177-184  own: out,a,aux0,aux1
        Lane a = th[keccak_rot(i)];
        S.lane(W(rw_rp(i, 2))); S.lane(a); S.lane(W(rw_rp(i, 0))); S.lane(W(rw_rp(i, 1)));
        S.unary(W(rw_rp(i, 0)), a);                                   // ShR(64,shr)
        S.unary(W(rw_rp(i, 1)), a);                                   // ShL(64,shl)
        S.gate_array(W(rw_rp(i, 2)), W(rw_rp(i, 0)), W(rw_rp(i, 1)));
    }
    // Chi :228-241  own: out, in
    for (int l = 0; l < 25; l++) S.lane(ch[l]);
    for (int l = 0; l < 25; l++) S.lane(rp[l]);
    for (int l = 0; l < 25; l++) {       // stepChi :212-221  own: out,a,b,c,bXor,bc
        Lane a = rp[l], b = rp[chi_b(l)], c = rp[chi_c(l)];
        S.lane(ch[l]); S.lane(a); S.lane(b); S.lane(c); S.lane(W(rw_ch(l, 0))); S.lane(W(rw_ch(l, 1)));
        S.unary(W(rw_ch(l, 0)), b);                                   // NotArray
        S.gate_array(W(rw_ch(l, 1)), W(rw_ch(l, 0)), c);              // AndArray
        S.gate_array(ch[l], a, W(rw_ch(l, 1)));                       // XorArray
    }
    // Iota(r) :273-283  own: out, in, roundConstants
    for (int l = 0; l < 25; l++) S.lane(out[l]);
    for (int l = 0; l < 25; l++) S.lane(ch[l]);
    S.lane(W(RW_RC));
    S.lane(W(RW_RC));                                                 // RoundConstants(r) :248-266  own: out[64]
    S.gate_array(out[0], ch[0], W(RW_RC));
}

// Absorb :304-323  (s: previous state words or NONE_IDX; blk: 17 block words).  Returns the state-out word base.
static uint32_t T_Absorb(Builder &B, uint32_t s_idx, uint32_t blk_idx, Blk *own) {
    uint32_t A = B.absorb(s_idx, blk_idx), out_w = A + RW * 24;
    Blk o = B.alloc(1600 + 1600 + 1088 + 1600); if (own) *own = o;
    LaneSink S{&B.at(o.pos), 0};
    for (uint32_t l = 0; l < 25; l++) S.lane(Lane{out_w + l});
    for (uint32_t l = 0; l < 25; l++) S.lane(Lane{s_idx == NONE_IDX ? NONE_IDX : s_idx + l});
    for (uint32_t l = 0; l < 17; l++) S.lane(Lane{blk_idx + l});
    for (uint32_t l = 0; l < 25; l++) S.lane(Lane{A + l});
    for (uint32_t l = 0; l < 17; l++) {                               // XorArray(64)(s[i], block[i])
        Blk x = B.alloc(384); LaneSink X{&B.at(x.pos), 0};
        X.gate_array(Lane{A + l}, Lane{s_idx == NONE_IDX ? NONE_IDX : s_idx + l}, Lane{blk_idx + l});
    }
    // Keccakf :356-367  own: out, in, midRound[25][25][64]
    Blk k = B.alloc(1600 + 1600 + 25 * 1600); LaneSink Kf{&B.at(k.pos), 0};
    for (uint32_t l = 0; l < 25; l++) Kf.lane(Lane{out_w + l});
    for (uint32_t l = 0; l < 25; l++) Kf.lane(Lane{A + l});
    for (uint32_t r = 0; r <= 24; r++) for (uint32_t l = 0; l < 25; l++) Kf.lane(Lane{A + RW * r + l});
    for (uint32_t r = 0; r < 24; r++) B.round_block(A + RW * r);
    return out_w;
}
// Final(n) :330-349 and Keccak(n) :374-385 ; in_words: n*17 block words ; returns Keccak's own block
static Blk T_Keccak(Builder &B, int n_, uint32_t in_words, Code blocks) {
    size_t n = (size_t)n_;
    Blk ko = B.alloc(256 + n * 1088 + 1 + 1600);                      // Keccak own: out[256], in, blocks, finalState
    { LaneSink S{&B.at(ko.pos + 256), 0}; for (uint32_t w = 0; w < n * 17; w++) S.lane(Lane{in_words + w}); }
    B.at(ko.pos + 256 + n * 1088) = blocks;
    Blk fo = B.alloc(1600 + n * 1088 + 1 + (n + 1) * 1600);           // Final own: out, in, blocks, s[n+1]
    { LaneSink S{&B.at(fo.pos + 1600), 0}; for (uint32_t w = 0; w < n * 17; w++) S.lane(Lane{in_words + w}); }
    B.at(fo.pos + 1600 + n * 1088) = blocks;
    size_t s = fo.pos + 1600 + n * 1088 + 1;
    for (size_t i = 0; i < 1600; i++) B.at(s + i) = ZERO;             // s[0] = 0
    uint32_t st = NONE_IDX;
    for (size_t b = 0; b < n; b++) {
        st = T_Absorb(B, st, in_words + 17 * (uint32_t)b, nullptr);
        LaneSink S{&B.at(s + 1600 * (b + 1)), 0}; for (uint32_t l = 0; l < 25; l++) S.lane(Lane{st + l});
    }
    Blk sel = T_SelectorArray(B, n_ + 1, 1600, &B.at(s), blocks);
    B.copy(fo.pos, &B.at(sel.pos), 1600);
    B.copy(ko.pos + 256 + n * 1088 + 1, &B.at(fo.pos), 1600);
    B.copy(ko.pos, &B.at(fo.pos), 256);
    return ko;
}
// Pad(maxBlocks, blockSize) :412-446  own: out[B], numBlocks, in[B], inLen, div, rem, filter[B+1], isEq[B], isLast[B]
static Blk T_Pad(Builder &B, int maxBlocks, int blockSize, const Code *in, Code inLen) {
    size_t Bn = (size_t)maxBlocks * (size_t)blockSize; Blk o = B.alloc(Bn + 1 + Bn + 1 + 2 + (Bn + 1) + Bn + Bn);
    size_t numBlocks = o.pos + Bn, iIn = numBlocks + 1, iLen = iIn + Bn, div = iLen + 1, rem = div + 1,
           filter = rem + 1, isEq = filter + Bn + 1, isLast = isEq + Bn;
    B.copy(iIn, in, Bn); B.at(iLen) = inLen;
    Blk d = T_Divide(B, 16, inLen, c_const((uint32_t)blockSize)); B.at(div) = B.at(d.pos); B.at(rem) = B.at(d.pos + 1);
    Code nbk = B.add(B.at(div), ONE); B.at(numBlocks) = nbk;
    T_AssertLessEqThan(B, 16, nbk, c_const((uint32_t)maxBlocks));
    B.at(filter) = ONE;
    for (size_t i = 0; i < Bn; i++) {
        Blk e = T_IsEqual(B, c_const((uint32_t)i), inLen); Code eq = B.at(e.pos); B.at(isEq + i) = eq;
        B.at(filter + i + 1) = B.gtc(inLen, (uint32_t)i);          // filter[i]*(1 - isEq[i]) == (inLen > i)
    }
    Code lastPos = B.sub(B.mul(nbk, c_const((uint32_t)blockSize)), ONE);
    for (size_t i = 0; i < Bn; i++) {
        Blk e = T_IsEqual(B, c_const((uint32_t)i), lastPos); Code l = B.at(e.pos); B.at(isLast + i) = l;
        B.at(o.pos + i) = B.fma(l, c_const(0x80), B.fma(in[i], B.at(filter + i + 1), B.at(isEq + i)));
    }
    return o;
}
// KeccakBytes(maxBlocks) :454-489
static Blk T_KeccakBytes(Builder &B, int maxBlocks, const Code *in, Code inLen) {
    size_t Bn = (size_t)maxBlocks * 136;
    Blk o = B.alloc(32 + Bn + 1 + Bn + 1 + 24 * Bn + 256 + 256);
    size_t iIn = o.pos + 32, iLen = iIn + Bn, padded = iLen + 1, numBlocks = padded + Bn, inBitsArray = numBlocks + 1,
           inBits = inBitsArray + 8 * Bn, inBlocks = inBits + 8 * Bn, outBits = inBlocks + 8 * Bn, outBytes = outBits + 256;
    B.copy(iIn, in, Bn); B.at(iLen) = inLen;
    T_AssertLessThan(B, 16, inLen, c_const((uint32_t)Bn));
    Blk p = T_Pad(B, maxBlocks, 136, in, inLen); B.copy(padded, &B.at(p.pos), Bn); B.at(numBlocks) = B.at(p.pos + Bn);
    uint32_t words = B.pack8_words(&B.at(padded), (uint32_t)(Bn / 8));         // bytes -> 17 lanes per block
    for (size_t i = 0; i < Bn; i++) {                                          // Num2Bits(8)(padded[i])
        Blk nb = B.alloc(9);
        for (uint32_t k = 0; k < 8; k++) B.at(nb.pos + k) = c_bit(words + (uint32_t)(i / 8), 8 * (uint32_t)(i % 8) + k);
        B.at(nb.pos + 8) = B.at(padded + i);
        B.chk_range(B.at(padded + i), 8, nb.sig);
        B.copy(inBitsArray + 8 * i, &B.at(nb.pos), 8);
    }
    Blk fl = T_CopyArray(B, 8 * Bn, &B.at(inBitsArray)); B.copy(inBits, &B.at(fl.pos), 8 * Bn);   // Flatten
    B.copy(inBlocks, &B.at(inBits), 8 * Bn);
    Blk k = T_Keccak(B, maxBlocks, words, B.at(numBlocks)); B.copy(outBits, &B.at(k.pos), 256);
    Blk rs = T_CopyArray(B, 256, &B.at(outBits)); B.copy(outBytes, &B.at(rs.pos), 256);           // Reshape
    for (size_t i = 0; i < 32; i++) { Blk bn = T_Bits2Num(B, 8, &B.at(outBytes + 8 * i)); B.at(o.pos + i) = B.at(bn.pos); }
    return o;
}

// ============================================================================================================
// circuits/utils/public_commitment.circom, constants.circom, burn_address.circom, proof_of_work.circom
// ============================================================================================================
// PublicCommitment(N) :18-42  own: out, in[N][32], flattenIn, block, hash[32], reducedHash[31]
static Blk T_PublicCommitment(Builder &B, int N, const Code *in) {
    size_t n32 = (size_t)N * 32; int nb = (N * 32) / 136 + ((N * 32) % 136 != 0); size_t blk = (size_t)nb * 136;
    Blk o = B.alloc(1 + n32 + n32 + blk + 32 + 31);
    size_t iIn = o.pos + 1, flat = iIn + n32, block = flat + n32, hash = block + blk, red = hash + 32;
    B.copy(iIn, in, n32);
    for (int i = 0; i < N; i++) T_AssertByteString(B, 32, in + 32 * i);
    Blk f = T_CopyArray(B, n32, in); B.copy(flat, &B.at(f.pos), n32);
    Blk ft = T_Fit(B, (int)n32, (int)blk, &B.at(flat)); B.copy(block, &B.at(ft.pos), blk);
    Blk k = T_KeccakBytes(B, nb, &B.at(block), c_const((uint32_t)n32)); B.copy(hash, &B.at(k.pos), 32);
    Blk f2 = T_Fit(B, 32, 31, &B.at(hash)); B.copy(red, &B.at(f2.pos), 31);
    Blk be = T_BigEndianBytes2Num(B, 31, &B.at(red)); B.at(o.pos) = B.at(be.pos);
    return o;
}
// constants.circom :3-15
static Code POSEIDON_PREFIX(Builder &B, int add) {
    // keccak("EIP-7503") mod p = 5265656504298861414514317065875120428884240036965045859626767452974705356670
    Fr r; const uint32_t l[8] = {0x3d892f7eu, 0xf0363f98u, 0x980a6b46u, 0xd115b780u, 0xcd46cec2u, 0x007d2482u, 0xee7876b8u, 0x0ba44186u};
    memcpy(r.l, l, 32);
    return B.konst(fr_add(r, fr_from_u64((uint64_t)add)));
}
// BurnAddress :47-58  own: addressBytes[20], burnKey, revealAmount, burnExtraCommitment, hash, hashBytes[32]
static Blk T_BurnAddress(Builder &B, Code burnKey, Code revealAmount, Code bec) {
    Blk o = B.alloc(20 + 3 + 1 + 32);
    B.at(o.pos + 20) = burnKey; B.at(o.pos + 21) = revealAmount; B.at(o.pos + 22) = bec;
    Code ins[4] = {POSEIDON_PREFIX(B, 0), burnKey, revealAmount, bec};
    Blk p = T_Poseidon(B, 4, ins); B.at(o.pos + 23) = B.at(p.pos);
    Blk b = T_Num2BigEndianBytes(B, 32, B.at(p.pos)); B.copy(o.pos + 24, &B.at(b.pos), 32);
    Blk f = T_Fit(B, 32, 20, &B.at(o.pos + 24)); B.copy(o.pos, &B.at(f.pos), 20);
    return o;
}
// BurnAddressHash :67-83  own: addressHashNibbles[64], 3 inputs, addressBytes[20], addressBytesBlock[136], addressHash[32]
static Blk T_BurnAddressHash(Builder &B, Code burnKey, Code revealAmount, Code bec) {
    Blk o = B.alloc(64 + 3 + 20 + 136 + 32);
    B.at(o.pos + 64) = burnKey; B.at(o.pos + 65) = revealAmount; B.at(o.pos + 66) = bec;
    Blk a = T_BurnAddress(B, burnKey, revealAmount, bec); B.copy(o.pos + 67, &B.at(a.pos), 20);
    Blk f = T_Fit(B, 20, 136, &B.at(o.pos + 67)); B.copy(o.pos + 87, &B.at(f.pos), 136);
    Blk k = T_KeccakBytes(B, 1, &B.at(o.pos + 87), c_const(20)); B.copy(o.pos + 223, &B.at(k.pos), 32);
    Blk nb = T_Bytes2Nibbles(B, 32, &B.at(o.pos + 223)); B.copy(o.pos, &B.at(nb.pos), 64);
    return o;
}
// EIP7503 :11-21
static Blk T_EIP7503(Builder &B) {
    static const uint8_t s[8] = {69, 73, 80, 45, 55, 53, 48, 51};
    Blk o = B.alloc(8); for (int i = 0; i < 8; i++) B.at(o.pos + (size_t)i) = c_const(s[i]); return o;
}
// ConcatFixed4(A,B,C,D) :28-48  own: out[A+B+C+D], a, b, c, d
static Blk T_ConcatFixed4(Builder &B, int A, int Bn, int C, int D, const Code *a, const Code *b, const Code *c, const Code *d) {
    size_t T = (size_t)(A + Bn + C + D); Blk o = B.alloc(2 * T);
    B.copy(o.pos, a, (size_t)A); B.copy(o.pos + (size_t)A, b, (size_t)Bn); B.copy(o.pos + (size_t)(A + Bn), c, (size_t)C);
    B.copy(o.pos + (size_t)(A + Bn + C), d, (size_t)D);
    B.copy(o.pos + T, &B.at(o.pos), T);
    return o;
}
// ProofOfWorkChecker :54-81
static Blk T_ProofOfWorkChecker(Builder &B, Code burnKey, Code revealAmount, Code bec, Code minimumZeroBytes) {
    Blk o = B.alloc(4 + 96 + 8 + 104 + 136 + 32 + 32);
    size_t bk = o.pos + 4, ra = bk + 32, be = ra + 32, eip = be + 32, hin = eip + 8, blk = hin + 104, kec = blk + 136, sbz = kec + 32;
    B.at(o.pos) = burnKey; B.at(o.pos + 1) = revealAmount; B.at(o.pos + 2) = bec; B.at(o.pos + 3) = minimumZeroBytes;
    Blk x = T_Num2BigEndianBytes(B, 32, burnKey); B.copy(bk, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, revealAmount); B.copy(ra, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, bec); B.copy(be, &B.at(x.pos), 32);
    x = T_EIP7503(B); B.copy(eip, &B.at(x.pos), 8);
    x = T_ConcatFixed4(B, 32, 32, 32, 8, &B.at(bk), &B.at(ra), &B.at(be), &B.at(eip)); B.copy(hin, &B.at(x.pos), 104);
    x = T_Fit(B, 104, 136, &B.at(hin)); B.copy(blk, &B.at(x.pos), 136);
    x = T_KeccakBytes(B, 1, &B.at(blk), c_const(104)); B.copy(kec, &B.at(x.pos), 32);
    x = T_Filter(B, 32, minimumZeroBytes); B.copy(sbz, &B.at(x.pos), 32);
    for (size_t i = 0; i < 32; i++) B.chk_eq(B.mul(B.at(kec + i), B.at(sbz + i)), ZERO, o.sig);
    return o;
}

// ============================================================================================================
// circuits/utils/rlp/*.circom
// ============================================================================================================
// CountBytes(N) integer.circom:16-49  own: len, bytes[N], isZero[N], stillZero[N]
static Blk T_CountBytes(Builder &B, int N, const Code *bytes) {
    size_t n = (size_t)N; Blk o = B.alloc(1 + 3 * n); B.copy(o.pos + 1, bytes, n);
    for (size_t i = 0; i < n; i++) { Blk z = T_IsZero(B, bytes[i]); B.at(o.pos + 1 + n + i) = B.at(z.pos); }
    std::vector<Code> terms;
    for (size_t i = 0; i < n; i++) {
        Code sz = i == 0 ? B.at(o.pos + 1 + n) : B.mul(B.at(o.pos + 1 + n + i), B.at(o.pos + 1 + 2 * n + i - 1));
        B.at(o.pos + 1 + 2 * n + i) = sz; terms.push_back(sz);
    }
    B.at(o.pos) = B.sub(c_const((uint32_t)n), B.sum_tree(terms)); return o;
}
// RlpInteger(N) integer.circom:67-110
static Blk T_RlpInteger(Builder &B, int N, Code in) {
    size_t n = (size_t)N; Blk o = B.alloc(n + 1 + 1 + 1 + n + 1 + n + 3);
    size_t outLen = o.pos + n + 1, iIn = outLen + 1, bytes = iIn + 1, length = bytes + n, bigEndian = length + 1,
           isSingle = bigEndian + n, isZero = isSingle + 1, first = isZero + 1;
    B.at(iIn) = in;
    Blk x = T_Num2BigEndianBytes(B, N, in); B.copy(bytes, &B.at(x.pos), n);
    x = T_CountBytes(B, N, &B.at(bytes)); Code len = B.at(x.pos); B.at(length) = len;
    x = T_ShiftLeft(B, N, &B.at(bytes), B.sub(c_const((uint32_t)n), len)); B.copy(bigEndian, &B.at(x.pos), n);
    x = T_LessThan(B, N * 8, in, c_const(128)); Code single = B.at(x.pos); B.at(isSingle) = single;
    x = T_IsZero(B, in); Code iz = B.at(x.pos); B.at(isZero) = iz;
    x = T_Mux1(B, B.add(c_const(0x80), len), in, single); B.at(first) = B.at(x.pos);
    B.at(o.pos) = B.fma(iz, c_const(0x80), B.at(first));
    Code ns = B.not1(single);
    for (size_t i = 1; i < n + 1; i++) B.at(o.pos + i) = B.mul(ns, B.at(bigEndian + i - 1));
    B.at(outLen) = B.add(B.add(ns, len), iz);
    return o;
}
// RlpEmptyAccount(maxBalanceBytes) empty_account.circom:20-134
static const uint8_t STORAGE_CODE_RLP[66] = {
    160, 86, 232, 31, 23, 27, 204, 85, 166, 255, 131, 69, 230, 146, 192, 248, 110, 91, 72, 224, 27, 153, 108, 173, 192, 1, 98, 47, 181, 227, 99, 180, 33,
    160, 197, 210, 70, 1, 134, 247, 35, 60, 146, 126, 125, 178, 220, 199, 3, 192, 229, 0, 182, 83, 202, 130, 39, 59, 123, 250, 216, 4, 93, 133, 164, 112};
static Blk T_RlpEmptyAccount(Builder &B, int mbb, Code balance) {
    size_t m = (size_t)mbb, OL = 4 + m + 66; Blk o = B.alloc(OL + 1 + 1 + (4 + m) + 1 + (m + 1) + 1 + 1 + 66);
    size_t outLen = o.pos + OL, iBal = outLen + 1, pre = iBal + 1, preLen = pre + 4 + m, balRlp = preLen + 1,
           balRlpLen = balRlp + m + 1, nabLen = balRlpLen + 1, sc = nabLen + 1;
    B.at(iBal) = balance;
    B.at(pre + 2) = c_const(0x80);
    Blk r = T_RlpInteger(B, mbb, balance); B.copy(balRlp, &B.at(r.pos), m + 1); B.at(balRlpLen) = B.at(r.pos + m + 1);
    for (size_t i = 0; i < m + 1; i++) B.at(pre + 3 + i) = B.at(balRlp + i);
    B.at(nabLen) = B.add(ONE, B.at(balRlpLen));
    B.at(preLen) = B.add(c_const(2), B.at(nabLen));
    for (size_t i = 0; i < 66; i++) B.at(sc + i) = c_const(STORAGE_CODE_RLP[i]);
    B.at(pre) = c_const(0xf8);
    B.at(pre + 1) = B.add(B.at(nabLen), c_const(66));
    Blk cc = T_Concat(B, 4 + mbb, 66, &B.at(pre), B.at(preLen), &B.at(sc), c_const(66));
    B.copy(o.pos, &B.at(cc.pos), OL); B.at(outLen) = B.at(cc.pos + OL);
    return o;
}
// TruncatedAddressHash(addressHashBytes) merkle_patricia_trie_leaf.circom:50-90 (`temp` :76 never assigned => 0)
static Blk T_TruncatedAddressHash(Builder &B, int ahb, const Code *nibbles, Code nibLen) {
    size_t a = (size_t)ahb; Blk o = B.alloc((a + 1) + 1 + 2 * a + 1 + 2 + 2 * a + (2 * a + 2) + (2 * a - 1));
    size_t outLen = o.pos + a + 1, iNib = outLen + 1, iLen = iNib + 2 * a, div = iLen + 1, rem = div + 1, shifted = rem + 1,
           outNib = shifted + 2 * a, temp = outNib + 2 * a + 2;
    B.copy(iNib, nibbles, 2 * a); B.at(iLen) = nibLen;
    for (size_t i = 0; i < 2 * a - 1; i++) B.at(temp + i) = ZERO;
    T_AssertLessEqThan(B, 7, nibLen, c_const((uint32_t)(2 * a)));
    Blk d = T_Divide(B, 7, nibLen, c_const(2)); B.at(div) = B.at(d.pos); Code rm = B.at(d.pos + 1); B.at(rem) = rm;
    Blk s = T_ShiftLeft(B, 2 * ahb, nibbles, B.sub(c_const((uint32_t)(2 * a)), nibLen)); B.copy(shifted, &B.at(s.pos), 2 * a);
    B.at(outNib) = B.add(c_const(2), rm);
    B.at(outNib + 1) = B.mul(rm, B.at(shifted));
    for (size_t i = 0; i < 2 * a; i++) {
        if (i < 2 * a - 1) { Blk m = T_Mux1(B, B.at(shifted + i), B.at(shifted + i + 1), rm); B.at(outNib + i + 2) = B.at(m.pos); }
        else B.at(outNib + i + 2) = B.mul(B.not1(rm), B.at(shifted + i));
    }
    Blk nb = T_Nibbles2Bytes(B, ahb + 1, &B.at(outNib)); B.copy(o.pos, &B.at(nb.pos), a + 1);
    B.at(outLen) = B.add(ONE, B.at(div));
    return o;
}
// RlpMerklePatriciaTrieLeaf(maxAddressHashBytes, maxBalanceBytes) :102-189
static Blk T_RlpMerklePatriciaTrieLeaf(Builder &B, int mahb, int mbb, const Code *nibbles, Code nibLen, Code balance) {
    size_t mrea = 4 + (size_t)mbb + 66, mvr = 2 + mrea, mkl = 1 + (size_t)mahb, mkr = 1 + mkl, mpk = 2 + mkr, MO = mpk + mvr;
    Blk o = B.alloc(MO + 1 + 2 * (size_t)mahb + 1 + 1 + mkl + 1 + mrea + 1 + mpk + 1 + mvr + 1);
    size_t outLen = o.pos + MO, iNib = outLen + 1, iLen = iNib + 2 * (size_t)mahb, iBal = iLen + 1, key = iBal + 1, keyLen = key + mkl,
           rea = keyLen + 1, reaLen = rea + mrea, pk = reaLen + 1, pkLen = pk + mpk, vr = pkLen + 1, vrLen = vr + mvr;
    B.copy(iNib, nibbles, 2 * (size_t)mahb); B.at(iLen) = nibLen; B.at(iBal) = balance;
    Blk t = T_TruncatedAddressHash(B, mahb, nibbles, nibLen); B.copy(key, &B.at(t.pos), mkl); Code kl = B.at(t.pos + mkl); B.at(keyLen) = kl;
    T_AssertGreaterEqThan(B, 16, kl, c_const(2));
    Blk e = T_RlpEmptyAccount(B, mbb, balance); B.copy(rea, &B.at(e.pos), mrea); Code rl = B.at(e.pos + mrea); B.at(reaLen) = rl;
    B.at(vr) = c_const(0xb8); B.at(vr + 1) = rl;
    for (size_t i = 0; i < mrea; i++) B.at(vr + i + 2) = B.at(rea + i);
    Code vl = B.add(c_const(2), rl); B.at(vrLen) = vl;
    B.at(pk) = c_const(0xf8);
    B.at(pk + 1) = B.add(B.add(kl, ONE), vl);
    B.at(pk + 2) = B.add(c_const(0x80), kl);
    for (size_t i = 0; i < mkl; i++) B.at(pk + i + 3) = B.at(key + i);
    B.at(pkLen) = B.add(c_const(3), kl);
    Blk cc = T_Concat(B, (int)mpk, (int)mvr, &B.at(pk), B.at(pkLen), &B.at(vr), vl);
    B.copy(o.pos, &B.at(cc.pos), MO); B.at(outLen) = B.at(cc.pos + MO);
    return o;
}
// IsInRange(B) :196-207  own: out, lower, value, upper, lowerLteValue, valueLteUpper
static Blk T_IsInRange(Builder &B, int nb, Code lower, Code value, Code upper) {
    Blk o = B.alloc(6); B.at(o.pos + 1) = lower; B.at(o.pos + 2) = value; B.at(o.pos + 3) = upper;
    T_AssertBits(B, nb, lower); T_AssertBits(B, nb, value); T_AssertBits(B, nb, upper);
    Blk a = T_LessEqThan(B, nb, lower, value); B.at(o.pos + 4) = B.at(a.pos);
    Blk b = T_LessEqThan(B, nb, value, upper); B.at(o.pos + 5) = B.at(b.pos);
    B.at(o.pos) = B.mul(B.at(a.pos), B.at(b.pos)); return o;
}
// LeafDetector(N) :247-294
static Blk T_LeafDetector(Builder &B, int N, const Code *layer, Code layerLen) {
    size_t n = (size_t)N; Blk o = B.alloc(1 + n + 1 + 16);
    B.copy(o.pos + 1, layer, n); B.at(o.pos + 1 + n) = layerLen;
    size_t v = o.pos + 2 + n;
    T_AssertLessEqThan(B, 16, layerLen, c_const((uint32_t)n));
    B.at(v + 0) = B.at(T_IsEqual(B, layer[0], c_const(0xf8)).pos);                       // leafPrefixIsF8
    Code totalLength = layer[1]; B.at(v + 1) = totalLength;
    B.at(v + 2) = B.at(T_IsEqual(B, B.add(totalLength, c_const(2)), layerLen).pos);      // isConsistentWithLayerLen
    Code keyPrefix = layer[2]; B.at(v + 3) = keyPrefix;
    B.at(v + 4) = B.at(T_LessEqThan(B, 16, keyPrefix, c_const(0xb7)).pos);               // keyPrefixIsValid
    Code multi = B.at(T_IsInRange(B, 16, c_const(0x81), keyPrefix, c_const(0xb7)).pos); B.at(v + 5) = multi;
    Code extra = B.mul(multi, B.sub(keyPrefix, c_const(0x80))); B.at(v + 6) = extra;
    Code keyLen = B.add(ONE, extra); B.at(v + 7) = keyLen;
    Code base = B.add(c_const(2), keyLen);
    Code vwp = B.at(T_Selector(B, N, layer, base).pos); B.at(v + 8) = vwp;
    B.at(v + 9) = B.at(T_IsEqual(B, vwp, c_const(0xb8)).pos);
    Code vwl = B.at(T_Selector(B, N, layer, B.add(base, ONE)).pos); B.at(v + 10) = vwl;
    Code vp = B.at(T_Selector(B, N, layer, B.add(base, c_const(2))).pos); B.at(v + 11) = vp;
    B.at(v + 12) = B.at(T_IsEqual(B, vp, c_const(0xf8)).pos);
    Code vlen = B.at(T_Selector(B, N, layer, B.add(base, c_const(3))).pos); B.at(v + 13) = vlen;
    B.at(v + 14) = B.at(T_IsEqual(B, vwl, B.add(vlen, c_const(2))).pos);
    B.at(v + 15) = B.at(T_IsEqual(B, B.add(B.add(keyLen, vlen), c_const(6)), layerLen).pos);
    Code ands[7] = {B.at(v + 0), B.at(v + 2), B.at(v + 4), B.at(v + 9), B.at(v + 14), B.at(v + 12), B.at(v + 15)};
    B.at(o.pos) = B.at(T_MultiAND(B, 7, ands).pos);
    return o;
}

// ============================================================================================================
// circuits/spend.circom, circuits/proof_of_burn.circom
// ============================================================================================================
// Spend(maxAmountBytes) :32-53
static Blk T_Spend(Builder &B, int mab, Code burnKey, Code balance, Code withdrawn, Code extra) {
    Blk o = B.alloc(1 + 4 + 2 + 128);
    B.at(o.pos + 1) = burnKey; B.at(o.pos + 2) = balance; B.at(o.pos + 3) = withdrawn; B.at(o.pos + 4) = extra;
    size_t coin = o.pos + 5, rem = o.pos + 6, by = o.pos + 7;
    T_AssertGreaterEqThan(B, mab * 8, balance, withdrawn);
    Code i1[3] = {POSEIDON_PREFIX(B, 2), burnKey, balance};
    B.at(coin) = B.at(T_Poseidon(B, 3, i1).pos);
    Code i2[3] = {POSEIDON_PREFIX(B, 2), burnKey, B.sub(balance, withdrawn)};
    B.at(rem) = B.at(T_Poseidon(B, 3, i2).pos);
    Blk x = T_Num2BigEndianBytes(B, 32, B.at(coin)); B.copy(by, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, withdrawn); B.copy(by + 32, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, B.at(rem)); B.copy(by + 64, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, extra); B.copy(by + 96, &B.at(x.pos), 32);
    x = T_PublicCommitment(B, 4, &B.at(by)); B.at(o.pos) = B.at(x.pos);
    return o;
}
struct PobParams { int maxNumLayers, maxNodeBlocks, maxHeaderBlocks, minLeafAddressNibbles, amountBytes, powMinimumZeroBytes; Fr maxIntendedBalance, maxActualBalance; };
// ProofOfBurn(...) :34-212
static Blk T_ProofOfBurn(Builder &B, const PobParams &P, const Code *in) {
    size_t L = (size_t)P.maxNumLayers, NB = (size_t)P.maxNodeBlocks * 136, HB = (size_t)P.maxHeaderBlocks * 136;
    size_t nIn = 6 + L * NB + L + 1 + HB + 3;
    size_t nMid = 2 + 64 + 32 + 32 + 5 * 32 + NB + 1 + L + (L - 1) + L * 32 + L * 31 + L + 1 + 139 + 1;
    Blk o = B.alloc(1 + nIn + nMid);
    B.copy(o.pos + 1, in, nIn);
    const Code *I = &B.at(o.pos + 1);
    Code burnKey = I[0], actualBalance = I[1], intendedBalance = I[2], revealAmount = I[3], bec = I[4], numLeafNib = I[5];
    const Code *layers = I + 6, *layerLens = layers + L * NB;
    Code numLayers = layerLens[L];
    const Code *blockHeader = layerLens + L + 1;
    Code blockHeaderLen = blockHeader[HB], relax = blockHeader[HB + 1], proofExtra = blockHeader[HB + 2];
    size_t remainingCoin = o.pos + 1 + nIn, nullifier = remainingCoin + 1, addrNib = nullifier + 1, blockRoot = addrNib + 64, stateRoot = blockRoot + 32,
           nullB = stateRoot + 32, remB = nullB + 32, revB = remB + 32, becB = revB + 32, ecB = becB + 32, lastLayer = ecB + 32,
           lastLayerLen = lastLayer + NB, layerExists = lastLayerLen + 1, subChk = layerExists + L, layerKec = subChk + (L - 1),
           redKec = layerKec + L * 32, isLeaf = redKec + L * 31, isLastLeaf = isLeaf + L, leaf = isLastLeaf + 1, leafLen = leaf + 139;
    int ab8 = P.amountBytes * 8;
    T_AssertLessEqThan(B, ab8, intendedBalance, B.konst(P.maxIntendedBalance));
    T_AssertLessEqThan(B, ab8, actualBalance, B.konst(P.maxActualBalance));
    T_AssertLessEqThan(B, ab8, intendedBalance, actualBalance);
    Code relax2 = B.mul(relax, c_const(2)), minNib = c_const((uint32_t)P.minLeafAddressNibbles);
    T_AssertLessEqThan(B, 16, relax2, minNib);
    T_AssertGreaterEqThan(B, 16, numLeafNib, B.sub(minNib, relax2));
    T_AssertBits(B, ab8, revealAmount);
    T_AssertLessEqThan(B, ab8, revealAmount, intendedBalance);
    for (size_t i = 0; i < L; i++) {
        T_AssertLessThan(B, 16, layerLens[i], c_const((uint32_t)(NB * 8)));
        T_AssertByteString(B, (int)NB, layers + i * NB);
    }
    T_AssertLessThan(B, 16, blockHeaderLen, c_const((uint32_t)(HB * 8)));
    T_AssertByteString(B, (int)HB, blockHeader);
    Code p3[3] = {POSEIDON_PREFIX(B, 2), burnKey, B.sub(intendedBalance, revealAmount)};
    B.at(remainingCoin) = B.at(T_Poseidon(B, 3, p3).pos);
    Code p2[2] = {POSEIDON_PREFIX(B, 1), burnKey};
    B.at(nullifier) = B.at(T_Poseidon(B, 2, p2).pos);
    Blk x = T_BurnAddressHash(B, burnKey, revealAmount, bec); B.copy(addrNib, &B.at(x.pos), 64);
    x = T_KeccakBytes(B, P.maxHeaderBlocks, blockHeader, blockHeaderLen); B.copy(blockRoot, &B.at(x.pos), 32);
    for (size_t i = 0; i < 32; i++) B.at(stateRoot + i) = blockHeader[91 + i];
    x = T_Num2BigEndianBytes(B, 32, B.at(nullifier)); B.copy(nullB, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, B.at(remainingCoin)); B.copy(remB, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, revealAmount); B.copy(revB, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, bec); B.copy(becB, &B.at(x.pos), 32);
    x = T_Num2BigEndianBytes(B, 32, proofExtra); B.copy(ecB, &B.at(x.pos), 32);
    {
        std::vector<Code> six(192);
        memcpy(&six[0], &B.at(blockRoot), 128); memcpy(&six[32], &B.at(nullB), 128); memcpy(&six[64], &B.at(remB), 128);
        memcpy(&six[96], &B.at(revB), 128); memcpy(&six[128], &B.at(becB), 128); memcpy(&six[160], &B.at(ecB), 128);
        x = T_PublicCommitment(B, 6, six.data()); B.at(o.pos) = B.at(x.pos);
    }
    Code selLast = B.sub(numLayers, ONE);
    x = T_SelectorArray(B, P.maxNumLayers, NB, layers, selLast); B.copy(lastLayer, &B.at(x.pos), NB);
    B.at(lastLayerLen) = B.at(T_Selector(B, P.maxNumLayers, layerLens, selLast).pos);
    x = T_Filter(B, P.maxNumLayers, numLayers); B.copy(layerExists, &B.at(x.pos), L);
    Code numLeaves = ZERO;
    for (size_t i = 0; i < L; i++) {
        Code lf = B.at(T_LeafDetector(B, (int)NB, layers + i * NB, layerLens[i]).pos); B.at(isLeaf + i) = lf;
        numLeaves = B.add(numLeaves, lf);
        x = T_KeccakBytes(B, P.maxNodeBlocks, layers + i * NB, layerLens[i]); B.copy(layerKec + 32 * i, &B.at(x.pos), 32);
        x = T_Fit(B, 32, 31, &B.at(layerKec + 32 * i)); B.copy(redKec + 31 * i, &B.at(x.pos), 31);
        if (i > 0) {
            Code sc = B.at(T_SubstringCheck(B, (int)NB, 31, layers + (i - 1) * NB, layerLens[i - 1], &B.at(redKec + 31 * i)).pos);
            B.at(subChk + i - 1) = sc;
            B.chk_eq(B.mul(B.not1(sc), B.at(layerExists + i)), ZERO, o.sig);
        }
    }
    B.chk_eq(numLeaves, ONE, o.sig);
    B.at(isLastLeaf) = B.at(T_LeafDetector(B, (int)NB, &B.at(lastLayer), B.at(lastLayerLen)).pos);
    B.chk_eq(B.at(isLastLeaf), ONE, o.sig);
    for (size_t i = 0; i < 32; i++) B.chk_eq(B.at(layerKec + i), B.at(stateRoot + i), o.sig);
    x = T_RlpMerklePatriciaTrieLeaf(B, 32, P.amountBytes, &B.at(addrNib), numLeafNib, actualBalance);
    B.copy(leaf, &B.at(x.pos), 139); B.at(leafLen) = B.at(x.pos + 139);
    for (size_t i = 0; i < 139; i++) B.chk_eq(B.at(leaf + i), B.at(lastLayer + i), o.sig);
    B.chk_eq(B.at(leafLen), B.at(lastLayerLen), o.sig);
    T_ProofOfWorkChecker(B, burnKey, revealAmount, bec, B.add(c_const((uint32_t)P.powMinimumZeroBytes), relax));
    return o;
}

// ============================================================================================================
// main dispatch
// ============================================================================================================
struct MainInfo { const char *name; int nparams; const char *schema; };
static const MainInfo MAINS[] = {
    {"Spend", 1, "burnKey,balance,withdrawnBalance,extraCommitment"},
    {"ProofOfBurn", 8, "burnKey,actualBalance,intendedBalance,revealAmount,burnExtraCommitment,numLeafAddressNibbles,layers[p0][p1*136],layerLens[p0],numLayers,blockHeader[p2*136],blockHeaderLen,byteSecurityRelax,_proofExtraCommitment"},
    {"EIP7503", 0, ""}, {"ConcatFixed4", 4, "a[p0],b[p1],c[p2],d[p3]"},
    {"ProofOfWorkChecker", 0, "burnKey,revealAmount,burnExtraCommitment,minimumZeroBytes"},
    {"PublicCommitment", 1, "in[p0][32]"}, {"Poseidon", 1, "inputs[p0]"}, {"Divide", 1, "a,b"},
    {"SubstringCheck", 2, "mainInput[p0],mainLen,subInput[p1]"}, {"ShiftLeft", 1, "in[p0],count"},
    {"ShiftRight", 2, "in[p0],count"}, {"Mask", 1, "in[p0],count"}, {"Concat", 2, "a[p0],aLen,b[p1],bLen"},
    {"Selector", 1, "vals[p0],select"}, {"SelectorArray1D", 2, "arrays[p0][p1],select"},
    {"SelectorArray2D", 3, "arrays[p0][p1][p2],select"}, {"BigEndianBytes2Num", 1, "in[p0]"},
    {"LittleEndianBytes2Num", 1, "in[p0]"}, {"Bytes2Nibbles", 1, "in[p0]"}, {"Num2BigEndianBytes", 1, "in"},
    {"Num2LittleEndianBytes", 1, "in"}, {"Nibbles2Bytes", 1, "nibbles[2*p0]"}, {"Num2BitsSafe", 1, "in"},
    {"Pad", 2, "in[p0*p1],inLen"}, {"KeccakBytes", 1, "in[p0*136],inLen"},
    {"BurnAddress", 0, "burnKey,revealAmount,burnExtraCommitment"}, {"BurnAddressHash", 0, "burnKey,revealAmount,burnExtraCommitment"},
    {"AssertBits", 1, "in"}, {"AssertByteString", 1, "in[p0]"}, {"AssertLessThan", 1, "a,b"}, {"AssertLessEqThan", 1, "a,b"},
    {"AssertGreaterEqThan", 1, "a,b"}, {"Filter", 1, "in"}, {"Fit", 2, "in[p0]"}, {"Reverse", 1, "in[p0]"},
    {"Flatten", 2, "in[p0][p1]"}, {"Reshape", 2, "in[p0*p1]"}, {"RlpInteger", 1, "in"}, {"CountBytes", 1, "bytes[p0]"},
    {"RlpEmptyAccount", 1, "balance"}, {"TruncatedAddressHash", 1, "addressHashNibbles[2*p0],addressHashNibblesLen"},
    {"IsInRange", 1, "lower,value,upper"}, {"LeafDetector", 1, "layer[p0],layerLen"},
    {"RlpMerklePatriciaTrieLeaf", 2, "addressHashNibbles[2*p0],addressHashNibblesLen,balance"},
    {nullptr, 0, nullptr}};

static int PI(const std::vector<Fr> &p, size_t i) {
    if (i >= p.size()) throw std::runtime_error("pob: missing template parameter");
    return (int)p[i].l[0];
}
// number of scalar inputs of a main, from its schema
static size_t count_inputs(const char *schema, const std::vector<Fr> &p) {
    size_t total = 0; const char *s = schema;
    while (*s) {
        while (*s && *s != '[' && *s != ',') s++;
        size_t n = 1;
        while (*s == '[') {
            s++; size_t term = 1, acc = 0; bool have = false;       // expression: factors joined by '*'
            while (*s && *s != ']') {
                if (*s == 'p') { s++; term *= (size_t)PI(p, (size_t)(*s - '0')); s++; have = true; }
                else if (*s >= '0' && *s <= '9') { size_t v = 0; while (*s >= '0' && *s <= '9') v = v * 10 + (size_t)(*s++ - '0'); term *= v; have = true; }
                else if (*s == '*') s++;
                else throw std::runtime_error("pob: bad schema expression");
            }
            if (have) acc = term;
            n *= acc; if (*s == ']') s++;
        }
        total += n;
        if (*s == ',') s++;
    }
    return total;
}

static uint32_t run_main(Builder &B, const std::string &name, const std::vector<Fr> &p, const Code *in) {
    auto IS = [&](const char *s) { return name == s; };
    if (IS("Spend")) { T_Spend(B, PI(p, 0), in[0], in[1], in[2], in[3]); return 1; }
    if (IS("ProofOfBurn")) {
        if (p.size() < 8) throw std::runtime_error("pob: ProofOfBurn needs 8 parameters");
        PobParams P{PI(p, 0), PI(p, 1), PI(p, 2), PI(p, 3), PI(p, 4), PI(p, 5), p[6], p[7]};
        // proof_of_burn.circom:195-204 hard-codes maxLeafLen = 139 = the output length of RlpMerklePatriciaTrieLeaf(32, 31)
        // (108 + amountBytes) and compares 139 bytes of lastLayer: any other amountBytes, or layers shorter than 139
        // bytes, do not compile in the reference either
        if (P.maxNumLayers < 1 || P.maxNodeBlocks < 1 || P.maxHeaderBlocks < 1 || P.amountBytes != 31 || (size_t)P.maxNodeBlocks * 136 < 139 ||
            (size_t)P.maxHeaderBlocks * 136 < 123)
            throw std::runtime_error("pob: unsupported ProofOfBurn shape (needs amountBytes == 31, maxNodeBlocks*136 >= 139, maxHeaderBlocks*136 >= 123)");
        T_ProofOfBurn(B, P, in); return 1;
    }
    if (IS("EIP7503")) { T_EIP7503(B); return 8; }
    if (IS("ConcatFixed4")) { int A = PI(p, 0), Bn = PI(p, 1), C = PI(p, 2), D = PI(p, 3); T_ConcatFixed4(B, A, Bn, C, D, in, in + A, in + A + Bn, in + A + Bn + C); return (uint32_t)(A + Bn + C + D); }
    if (IS("ProofOfWorkChecker")) { T_ProofOfWorkChecker(B, in[0], in[1], in[2], in[3]); return 0; }
    if (IS("PublicCommitment")) { T_PublicCommitment(B, PI(p, 0), in); return 1; }
    if (IS("Poseidon")) { T_Poseidon(B, PI(p, 0), in); return 1; }
    if (IS("Divide")) { T_Divide(B, PI(p, 0), in[0], in[1]); return 2; }
    if (IS("SubstringCheck")) { int mm = PI(p, 0); T_SubstringCheck(B, mm, PI(p, 1), in, in[mm], in + mm + 1); return 1; }
    if (IS("ShiftLeft")) { int n = PI(p, 0); T_ShiftLeft(B, n, in, in[n]); return (uint32_t)n; }
    if (IS("ShiftRight")) { int n = PI(p, 0), ms = PI(p, 1); T_ShiftRight(B, n, ms, in, in[n]); return (uint32_t)(n + ms); }
    if (IS("Mask")) { int n = PI(p, 0); T_Mask(B, n, in, in[n]); return (uint32_t)n; }
    if (IS("Concat")) { int A = PI(p, 0), Bn = PI(p, 1); T_Concat(B, A, Bn, in, in[A], in + A + 1, in[A + 1 + Bn]); return (uint32_t)(A + Bn + 1); }
    if (IS("Selector")) { int n = PI(p, 0); T_Selector(B, n, in, in[n]); return 1; }
    if (IS("SelectorArray1D")) { int n = PI(p, 0), q = PI(p, 1); T_SelectorArray(B, n, (size_t)q, in, in[n * q]); return (uint32_t)q; }
    if (IS("SelectorArray2D")) { int n = PI(p, 0), q = PI(p, 1) * PI(p, 2); T_SelectorArray(B, n, (size_t)q, in, in[n * q]); return (uint32_t)q; }
    if (IS("BigEndianBytes2Num")) { T_BigEndianBytes2Num(B, PI(p, 0), in); return 1; }
    if (IS("LittleEndianBytes2Num")) { T_LittleEndianBytes2Num(B, PI(p, 0), in); return 1; }
    if (IS("Bytes2Nibbles")) { T_Bytes2Nibbles(B, PI(p, 0), in); return (uint32_t)(2 * PI(p, 0)); }
    if (IS("Num2BigEndianBytes")) { T_Num2BigEndianBytes(B, PI(p, 0), in[0]); return (uint32_t)PI(p, 0); }
    if (IS("Num2LittleEndianBytes")) { T_Num2LittleEndianBytes(B, PI(p, 0), in[0]); return (uint32_t)PI(p, 0); }
    if (IS("Nibbles2Bytes")) { T_Nibbles2Bytes(B, PI(p, 0), in); return (uint32_t)PI(p, 0); }
    if (IS("Num2BitsSafe")) { T_Num2BitsSafe(B, PI(p, 0), in[0]); return (uint32_t)PI(p, 0); }
    if (IS("Pad")) { int Bn = PI(p, 0) * PI(p, 1); T_Pad(B, PI(p, 0), PI(p, 1), in, in[Bn]); return (uint32_t)(Bn + 1); }
    if (IS("KeccakBytes")) { int Bn = PI(p, 0) * 136; T_KeccakBytes(B, PI(p, 0), in, in[Bn]); return 32; }
    if (IS("BurnAddress")) { T_BurnAddress(B, in[0], in[1], in[2]); return 20; }
    if (IS("BurnAddressHash")) { T_BurnAddressHash(B, in[0], in[1], in[2]); return 64; }
    if (IS("AssertBits")) { T_AssertBits(B, PI(p, 0), in[0]); return 0; }
    if (IS("AssertByteString")) { T_AssertByteString(B, PI(p, 0), in); return 0; }
    if (IS("AssertLessThan")) { T_AssertLessThan(B, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("AssertLessEqThan")) { T_AssertLessEqThan(B, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("AssertGreaterEqThan")) { T_AssertGreaterEqThan(B, PI(p, 0), in[0], in[1]); return 0; }
    if (IS("Filter")) { T_Filter(B, PI(p, 0), in[0]); return (uint32_t)PI(p, 0); }
    if (IS("Fit")) { T_Fit(B, PI(p, 0), PI(p, 1), in); return (uint32_t)PI(p, 1); }
    if (IS("Reverse")) { T_Reverse(B, PI(p, 0), in); return (uint32_t)PI(p, 0); }
    if (IS("Flatten") || IS("Reshape")) { size_t n = (size_t)PI(p, 0) * (size_t)PI(p, 1); T_CopyArray(B, n, in); return (uint32_t)n; }
    if (IS("RlpInteger")) { T_RlpInteger(B, PI(p, 0), in[0]); return (uint32_t)(PI(p, 0) + 2); }
    if (IS("CountBytes")) { T_CountBytes(B, PI(p, 0), in); return 1; }
    if (IS("RlpEmptyAccount")) { T_RlpEmptyAccount(B, PI(p, 0), in[0]); return (uint32_t)(4 + PI(p, 0) + 66 + 1); }
    if (IS("TruncatedAddressHash")) { int a = PI(p, 0); T_TruncatedAddressHash(B, a, in, in[2 * a]); return (uint32_t)(a + 2); }
    if (IS("IsInRange")) { T_IsInRange(B, PI(p, 0), in[0], in[1], in[2]); return 1; }
    if (IS("LeafDetector")) { int n = PI(p, 0); T_LeafDetector(B, n, in, in[n]); return 1; }
    if (IS("RlpMerklePatriciaTrieLeaf")) {
        int a = PI(p, 0), b = PI(p, 1); T_RlpMerklePatriciaTrieLeaf(B, a, b, in, in[2 * a], in[2 * a + 1]);
        return (uint32_t)((4 + a) + (2 + 4 + b + 66) + 1);
    }
    throw std::runtime_error("pob: unknown main template '" + name + "'");
}

struct BuildOut { uint32_t n_words; };

static void build(Builder &B, const std::string &name, const std::vector<Fr> &params, size_t n_in, uint32_t *n_out) {
    std::vector<Code> in(n_in ? n_in : 1);
    for (size_t i = 0; i < n_in; i++) { uint32_t s = B.new_val(0); in[i] = c_val(s); }
    *n_out = run_main(B, name, params, in.data());
}

}  // namespace

const char *main_input_schema(const std::string &main_name, int *nparams) {
    for (const MainInfo *m = MAINS; m->name; m++)
        if (main_name == m->name) { if (nparams) *nparams = m->nparams; return m->schema; }
    return nullptr;
}

Program compile_circuit(const std::string &main_name, const std::vector<Fr> &params, bool hcreate) {
    int np = 0; const char *schema = main_input_schema(main_name, &np);
    if (!schema) throw std::runtime_error("pob: unknown main template '" + main_name + "'");
    if ((int)params.size() < np) throw std::runtime_error("pob: too few template parameters for " + main_name);
    size_t n_in = count_inputs(schema, params);
    uint32_t n_out = 0;
    uint32_t n_words;
    { Builder dry(hcreate, true, 0); build(dry, main_name, params, n_in, &n_out); n_words = dry.n_words; }
    uint32_t val_base = (n_words + 3u) & ~3u;
    Builder B(hcreate, false, val_base);
    build(B, main_name, params, n_in, &n_out);

    Program P;
    P.main_name = main_name; P.params = params; P.hcreate = hcreate;
    P.n_signals = B.nsig; P.n_outputs = n_out; P.n_inputs = (uint32_t)n_in; P.input_schema = schema;
    P.n_words = B.n_words; P.val_base = val_base; P.n_vals = B.n_vals;
    if (P.store_u64() >= MAX_STORE_U64) throw std::runtime_error("pob: instance store exceeds the 128 MiB code range");
    P.aux = B.aux; P.konst = B.konsts;
    if (P.konst.empty()) P.konst.push_back(fr_zero());
    if (P.aux.empty()) P.aux.push_back(0);
    // ---- levelise ----
    // OP_INV results (the `inv` hint signal of IsZero) are consumed by no other op, so they are pulled out of the
    // dependency levels and run once at the end, batch-inverted (vm_inv_batch).
    std::vector<uint8_t> is_inv_slot(B.n_vals, 0);
    size_t n_inv = 0;
    for (auto &o : B.ops) if (op_opc(o.op) == OP_INV) { is_inv_slot[op_dst(o.op)] = 1; n_inv++; }
    auto uses_inv = [&](Code c) { return code_kind(c) == K_VAL && is_inv_slot[code_payload(c)]; };
    for (auto &o : B.ops) {
        uint32_t opc = op_opc(o.op);
        bool bad = false;
        if (opc == OP_FMA) bad = uses_inv(o.op.a) || uses_inv(o.op.b) || uses_inv(o.op.c);
        else if (opc == OP_CHK_EQ || opc == OP_DIV || opc == OP_MOD) bad = uses_inv(o.op.a) || uses_inv(o.op.b);
        else if (opc != OP_PACK8) bad = uses_inv(o.op.a);
        if (bad) throw std::runtime_error("pob: internal: an IsZero inverse is consumed by another op");
    }
    for (Code c : B.aux) if (uses_inv(c)) throw std::runtime_error("pob: internal: an IsZero inverse is consumed by an operand list");
    uint32_t max_level = 0;
    for (auto &o : B.ops) if (op_opc(o.op) != OP_INV) max_level = std::max(max_level, o.level);
    for (auto &a : B.absorbs) max_level = std::max(max_level, a.level);
    for (auto &q : B.poseidons) max_level = std::max(max_level, q.level);
    for (auto &q : B.psums) max_level = std::max(max_level, q.level);
    std::vector<uint32_t> scount(max_level + 2, 0), sstart(max_level + 2, 0);
    for (auto &q : B.psums) scount[q.level]++;
    for (uint32_t l = 1; l <= max_level + 1; l++) sstart[l] = sstart[l - 1] + scount[l - 1];
    P.psums.resize(B.psums.size());
    { std::vector<uint32_t> sp = sstart; for (auto &q : B.psums) P.psums[sp[q.level]++] = q.op; }
    for (auto &q : B.psums) if (uses_inv(q.op.x0)) throw std::runtime_error("pob: internal: an IsZero inverse feeds a prefix sum");
    std::vector<uint32_t> pcount(max_level + 2, 0), pstart(max_level + 2, 0);
    for (auto &q : B.poseidons) pcount[q.level]++;
    for (uint32_t l = 1; l <= max_level + 1; l++) pstart[l] = pstart[l - 1] + pcount[l - 1];
    P.poseidons.resize(B.poseidons.size());
    { std::vector<uint32_t> pp = pstart; for (auto &q : B.poseidons) P.poseidons[pp[q.level]++] = q.op; }
    P.pos_konst = B.pos_konst; if (P.pos_konst.empty()) P.pos_konst.push_back(fr_zero());
    for (auto &q : B.poseidons) for (uint32_t j = 0; j < q.op.t; j++) if (uses_inv(B.aux[q.op.in_aux + j])) throw std::runtime_error("pob: internal: an IsZero inverse feeds a Poseidon");
    std::vector<uint32_t> tcount(max_level + 2, 0), wcount(max_level + 2, 0);
    for (auto &o : B.ops) if (op_opc(o.op) != OP_INV) tcount[o.level]++;
    for (auto &a : B.absorbs) wcount[a.level]++;
    std::vector<uint32_t> tstart(max_level + 2, 0), wstart(max_level + 2, 0);
    for (uint32_t l = 1; l <= max_level + 1; l++) { tstart[l] = tstart[l - 1] + tcount[l - 1]; wstart[l] = wstart[l - 1] + wcount[l - 1]; }
    P.ops.resize(B.ops.size()); P.absorbs.resize(B.absorbs.size());
    P.inv_begin = (uint32_t)(B.ops.size() - n_inv); P.inv_end = (uint32_t)B.ops.size();
    B.inv_generic.resize(B.n_vals, 0);
    size_t n_ginv = 0; for (auto &o : B.ops) if (op_opc(o.op) == OP_INV && B.inv_generic[op_dst(o.op)]) n_ginv++;
    P.ginv_begin = (uint32_t)(P.inv_end - n_ginv);
    { std::vector<uint32_t> tp = tstart, wp = wstart; uint32_t ip = P.inv_begin, gp = P.ginv_begin;
      for (auto &o : B.ops) {
          if (op_opc(o.op) == OP_INV) { if (B.inv_generic[op_dst(o.op)]) P.ops[gp++] = o.op; else P.ops[ip++] = o.op; }
          else P.ops[tp[o.level]++] = o.op;
      }
      for (auto &a : B.absorbs) P.absorbs[wp[a.level]++] = a.op; }
    // inside a level: long sequential ops first, then grouped by opcode / operand class so that the 32 lanes of a
    // warp run the same case of the interpreter switch (and the same fast or slow multiplication path)
    auto op_key = [](const Op &o) -> uint32_t {
        uint32_t opc = op_opc(o);
        uint32_t k = (opc + 2) << 4;
        if (opc == OP_FMA) k |= (o.b == c_const(1)) ? 0u : (o.b == c_konst(0)) ? 1u : (code_kind(o.b) == K_KONST) ? 3u : 2u;
        return k;
    };
    for (uint32_t l = 1; l <= max_level; l++) {
        if (tcount[l] == 0 && wcount[l] == 0 && pcount[l] == 0 && scount[l] == 0) continue;
        std::stable_sort(P.ops.begin() + tstart[l], P.ops.begin() + tstart[l] + tcount[l], [&](const Op &x, const Op &y) { return op_key(x) < op_key(y); });
        P.levels.push_back(Level{tstart[l], tstart[l] + tcount[l], wstart[l], wstart[l] + wcount[l], pstart[l], pstart[l] + pcount[l], sstart[l], sstart[l] + scount[l]});
    }
    // ---- codes + tiles ----
    P.codes.resize(ROUND_SIGNALS);
    { LaneSink S{P.codes.data(), 0}; emit_round(S);
      if ((size_t)(S.p - P.codes.data()) != ROUND_SIGNALS) throw std::runtime_error("pob: internal: round table size mismatch"); }
    // derive (and thereby verify) the 64-signal group descriptors of the round table
    P.round_desc.resize(ROUND_SIGNALS / 64);
    for (uint32_t g = 0; g < ROUND_SIGNALS / 64; g++) {
        const Code *c = &P.codes[64 * g];
        auto W = [&](int j) { return (code_payload(c[j]) >> 6); };
        auto Bt = [&](int j) { return (code_payload(c[j]) & 63u); };
        bool ok = true, lane = true;
        for (int j = 0; j < 64; j++) { if (code_kind(c[j]) != K_BIT) ok = false; if (W(j) != W(0) || Bt(j) != (uint32_t)j) lane = false; }
        uint64_t d = 0;
        if (ok && lane) d = (uint64_t)W(0);
        else if (ok) {
            bool found = false;
            for (uint32_t f = 0; f < 3 && !found; f++) {
                uint32_t w[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}; bool m = true;
                for (uint32_t j = 0; j < 64 && m; j++) {
                    uint32_t sidx = 64 * f + j, gi = sidx / 3, mem = sidx % 3;
                    if (Bt((int)j) != gi) m = false;
                    else if (w[mem] == 0xffffffffu) w[mem] = W((int)j);
                    else if (w[mem] != W((int)j)) m = false;
                }
                if (m) { d = (uint64_t)w[0] | ((uint64_t)w[1] << 16) | ((uint64_t)w[2] << 32) | ((uint64_t)(1 + f) << 48); found = true; }
            }
            ok = found;
        }
        // k_expand_round indexes its shared-memory word table with all three descriptor words
        if (!ok || (d & 0xffff) >= ROUND_WORDS_SPAN || ((d >> 48) && (((d >> 16) & 0xffff) >= ROUND_WORDS_SPAN || ((d >> 32) & 0xffff) >= ROUND_WORDS_SPAN)))
            throw std::runtime_error("pob: internal: round table group does not fit a descriptor");
        P.round_desc[g] = d;
    }
    P.codes.insert(P.codes.end(), B.flat, B.flat + B.flat_n);
    P.n_round_blocks = B.n_round_blocks; P.n_flat_signals = B.flat_n;
    uint32_t tile_signals = TILE_SIGNALS;
#ifdef POB_TUNING
    if (const char *v = getenv("POB_TILE_SIGNALS")) { uint32_t t = (uint32_t)atoi(v); if (t >= 64 && t <= MAX_TILE_SIGNALS && t % 64 == 0) tile_signals = t; }
#endif
    for (auto &s : B.segs) {
        uint64_t done = 0;
        while (done < s.n) {
            uint32_t n = (uint32_t)std::min<uint64_t>(tile_signals, s.n - done);
            Tile t; t.dst = s.dst + done; t.n = n; t.pad = 0;
            if (s.round) { t.code_off = (uint32_t)done; t.ubase = s.ubase; t.pad = 1; }
            else { t.code_off = (uint32_t)(ROUND_SIGNALS + s.pos + done); t.ubase = 0; }
            P.tiles.push_back(t); done += n;
        }
    }
    // Tile order is free (every tile carries its own destination).  KeccakfRound tiles only read L1-resident tables, the
    // other tiles read their code stream and store values through L2/DRAM; interleaving those reads with the write
    // stream costs DRAM efficiency (profiles/r01_expand_sweep.md), so all round tiles go first, the rest last.
    std::stable_sort(P.tiles.begin(), P.tiles.end(), [](const Tile &a, const Tile &b) { return a.pad > b.pad; });
    return P;
}

std::vector<Fr> build_inverse_table() {
    const uint32_t N = 1u << 16;
    std::vector<Fr> inv(N), pre(N);
    Fr acc = fr_from_u64(1);
    for (uint32_t i = 1; i < N; i++) { pre[i] = acc; acc = fr_mul(acc, fr_from_u64(i)); }
    Fr ai = fr_inv(acc);
    for (uint32_t i = N - 1; i >= 1; i--) { inv[i] = fr_mul(ai, pre[i]); ai = fr_mul(ai, fr_from_u64(i)); }
    inv[0] = fr_zero();
    return inv;
}

}  // namespace pob
