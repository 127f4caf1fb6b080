// pob_b200.cu -- CUDA kernels (sm_100a) and the C-ABI of the batched witness generator (include/pob_b200.h).
//
// Kernels
//   k_eval    one CTA per proof instance: runs the levelised witness program over the instance store.
//             Thread ops: BN254-Fr FMA / IsZero / inverse / div-mod / byte packing / constraint checks
//             (vm_exec.h).  Warp ops: one Keccak absorb per warp, state lane l held by thread l, Theta
//             column parities / D, RhoPi lane walk and Chi neighbours exchanged with warp shuffles; every
//             intermediate lane word the bit-level circuit exposes (utils/keccak.circom:58-297) is written
//             to the store (238 words per round).  One Poseidon permutation per warp (state element j in
//             Montgomery form on lane j, Mix/MixS via shuffles).  Prefix sums by warp scan.  IsZero inverse
//             hints batch-inverted at the end (table for small inputs, one binary-EEA inversion per thread).
//   k_pow_grind  proof-of-work burn-key search (the step before the path): one candidate key per thread.
//   k_expand_round / k_expand_codes  the HBM-bound kernels: materialise every witness entry as a 32-byte little-endian
//             field element with one 256-bit store (STG.E.ENL2.256).  Algorithmic bytes = 32 * n_signals per instance
//             (6.909 GB for main_proof_of_burn).  KeccakfRound blocks (95.8 %) are driven by 8-byte group descriptors
//             and the round's lane words staged in shared memory; everything else by one 32-bit code per entry.
//   k_digest  optional 64-bit digest of a materialised witness (parity tests at full size).
// There is no host execution path for any of this: without a device pob_create fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/pob_b200.h"
#include "compiler.h"
#include "vm_exec.h"

using namespace pob;

// =============================================================================================================
// device code
// =============================================================================================================
namespace {

__device__ __forceinline__ int inv_rot(int L) {      // i such that rot[i + 1] == L  (L = 1..24)
    constexpr int INV[25] = {0, 23, 17, 5, 11, 6, 22, 1, 8, 21, 0, 2, 16, 15, 19, 12, 7, 3, 4, 14, 18, 9, 20, 13, 10};
    return INV[L];
}

// One Absorb (utils/keccak.circom:304-323) by one warp; lane l < 25 owns state lane l.
__device__ void absorb_warp(uint64_t *W, const AbsorbOp op) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool act = lane < 25;
    const int l = act ? lane : 0;
    uint64_t st = (act && op.s_idx != NONE_IDX) ? W[op.s_idx + l] : 0ull;
    if (lane < 17) st ^= W[op.blk_idx + lane];
    if (act) W[op.out_idx + l] = st;
    const int col = l % 5;
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t *B = W + op.out_idx + RW * r;
        // Theta: Xor5 chain of my column (all lanes of a column compute it redundantly; lanes 0..4 store it)
        uint64_t v0 = __shfl_sync(FULL, st, col), v1 = __shfl_sync(FULL, st, col + 5), v2 = __shfl_sync(FULL, st, col + 10),
                 v3 = __shfl_sync(FULL, st, col + 15), v4 = __shfl_sync(FULL, st, col + 20);
        uint64_t x0 = v0 ^ v1, x1 = x0 ^ v2, x2 = x1 ^ v3, c = x2 ^ v4;
        if (lane < 5) { B[rw_x5(lane, 0)] = x0; B[rw_x5(lane, 1)] = x1; B[rw_x5(lane, 2)] = x2; B[rw_x5(lane, 3)] = c; }
        // D(i) = c[(i+4)%5] ^ rotl(c[(i+1)%5], 1)
        uint64_t ca = __shfl_sync(FULL, c, (col + 1) % 5), cb = __shfl_sync(FULL, c, (col + 4) % 5);
        uint64_t s0 = ca << 1, s1 = ca >> 63, so = s0 | s1, d = cb ^ so;
        if (lane < 5) { B[rw_dd(lane, 0)] = s0; B[rw_dd(lane, 1)] = s1; B[rw_dd(lane, 2)] = so; B[rw_dd(lane, 3)] = d; }
        uint64_t th = st ^ d;
        if (act) B[rw_th(l)] = th;
        // RhoPi: lane i < 24 performs step i on theta[rot[i]], the result belongs to lane rot[i+1]
        const int i = lane < 24 ? lane : 0;
        uint64_t a = __shfl_sync(FULL, th, keccak_rot(i));
        const int shl = keccak_shl(i);
        uint64_t a0 = a >> (64 - shl), a1 = a << shl, ro = a0 | a1;
        if (lane < 24) { B[rw_rp(lane, 0)] = a0; B[rw_rp(lane, 1)] = a1; B[rw_rp(lane, 2)] = ro; }
        uint64_t rp = __shfl_sync(FULL, ro, inv_rot(l));
        if (lane == 0) rp = th;
        // Chi
        uint64_t vb = __shfl_sync(FULL, rp, chi_b(l)), vc = __shfl_sync(FULL, rp, chi_c(l));
        uint64_t nb = ~vb, bc = nb & vc, ch = rp ^ bc;
        if (act) { B[rw_ch(l, 0)] = nb; B[rw_ch(l, 1)] = bc; B[rw_ch(l, 2)] = ch; }
        // Iota
        const uint64_t rc = keccak_rc(r);
        if (lane == 0) { B[RW_RC] = rc; ch ^= rc; }
        if (act) B[rw_out(l)] = ch;
        st = ch;
    }
}

// ---- one Poseidon permutation by one warp (circomlib/circuits/poseidon.circom:67-196) ---------------------------
// Lane j < t owns state element j in Montgomery form; Mix / MixS exchange elements with warp shuffles; every
// intermediate signal is converted back to canonical form and written to its slot (layout: program.h PosLayout).
__device__ __forceinline__ Fr shfl_fr(const Fr &v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
__device__ void poseidon_warp(const VmCtx &x, const PoseidonOp op, const Fr *pk) {
    const int lane = threadIdx.x & 31;
    const uint32_t t = op.t; const PosLayout L = pos_layout(t);
    const bool act = (uint32_t)lane < t; const uint32_t j = act ? (uint32_t)lane : 0u;
    const Fr *K = pk + op.koff;
    uint64_t *V = x.U + x.val_base + 4ull * op.base;
    // values are parked in their slots in MONTGOMERY form (nothing but this warp reads them before the sweep below):
    // no conversion sits on the dependency chain of the 65 rounds
    auto put = [&](uint32_t off, const Fr &m) { if (act) vm_store_val(V + 4ull * off, m); };
    Fr s = fr_add(fr_to_mont(vm_load(x, x.aux[op.in_aux + j])), K[L.kC + j]);     // ark[0]
    put(j, s);
    auto full = [&](uint32_t F, uint32_t coff, uint32_t moff) {
        Fr x2 = fr_mont(s, s), x4 = fr_mont(x2, x2), x5 = fr_mont(x4, s);
        put(F + 3 * j, x2); put(F + 3 * j + 1, x4); put(F + 3 * j + 2, x5);
        Fr y = fr_add(x5, K[L.kC + coff + j]); put(F + 3 * t + j, y);
        Fr acc = fr_zero();
        for (uint32_t k = 0; k < t; k++) { Fr yk = shfl_fr(y, (int)k); acc = fr_add(acc, fr_mont(K[moff + k * t + j], yk)); }
        put(F + 4 * t + j, acc); s = acc;
    };
    for (uint32_t f = 0; f < 4; f++) full(L.F1 + 5 * t * f, (f + 1) * t, f == 3 ? L.kP : L.kM);
#pragma unroll 1
    for (uint32_t r = 0; r < L.rp; r++) {
        const uint32_t B = L.PB + r * (4 + t), so = (2 * t - 1) * r;
        Fr x2 = fr_mont(s, s), x4 = fr_mont(x2, x2), x5 = fr_mont(x4, s);                // meaningful on lane 0 only
        Fr z0 = fr_add(x5, K[L.kC + 5 * t + r]);
        if (lane == 0) { vm_store_val(V + 4ull * B, x2); vm_store_val(V + 4ull * (B + 1), x4); vm_store_val(V + 4ull * (B + 2), x5); vm_store_val(V + 4ull * (B + 3), z0); }
        z0 = shfl_fr(z0, 0);
        const Fr in = (lane == 0) ? z0 : s;
        Fr prod = fr_mont(K[L.kS + so + j], in);                                          // S[so + i] * in[i]
        Fr o0 = fr_zero();
        for (uint32_t k = 0; k < t; k++) o0 = fr_add(o0, shfl_fr(prod, (int)k));
        Fr oj = fr_add(s, fr_mont(z0, K[L.kS + so + t + (j ? j : 1) - 1]));                // lanes 1..t-1
        s = (lane == 0) ? o0 : oj;
        put(B + 4 + j, s);
    }
    for (uint32_t f = 0; f < 3; f++) full(L.SB + 5 * t * f, 5 * t + L.rp + f * t, L.kM);
    {
        Fr x2 = fr_mont(s, s), x4 = fr_mont(x2, x2), x5 = fr_mont(x4, s);
        put(L.LB + 3 * j, x2); put(L.LB + 3 * j + 1, x4); put(L.LB + 3 * j + 2, x5);
        Fr prod = fr_mont(K[L.kM + j * t], x5), out = fr_zero();
        for (uint32_t k = 0; k < t; k++) out = fr_add(out, shfl_fr(prod, (int)k));
        if (lane == 0) vm_store_val(V + 4ull * (L.LB + 3 * t), out);
    }
    __syncwarp();
    for (uint32_t i = (uint32_t)lane; i < L.total; i += 32) {         // all 32 lanes: Montgomery -> canonical, in place
        uint64_t *p = V + 4ull * i;
        vm_store_val(p, fr_from_mont(vm_load_val(p)));
    }
}

// ---- prefix sum with every partial sum a signal (substring_check.circom:47-49, :95) by one warp -----------------
__device__ void psum_warp(const VmCtx &x, const PsumOp op) {
    const uint32_t lane = threadIdx.x & 31, per = (op.n + 31) / 32;
    const uint32_t lo = min(op.n, lane * per), hi = min(op.n, lo + per);
    Fr acc = fr_zero();
    for (uint32_t k = lo; k < hi; k++) acc = fr_add(acc, vm_load(x, x.aux[op.aux0 + k]));
    Fr incl = acc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_up_sync(0xffffffffu, incl.l[i], off);
        if ((int)lane >= off) incl = fr_add(incl, o);
    }
    Fr run = fr_add(vm_load(x, op.x0), fr_sub(incl, acc));            // x0 + sum of all earlier lanes
    uint64_t *V = x.U + x.val_base + 4ull * op.dst;
    for (uint32_t k = lo; k < hi; k++) { run = fr_add(run, vm_load(x, x.aux[op.aux0 + k])); vm_store_val(V + 4ull * k, run); }
}

struct EvalArgs {
    const Op *ops; const AbsorbOp *absorbs; const PoseidonOp *poseidons; const Fr *pos_konst; const PsumOp *psums;
    const Level *levels; uint32_t n_levels, inv_begin, ginv_begin, inv_end;
    const Code *aux; const Fr *konst; const Fr *invtab;
    const Code *out_codes; uint32_t n_outputs, n_inputs, val_base;
    uint64_t *stores; uint64_t store_stride;     // u64 units
    const uint64_t *inputs;                       // chunk base: instance j at inputs + j * n_inputs * 4
    uint32_t *status; uint64_t *outputs;          // chunk base
    long long *prof;                              // tuning only: per-level clock64 stamps of instance 0 (or null)
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_eval(const EvalArgs a) {
    const uint32_t inst = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    uint64_t *U = a.stores + (uint64_t)inst * a.store_stride;
    __shared__ uint32_t s_status;
    if (tid == 0) s_status = STATUS_OK;
    const uint64_t *in = a.inputs + (uint64_t)inst * a.n_inputs * 4;
    // inputs -> first value slots, reduced mod p like the circom loader does (a caller may hand over limbs >= p)
    for (uint32_t i = tid; i < a.n_inputs; i += nthr) {
        Fr v = vm_load_val(in + 4ull * i);
        while (fr_geq_p(v)) { Fr t; fr_raw_sub(t, v, fr_p()); v = t; }
        vm_store_val(U + a.val_base + 4ull * i, v);
    }
    __syncthreads();
    VmCtx x{U, a.val_base, a.konst, a.aux, a.invtab, &s_status};
    const uint32_t warp = tid >> 5, nwarp = nthr >> 5;
    for (uint32_t lv = 0; lv < a.n_levels; lv++) {
        if (a.prof && inst == 0 && tid == 0) a.prof[lv] = clock64();
        const Level L = a.levels[lv];
        for (uint32_t i = L.t_begin + tid; i < L.t_end; i += nthr) vm_exec_op(x, a.ops[i]);
        // warp ops: Poseidons take the first warps (long), Keccak absorbs the next ones
        for (uint32_t q = L.p_begin + warp; q < L.p_end; q += nwarp) poseidon_warp(x, a.poseidons[q], a.pos_konst);
        { const uint32_t np = (L.p_end - L.p_begin) % nwarp, wv = (warp + nwarp - np) % nwarp;
          for (uint32_t w = L.w_begin + wv; w < L.w_end; w += nwarp) absorb_warp(U, a.absorbs[w]);
          const uint32_t nw2 = (np + (L.w_end - L.w_begin)) % nwarp, sv = (warp + nwarp - nw2) % nwarp;
          for (uint32_t q = L.s_begin + sv; q < L.s_end; q += nwarp) psum_warp(x, a.psums[q]); }
        __syncthreads();
    }
    if (a.prof && inst == 0 && tid == 0) a.prof[a.n_levels] = clock64();
    // IsZero inverse hints: no consumers, done last.  Table-sized inputs are spread over all threads; the ones expected
    // to need a real inversion go to 256 threads so that only 8 warps pay for a Fermat ladder (one per thread).
    vm_inv_batch(x, a.ops, a.inv_begin, a.ginv_begin, tid, nthr);
    if (tid < 256) vm_inv_batch(x, a.ops, a.ginv_begin, a.inv_end, tid, 256);
    if (a.prof && inst == 0) { __syncthreads(); if (tid == 0) a.prof[a.n_levels + 1] = clock64(); }
    if (tid == 0) a.status[inst] = (s_status == STATUS_OK) ? 0u : s_status;
    for (uint32_t i = tid; i < a.n_outputs; i += nthr) {
        uint64_t v[4]; vm_expand(a.out_codes[i], U, 0, a.val_base, a.konst, v);
        uint64_t *o = a.outputs + ((uint64_t)inst * a.n_outputs + i) * 4;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
    }
}

// 256-bit store (STG.E.ENL2.256).  No "memory" clobber on purpose: the compiler must be free to hoist the next
// entries' loads above it so that several loads are in flight per thread (the witness is written, never read, here).
__device__ __forceinline__ void st256(uint64_t *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d));
}

struct ExpandArgs {
    const Tile *tiles; const Code *codes; const Fr *konst; const uint2 *round_desc;
    const uint64_t *stores; uint64_t store_stride; uint32_t val_base;
    uint64_t *const *wit;                         // per instance of the group: witness slot base
    uint32_t tile0;                               // first tile of this launch (k_expand_codes)
};

// k_expand_round: grid = (KeccakfRound tiles, instances in the group) -- 95.8 % of the witness.  One CTA streams one
// tile (<= 8192 entries = 256 KiB) with one 256-bit store per entry.  The source of every entry follows from an 8-byte
// descriptor per 64 entries and a lane word of the round; the tile's <= 128 descriptors and the round's 263 words are
// staged in shared memory in one burst, so the streaming loop touches no global memory but the witness itself.
template <int T>
__global__ void __launch_bounds__(T) k_expand_round(const ExpandArgs a) {
    const Tile t = a.tiles[blockIdx.x];
    const uint64_t *Ub = a.stores + (uint64_t)blockIdx.y * a.store_stride + t.ubase;
    uint64_t *W = a.wit[blockIdx.y] + t.dst * 4;
    __shared__ uint2 sD[MAX_TILE_SIGNALS / 64]; __shared__ uint64_t sW[ROUND_WORDS_SPAN + 1];
    const uint2 *D = a.round_desc + (t.code_off >> 6);
    for (uint32_t i = threadIdx.x; i < ((t.n + 63) >> 6); i += T) sD[i] = __ldg(D + i);
    for (uint32_t i = threadIdx.x; i < ROUND_WORDS_SPAN; i += T) sW[i] = Ub[i];
    __syncthreads();
#pragma unroll 8
    for (uint32_t k = threadIdx.x; k < t.n; k += T) {
        const uint2 d = sD[k >> 6];
        const uint32_t tt = k & 63, mode = d.y >> 16;
        uint32_t w = d.x & 0xffffu, b = tt;
        if (mode) {                          // phase (mode-1) of a gate block [out_i, a_i, b_i]_i
            const uint32_t sidx = (mode - 1) * 64 + tt, g = sidx / 3, m = sidx - 3 * g;
            b = g; w = (m == 0) ? (d.x & 0xffffu) : (m == 1) ? (d.x >> 16) : (d.y & 0xffffu);
        }
        st256(W + 4ull * k, (sW[w] >> b) & 1ull, 0, 0, 0);
    }
}

// k_expand_codes: grid = (instances in the group, code tiles) -- INSTANCE-major, so that a tile's code stream is fetched
// from DRAM once and served from L2 to the other witnesses of the group.  One 32-bit code per entry, loads issued 4
// entries ahead of the stores.
__global__ void __launch_bounds__(256, 5) k_expand_codes(const ExpandArgs a) {
    const uint32_t inst = blockIdx.x;
    const Tile t = a.tiles[a.tile0 + blockIdx.y];
    const uint64_t *U = a.stores + (uint64_t)inst * a.store_stride;
    uint64_t *W = a.wit[inst] + t.dst * 4;
    const uint64_t *Ub = U + t.ubase;
    const Code *c = a.codes + t.code_off;
    constexpr int UG = 4;
    for (uint32_t base = threadIdx.x; base < t.n; base += 256 * UG) {
        Code cd[UG];
#pragma unroll
        for (int u = 0; u < UG; u++) { const uint32_t k = base + 256 * u; cd[u] = k < t.n ? __ldg(c + k) : 0u; }
        uint64_t v[UG][4];
#pragma unroll
        for (int u = 0; u < UG; u++) {
            const uint32_t kind = code_kind(cd[u]), p = code_payload(cd[u]);
            v[u][1] = v[u][2] = v[u][3] = 0;
            if (kind == K_BIT) v[u][0] = (Ub[p >> 6] >> (p & 63)) & 1ull;
            else if (kind == K_CONST) v[u][0] = p;
            else {
                const uint64_t *s = (kind == K_VAL) ? U + a.val_base + 4ull * p : reinterpret_cast<const uint64_t *>(a.konst + p);
                v[u][0] = s[0]; v[u][1] = s[1]; v[u][2] = s[2]; v[u][3] = s[3];
            }
        }
#pragma unroll
        for (int u = 0; u < UG; u++) { const uint32_t k = base + 256 * u; if (k < t.n) st256(W + 4ull * k, v[u][0], v[u][1], v[u][2], v[u][3]); }
    }
}

// digest = sum_i mix(i, limbs) mod 2^64 -- must equal oracle/pob_oracle.c:pob_oracle_digest
__global__ void __launch_bounds__(256) k_digest(const uint64_t *wit, uint64_t n_signals, unsigned long long *out) {
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_signals; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong4 v = *reinterpret_cast<const ulonglong4 *>(wit + 4 * i);
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ULL;
        h ^= v.x * 0xBF58476D1CE4E5B9ULL + v.y * 0x94D049BB133111EBULL + v.z * 0xD6E8FEB86659FD93ULL + v.w * 0xA0761D6478BD642FULL;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
        acc += h;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ uint64_t part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t s = 0; for (int w = 0; w < 8; w++) s += part[w]; atomicAdd(out, (unsigned long long)s); }
}

// ---- proof-of-work burn-key grinder (reference tests/main.py:47-56; circuits/utils/proof_of_work.circom:54-81) ----
// One candidate key per thread: a single-block keccak256 of key|reveal|extra|"EIP-7503" held in 25 registers.
struct GrindArgs { uint64_t start[4], lanes_tail[9]; uint64_t first, count; uint32_t zero_bytes; unsigned long long *hit; };
__device__ __forceinline__ uint64_t bswap64(uint64_t x) { return __byte_perm((uint32_t)(x >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32); }
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
__global__ void __launch_bounds__(256) k_pow_grind(const GrindArgs a) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.count) return;
    // key = start + first + idx (256-bit add)
    uint64_t k0 = a.start[0], k1 = a.start[1], k2 = a.start[2], k3 = a.start[3];
    const uint64_t add = a.first + idx;
    k0 += add; if (k0 < add) { if (++k1 == 0) { if (++k2 == 0) ++k3; } }
    uint64_t s[25];
    s[0] = bswap64(k3); s[1] = bswap64(k2); s[2] = bswap64(k1); s[3] = bswap64(k0);      // 32-byte big-endian key
#pragma unroll
    for (int i = 0; i < 9; i++) s[4 + i] = a.lanes_tail[i];                                // reveal | extra | "EIP-7503"
    s[13] = 0x01; s[14] = 0; s[15] = 0; s[16] = 0x8000000000000000ull;                     // 0x01 ... 0x80 padding of a 104-byte message
#pragma unroll
    for (int i = 17; i < 25; i++) s[i] = 0;
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
        b[0] = s[0];
#pragma unroll
        for (int i = 0; i < 24; i++) b[keccak_rot(i + 1)] = rotl64(s[keccak_rot(i)], keccak_shl(i));
#pragma unroll
        for (int i = 0; i < 25; i++) s[i] = b[i] ^ (~b[chi_b(i)] & b[chi_c(i)]);
        s[0] ^= keccak_rc(r);
    }
    const uint64_t mask = a.zero_bytes >= 8 ? ~0ull : ((1ull << (8 * a.zero_bytes)) - 1);
    if ((s[0] & mask) == 0) atomicMin(a.hit, (unsigned long long)idx);
}

// ---- self-check: every KeccakfRound block of a materialised witness satisfies out == KeccakRound_r(in) -------------
// One warp per block.  Lane l < 25 assembles lane word l of `in` (witness entries base+1600+64l .. +63) and of `out`
// (base+64l ..) from the 32-byte entries, lane 0 gathers the 25 input words and runs one textbook round.
__global__ void __launch_bounds__(256) k_check_rounds(const uint64_t *wit, const uint64_t *block_base, uint32_t n_blocks, unsigned long long *n_bad) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_blocks) return;
    const uint64_t base = block_base[warp];
    const int r = (int)(warp % 24);                       // round blocks are emitted 24 per Keccakf, in order
    uint64_t win = 0, wout = 0; bool bad = false;
    if (lane < 25) {
        for (uint32_t k = 0; k < 64; k++) {
            const ulonglong4 a = *reinterpret_cast<const ulonglong4 *>(wit + 4 * (base + 1600 + 64 * lane + k));
            const ulonglong4 b = *reinterpret_cast<const ulonglong4 *>(wit + 4 * (base + 64 * lane + k));
            bad |= (a.x > 1) | (b.x > 1) | ((a.y | a.z | a.w | b.y | b.z | b.w) != 0);
            win |= (a.x & 1ull) << k; wout |= (b.x & 1ull) << k;
        }
    }
    uint64_t s[25], o[25];
#pragma unroll
    for (int i = 0; i < 25; i++) { s[i] = __shfl_sync(0xffffffffu, win, i); o[i] = __shfl_sync(0xffffffffu, wout, i); }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        uint64_t c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
        b[0] = s[0];
#pragma unroll
        for (int i = 0; i < 24; i++) b[keccak_rot(i + 1)] = rotl64(s[keccak_rot(i)], keccak_shl(i));
#pragma unroll
        for (int i = 0; i < 25; i++) s[i] = b[i] ^ (~b[chi_b(i)] & b[chi_c(i)]);
        s[0] ^= keccak_rc(r);
#pragma unroll
        for (int i = 0; i < 25; i++) bad |= (s[i] != o[i]);
        if (bad) atomicAdd(n_bad, 1ull);
    }
}

}  // namespace

// =============================================================================================================
// host side of the C-ABI
// =============================================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct pob_handle {
    Program P; int device = 0;
    // device program
    Op *d_ops = nullptr; PsumOp *d_psums = nullptr; PoseidonOp *d_pos = nullptr; Fr *d_pos_konst = nullptr; AbsorbOp *d_abs = nullptr; Level *d_levels = nullptr; Code *d_aux = nullptr; Fr *d_konst = nullptr;
    Code *d_codes = nullptr; Tile *d_tiles = nullptr; Fr *d_invtab = nullptr; uint64_t *d_round_desc = nullptr;
    // stores (ring of RING chunks)
    static const uint32_t RING = 2;
    uint32_t chunk = 0; uint64_t store_stride = 0; uint64_t *d_stores = nullptr; uint64_t *d_inputs = nullptr;
    // witness slots
    std::vector<uint64_t *> slots;
    // per-batch buffers
    uint32_t cap_n = 0; uint32_t *d_status = nullptr; uint64_t *d_outputs = nullptr; unsigned long long *d_digests = nullptr;
    uint64_t **d_witptr = nullptr; uint32_t *h_status = nullptr; uint64_t *h_outputs = nullptr; uint64_t *h_digests = nullptr;
    uint64_t **h_witptr = nullptr;
    uint64_t *d_staged = nullptr; uint32_t n_staged = 0;
    long long *d_prof = nullptr;               // POB_EVAL_PROFILE: per-level clock stamps (tuning only)
    uint32_t xgroup = 0;                       // instances per expand launch (distinct witness slots)
    uint32_t n_round_tiles = 0;                // tiles [0, n_round_tiles) are KeccakfRound tiles, the rest code tiles
    uint64_t *d_block_base = nullptr; uint32_t n_blocks = 0;   // witness index of every KeccakfRound block (self-check)
    // k_expand_round is launched with 85 KiB of (unused) dynamic shared memory so that only TWO of its CTAs are resident
    // per SM: fewer concurrent write streams give the DRAM controllers longer same-row bursts -- measured 7.35 TB/s with
    // 2 CTAs/SM vs 7.26 / 7.19 / 7.09 / 6.97 TB/s with 3 / 4 / 5 / 8, and 5.96 TB/s with 1 (profiles/r01_expand_sweep.md)
    uint32_t round_dyn_smem = 85 * 1024;
    uint32_t round_threads = 256;              // CTA size of k_expand_round (POB_EXPAND_THREADS), tuning only
    uint32_t codes_dyn_smem = 0;               // same occupancy cap for k_expand_codes (POB_CODES_SMEM_KB), tuning only
    int eval_threads = 1024;                   // k_eval CTA size (POB_EVAL_THREADS), tuning only
    bool serialize = false;                    // POB_SERIALIZE=1: eval and expand on one stream (no overlap), tuning only
    cudaStream_t s_eval = nullptr, s_exp = nullptr, s_h2d = nullptr;
    cudaEvent_t ev_eval_done[RING] = {nullptr, nullptr}, ev_exp_done[RING] = {nullptr, nullptr}, ev_h2d[RING] = {nullptr, nullptr},
                ev_start = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    uint32_t last_n = 0; bool last_expanded = false;
    pob_timing timing{};
};

static void fill_desc(const Program &P, pob_desc *d) {
    memset(d, 0, sizeof *d);
    d->n_signals = P.n_signals; d->n_outputs = P.n_outputs; d->n_inputs = P.n_inputs;
    d->witness_bytes = 32ull * P.n_signals; d->wtns_file_bytes = 76ull + 32ull * P.n_signals;
    d->store_bytes = 8ull * P.store_u64(); d->n_ops = P.ops.size(); d->n_absorbs = (uint32_t)P.absorbs.size();
    d->n_levels = (uint32_t)P.levels.size(); d->n_tiles = (uint32_t)P.tiles.size();
}
static std::vector<Fr> params_vec(const uint64_t *params, int nparams) {
    std::vector<Fr> ps((size_t)(nparams > 0 ? nparams : 0));
    for (int i = 0; i < nparams; i++) memcpy(ps[(size_t)i].l, params + 4 * i, 32);
    return ps;
}
template <class T> static T *upload(const std::vector<T> &v) {
    T *d = nullptr; size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    if (!v.empty()) CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

extern "C" {

const char *pob_last_error(void) { return g_err.c_str(); }
const char *pob_version(void) { return "pob_b200 0.1 (sm_100a)"; }

const char *pob_input_schema(const char *main_name, int *nparams) {
    if (!main_name) return nullptr;
    return main_input_schema(main_name, nparams);
}

int pob_layout_info(const char *main_name, const uint64_t *params, int nparams, int hcreate, pob_desc *out) {
    if (!main_name || !out || (nparams > 0 && !params)) return fail(POB_E_BAD_ARG, "pob_layout_info: null argument");
    try { Program P = compile_circuit(main_name, params_vec(params, nparams), hcreate != 0); fill_desc(P, out); }
    catch (const std::exception &e) { return fail(POB_E_COMPILE, e.what()); }
    return POB_OK;
}

void pob_destroy(pob_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->s_eval) cudaStreamSynchronize(h->s_eval);
    if (h->s_exp) cudaStreamSynchronize(h->s_exp);
    for (void *p : {(void *)h->d_ops, (void *)h->d_psums, (void *)h->d_pos, (void *)h->d_pos_konst, (void *)h->d_abs, (void *)h->d_levels, (void *)h->d_aux, (void *)h->d_konst, (void *)h->d_codes,
                    (void *)h->d_tiles, (void *)h->d_invtab, (void *)h->d_round_desc, (void *)h->d_stores, (void *)h->d_inputs, (void *)h->d_status,
                    (void *)h->d_outputs, (void *)h->d_digests, (void *)h->d_witptr, (void *)h->d_staged, (void *)h->d_prof, (void *)h->d_block_base})
        if (p) cudaFree(p);
    for (uint64_t *s : h->slots) cudaFree(s);
    for (void *p : {(void *)h->h_status, (void *)h->h_outputs, (void *)h->h_digests, (void *)h->h_witptr}) if (p) cudaFreeHost(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (uint32_t r = 0; r < pob_handle::RING; r++) {
        if (h->ev_eval_done[r]) cudaEventDestroy(h->ev_eval_done[r]);
        if (h->ev_exp_done[r]) cudaEventDestroy(h->ev_exp_done[r]);
        if (h->ev_h2d[r]) cudaEventDestroy(h->ev_h2d[r]);
    }
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    if (h->s_eval) cudaStreamDestroy(h->s_eval);
    if (h->s_exp) cudaStreamDestroy(h->s_exp);
    delete h;
}

int pob_create(const char *main_name, const uint64_t *params, int nparams, int hcreate, int device, uint32_t max_slots, pob_handle **out) {
    if (!main_name || !out || (nparams > 0 && !params)) return fail(POB_E_BAD_ARG, "pob_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(POB_E_NO_DEVICE, "pob_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(POB_E_NO_DEVICE, "pob_create: device index out of range");
    pob_handle *h = new pob_handle(); h->device = device;
    try { h->P = compile_circuit(main_name, params_vec(params, nparams), hcreate != 0); }
    catch (const std::exception &e) { delete h; return fail(POB_E_COMPILE, e.what()); }
    try {
        const Program &P = h->P;
        CU(cudaSetDevice(device));
        h->d_ops = upload(P.ops); h->d_psums = upload(P.psums); h->d_pos = upload(P.poseidons); h->d_pos_konst = upload(P.pos_konst); h->d_abs = upload(P.absorbs); h->d_levels = upload(P.levels); h->d_aux = upload(P.aux);
        h->d_konst = upload(P.konst); h->d_codes = upload(P.codes);
        if (const char *v = getenv("POB_TILE_FILTER")) {      // tuning only: 1 = KeccakfRound tiles only, 2 = the others only (witness incomplete!)
            std::vector<Tile> sub; for (const Tile &t : P.tiles) if ((atoi(v) == 1) == (t.pad != 0)) sub.push_back(t);
            h->P.tiles = sub;
        }
        if (getenv("POB_FLAT_FIRST")) std::stable_sort(h->P.tiles.begin(), h->P.tiles.end(), [](const Tile &a, const Tile &b) { return a.pad < b.pad; });   // tuning only
        h->d_tiles = upload(h->P.tiles);
        for (const Tile &t : h->P.tiles) if (t.pad) h->n_round_tiles++;
        { std::vector<uint64_t> bases; for (const Tile &t : P.tiles) if (t.pad && t.code_off == 0) bases.push_back(t.dst);
          std::sort(bases.begin(), bases.end()); h->n_blocks = (uint32_t)bases.size(); h->d_block_base = upload(bases); }
        h->d_invtab = upload(build_inverse_table());
        h->d_round_desc = upload(P.round_desc);
        // the small eval grid must get SMs while the expand grid (hundreds of thousands of CTAs) is draining:
        // eval runs on the highest-priority stream, expand on the lowest
        int pr_least = 0, pr_greatest = 0; CU(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
        CU(cudaStreamCreateWithPriority(&h->s_eval, cudaStreamNonBlocking, pr_greatest));
        CU(cudaStreamCreateWithPriority(&h->s_h2d, cudaStreamNonBlocking, pr_greatest));
        CU(cudaStreamCreateWithPriority(&h->s_exp, cudaStreamNonBlocking, pr_least));
        for (uint32_t r = 0; r < pob_handle::RING; r++) {
            CU(cudaEventCreateWithFlags(&h->ev_eval_done[r], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_exp_done[r], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_h2d[r], cudaEventDisableTiming));
        }
        CU(cudaEventCreate(&h->ev_start)); CU(cudaEventCreate(&h->ev_end));
        // (an L2 persisting access-policy window for the code stream was tried and REDUCED the expand kernel to 4.9 TB/s: the
        // carve-out takes L2 away from write combining -- profiles/r01_expand_sweep.md)
        if (getenv("POB_EVAL_PROFILE")) CU(cudaMalloc(&h->d_prof, (P.levels.size() + 2) * sizeof(long long)));
        // witness slots: as many as fit in 80 % of free HBM after the store ring
        size_t free_b = 0, total_b = 0; CU(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t wbytes = 32ull * P.n_signals;
        h->store_stride = std::max<uint64_t>(32, (P.store_u64() + 31) & ~31ull);     // never 0 (constant-only gadgets such as EIP7503())
        // eval chunk: enough instances per launch to keep the SMs busy on small circuits, bounded by a ~0.5 GB store ring
        // half (main_proof_of_burn: 32; Spend: 1024)
        uint32_t chunk = (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(32, (512ull << 20) / (h->store_stride * 8)));
        chunk -= chunk % 32;
        if (const char *v = getenv("POB_EVAL_CHUNK")) chunk = (uint32_t)std::max(1, atoi(v));
        const uint64_t ring_bytes_per_inst = pob_handle::RING * (h->store_stride * 8 + (uint64_t)P.n_inputs * 32);
        uint64_t budget = (uint64_t)(free_b * 0.8);
        uint64_t nslots = budget > chunk * ring_bytes_per_inst ? (budget - chunk * ring_bytes_per_inst) / wbytes : 0;
        if (max_slots && nslots > max_slots) nslots = max_slots;
        if (nslots > 4096) nslots = 4096;
        if (nslots == 0) throw std::runtime_error("not even one witness slot fits in free HBM");
        // expand group: ~100 GB of witness per launch pair (main_proof_of_burn: 16 witnesses; Spend: up to the whole chunk)
        h->xgroup = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(nslots, chunk), std::max<uint64_t>(16, (100ull << 30) / wbytes));
        if (const char *v = getenv("POB_EVAL_THREADS")) h->eval_threads = atoi(v);
        if (const char *v = getenv("POB_SERIALIZE")) h->serialize = atoi(v) != 0;
        if (const char *v = getenv("POB_EXPAND_SMEM_KB")) h->round_dyn_smem = (uint32_t)atoi(v) * 1024u;
        if (const char *v = getenv("POB_EXPAND_THREADS")) h->round_threads = (uint32_t)atoi(v);
        if (const char *v = getenv("POB_CODES_SMEM_KB")) { h->codes_dyn_smem = (uint32_t)atoi(v) * 1024u; if (h->codes_dyn_smem > 48 * 1024) CU(cudaFuncSetAttribute(k_expand_codes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->codes_dyn_smem)); }
        if (h->round_dyn_smem > 48 * 1024) {
            CU(cudaFuncSetAttribute(k_expand_round<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->round_dyn_smem));
            CU(cudaFuncSetAttribute(k_expand_round<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->round_dyn_smem));
            CU(cudaFuncSetAttribute(k_expand_round<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->round_dyn_smem));
            CU(cudaFuncSetAttribute(k_expand_round<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->round_dyn_smem));
        }
        if (const char *v = getenv("POB_EXPAND_GROUP")) h->xgroup = (uint32_t)std::max(1, std::min<int>(atoi(v), (int)std::min<uint64_t>(nslots, chunk)));
        h->chunk = chunk;
        CU(cudaMalloc(&h->d_stores, (size_t)pob_handle::RING * chunk * h->store_stride * 8));
        CU(cudaMalloc(&h->d_inputs, std::max<size_t>(32, (size_t)pob_handle::RING * chunk * P.n_inputs * 32)));
        for (uint64_t s = 0; s < nslots; s++) { uint64_t *p = nullptr; CU(cudaMalloc(&p, wbytes)); h->slots.push_back(p); }
    } catch (const std::exception &e) {
        std::string m = e.what(); pob_destroy(h);
        return fail(m.find("slot") != std::string::npos ? POB_E_NO_MEMORY : POB_E_CUDA, "pob_create: " + m);
    }
    *out = h; return POB_OK;
}

int pob_describe(const pob_handle *h, pob_desc *out) {
    if (!h || !out) return fail(POB_E_BAD_ARG, "pob_describe: null argument");
    fill_desc(h->P, out); out->n_slots = (uint32_t)h->slots.size(); out->chunk = h->chunk; out->expand_group = h->xgroup; return POB_OK;
}

void *pob_alloc_pinned(uint64_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; } return p; }
void pob_free_pinned(void *p) { if (p) cudaFreeHost(p); }

static void ensure_batch_buffers(pob_handle *h, uint32_t n) {
    if (n <= h->cap_n) return;
    const Program &P = h->P;
    for (void *p : {(void *)h->d_status, (void *)h->d_outputs, (void *)h->d_digests, (void *)h->d_witptr}) if (p) cudaFree(p);
    for (void *p : {(void *)h->h_status, (void *)h->h_outputs, (void *)h->h_digests, (void *)h->h_witptr}) if (p) cudaFreeHost(p);
    const size_t no = std::max<uint32_t>(1, P.n_outputs);
    CU(cudaMalloc(&h->d_status, (size_t)n * 4)); CU(cudaMalloc(&h->d_outputs, (size_t)n * no * 32));
    CU(cudaMalloc(&h->d_digests, (size_t)n * 8)); CU(cudaMalloc(&h->d_witptr, (size_t)n * sizeof(uint64_t *)));
    CU(cudaMallocHost(&h->h_status, (size_t)n * 4)); CU(cudaMallocHost(&h->h_outputs, (size_t)n * no * 32));
    CU(cudaMallocHost(&h->h_digests, (size_t)n * 8)); CU(cudaMallocHost(&h->h_witptr, (size_t)n * sizeof(uint64_t *)));
    h->cap_n = n;
}

int pob_stage_inputs(pob_handle *h, const uint64_t *inputs, uint32_t n) {
    if (!h || !inputs || n == 0) return fail(POB_E_BAD_ARG, "pob_stage_inputs: bad argument");
    try {
        CU(cudaSetDevice(h->device));
        if (h->d_staged) { cudaFree(h->d_staged); h->d_staged = nullptr; h->n_staged = 0; }
        size_t bytes = std::max<size_t>(32, (size_t)n * h->P.n_inputs * 32);
        CU(cudaMalloc(&h->d_staged, bytes));
        if (h->P.n_inputs) CU(cudaMemcpy(h->d_staged, inputs, (size_t)n * h->P.n_inputs * 32, cudaMemcpyHostToDevice));
        h->n_staged = n;
    } catch (const std::exception &e) { return fail(POB_E_CUDA, e.what()); }
    return POB_OK;
}

int pob_run_batch(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags, uint32_t *status, uint64_t *outputs, uint64_t *digests) {
    if (!h || n == 0 || !status) return fail(POB_E_BAD_ARG, "pob_run_batch: bad argument");
    const bool staged = (flags & POB_RUN_INPUTS_STAGED) != 0, expand = (flags & (POB_RUN_EXPAND | POB_RUN_DIGEST)) != 0, digest = (flags & POB_RUN_DIGEST) != 0;
    if (staged ? (h->n_staged < n) : (inputs == nullptr && h->P.n_inputs)) return fail(POB_E_BAD_ARG, "pob_run_batch: no inputs");
    if (digest && !digests) return fail(POB_E_BAD_ARG, "pob_run_batch: POB_RUN_DIGEST needs a digests array");
    try {
        const Program &P = h->P;
        CU(cudaSetDevice(h->device));
        ensure_batch_buffers(h, n);
        const uint32_t E = h->chunk, R = pob_handle::RING, nchunks = (n + E - 1) / E, nslots = (uint32_t)h->slots.size(), X = h->xgroup;
        const size_t in_stride = (size_t)P.n_inputs * 4, no = std::max<uint32_t>(1, P.n_outputs);
        const uint32_t groups_per_chunk = (E + X - 1) / X, ev_per_chunk = 2 + 2 * groups_per_chunk;
        while (h->ev_pool.size() < (size_t)nchunks * ev_per_chunk) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->ev_pool.push_back(e); }
        for (uint32_t i = 0; i < n; i++) h->h_witptr[i] = h->slots[i % nslots];
        pob_timing T{};
        CU(cudaEventRecord(h->ev_start, h->s_eval));
        CU(cudaMemcpyAsync(h->d_witptr, h->h_witptr, (size_t)n * sizeof(uint64_t *), cudaMemcpyHostToDevice, h->s_eval));
        if (digest) CU(cudaMemsetAsync(h->d_digests, 0, (size_t)n * 8, h->s_eval));
        CU(cudaEventRecord(h->ev_eval_done[0], h->s_eval));
        CU(cudaStreamWaitEvent(h->s_h2d, h->ev_eval_done[0], 0));
        std::vector<uint32_t> group_count;
        for (uint32_t c = 0; c < nchunks; c++) {
            const uint32_t r = c % R, first = c * E, cnt = std::min(E, n - first);
            cudaEvent_t *ev = &h->ev_pool[(size_t)c * ev_per_chunk];
            const uint64_t *d_in;
            if (staged) d_in = h->d_staged + (size_t)first * in_stride;
            else {
                // inputs travel on their own stream, one chunk ahead of the eval kernel that consumes them
                uint64_t *dst = h->d_inputs + (size_t)r * E * in_stride;
                if (c >= R) CU(cudaStreamWaitEvent(h->s_h2d, h->ev_eval_done[r], 0));
                if (in_stride) { CU(cudaMemcpyAsync(dst, inputs + (size_t)first * in_stride, (size_t)cnt * in_stride * 8, cudaMemcpyHostToDevice, h->s_h2d)); T.h2d_bytes += (uint64_t)cnt * in_stride * 8; }
                CU(cudaEventRecord(h->ev_h2d[r], h->s_h2d));
                CU(cudaStreamWaitEvent(h->s_eval, h->ev_h2d[r], 0));
                d_in = dst;
            }
            if (c >= R) CU(cudaStreamWaitEvent(h->s_eval, h->ev_exp_done[r], 0));      // store ring slot r is free again
            uint64_t *stores = h->d_stores + (size_t)r * E * h->store_stride;
            EvalArgs ea{h->d_ops, h->d_abs, h->d_pos, h->d_pos_konst, h->d_psums, h->d_levels, (uint32_t)P.levels.size(), P.inv_begin, P.ginv_begin, P.inv_end, h->d_aux, h->d_konst, h->d_invtab,
                        h->d_codes + ROUND_SIGNALS + 1, P.n_outputs, P.n_inputs, P.val_base, stores, h->store_stride, d_in,
                        h->d_status + first, h->d_outputs + (size_t)first * no * 4, (c == 0) ? h->d_prof : nullptr};
            CU(cudaEventRecord(ev[0], h->s_eval));
            switch (h->eval_threads) {
            case 256: k_eval<256><<<cnt, 256, 0, h->s_eval>>>(ea); break;
            case 512: k_eval<512><<<cnt, 512, 0, h->s_eval>>>(ea); break;
            default: k_eval<1024><<<cnt, 1024, 0, h->s_eval>>>(ea); break;
            }
            CU(cudaEventRecord(ev[1], h->s_eval));
            CU(cudaEventRecord(h->ev_eval_done[r], h->s_eval));
            T.eval_launches++;
            if (expand) {
                CU(cudaStreamWaitEvent(h->s_exp, h->ev_eval_done[r], 0));
                uint32_t g = 0;
                for (uint32_t off = 0; off < cnt; off += X, g++) {
                    const uint32_t gc = std::min(X, cnt - off);
                    ExpandArgs xa{h->d_tiles, h->d_codes, h->d_konst, reinterpret_cast<const uint2 *>(h->d_round_desc), stores + (size_t)off * h->store_stride, h->store_stride, P.val_base, h->d_witptr + first + off, 0};
                    CU(cudaEventRecord(ev[2 + 2 * g], h->s_exp));
                    // launch 1: KeccakfRound tiles, tile-major (each CTA's tables are L1-resident);
                    // launch 2: code tiles, INSTANCE-major, so that a tile's code stream is fetched from DRAM once and
                    // served from L2 to the other witnesses of the group
                    const uint32_t n_round = h->n_round_tiles, n_code = (uint32_t)P.tiles.size() - n_round;
                    if (n_round) {
                        xa.tile0 = 0;
                        if (h->round_threads == 128) k_expand_round<128><<<dim3(n_round, gc), 128, h->round_dyn_smem, h->s_exp>>>(xa);
                        else if (h->round_threads == 512) k_expand_round<512><<<dim3(n_round, gc), 512, h->round_dyn_smem, h->s_exp>>>(xa);
                        else if (h->round_threads == 1024) k_expand_round<1024><<<dim3(n_round, gc), 1024, h->round_dyn_smem, h->s_exp>>>(xa);
                        else k_expand_round<256><<<dim3(n_round, gc), 256, h->round_dyn_smem, h->s_exp>>>(xa);
                    }
                    if (n_code) { xa.tile0 = n_round; k_expand_codes<<<dim3(gc, n_code), 256, h->codes_dyn_smem, h->s_exp>>>(xa); T.other_launches++; }
                    CU(cudaEventRecord(ev[3 + 2 * g], h->s_exp));
                    T.expand_launches++;
                    if (digest) for (uint32_t j = 0; j < gc; j++) {
                        k_digest<<<1184, 256, 0, h->s_exp>>>(h->slots[(first + off + j) % nslots], P.n_signals, h->d_digests + first + off + j);
                        T.other_launches++;
                    }
                }
                group_count.push_back(g);
                CU(cudaEventRecord(h->ev_exp_done[r], h->s_exp));
                if (h->serialize) CU(cudaStreamWaitEvent(h->s_eval, h->ev_exp_done[r], 0));
            } else {
                group_count.push_back(0);
                CU(cudaEventRecord(h->ev_exp_done[r], h->s_eval));
            }
        }
        CU(cudaGetLastError());
        // results: status + outputs after the last eval; digests after the last expand
        CU(cudaMemcpyAsync(h->h_status, h->d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, h->s_eval));
        T.d2h_bytes += (uint64_t)n * 4;
        if (P.n_outputs) { CU(cudaMemcpyAsync(h->h_outputs, h->d_outputs, (size_t)n * no * 32, cudaMemcpyDeviceToHost, h->s_eval)); T.d2h_bytes += (uint64_t)n * no * 32; }
        CU(cudaEventRecord(h->ev_eval_done[0], h->s_eval));
        CU(cudaStreamWaitEvent(h->s_exp, h->ev_eval_done[0], 0));
        if (digest) { CU(cudaMemcpyAsync(h->h_digests, h->d_digests, (size_t)n * 8, cudaMemcpyDeviceToHost, h->s_exp)); T.d2h_bytes += (uint64_t)n * 8; }
        CU(cudaEventRecord(h->ev_end, h->s_exp));
        CU(cudaStreamSynchronize(h->s_exp)); CU(cudaStreamSynchronize(h->s_eval)); CU(cudaStreamSynchronize(h->s_h2d));
        CU(cudaGetLastError());
        memcpy(status, h->h_status, (size_t)n * 4);
        if (outputs && P.n_outputs) memcpy(outputs, h->h_outputs, (size_t)n * no * 32);
        if (digest) memcpy(digests, h->h_digests, (size_t)n * 8);
        CU(cudaEventElapsedTime(&T.total_ms, h->ev_start, h->ev_end));
        for (uint32_t c = 0; c < nchunks; c++) {
            cudaEvent_t *ev = &h->ev_pool[(size_t)c * ev_per_chunk];
            float ms = 0; CU(cudaEventElapsedTime(&ms, ev[0], ev[1])); T.eval_ms += ms;
            for (uint32_t g = 0; g < group_count[c]; g++) { CU(cudaEventElapsedTime(&ms, ev[2 + 2 * g], ev[3 + 2 * g])); T.expand_ms += ms; }
        }
        if (h->d_prof) {
            std::vector<long long> st(P.levels.size() + 2);
            CU(cudaMemcpy(st.data(), h->d_prof, st.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            FILE *f = fopen(getenv("POB_EVAL_PROFILE"), "w");
            if (f) {
                for (size_t l = 0; l + 1 < st.size(); l++) {
                    if (l < P.levels.size()) fprintf(f, "level %zu ops %u absorbs %u poseidons %u cycles %lld\n", l, P.levels[l].t_end - P.levels[l].t_begin, P.levels[l].w_end - P.levels[l].w_begin, P.levels[l].p_end - P.levels[l].p_begin, st[l + 1] - st[l]);
                    else fprintf(f, "inverse-batch ops %u cycles %lld\n", P.inv_end - P.inv_begin, st[l + 1] - st[l]);
                }
                fclose(f);
            }
        }
        h->timing = T; h->last_n = n; h->last_expanded = expand;
    } catch (const std::exception &e) { return fail(POB_E_CUDA, std::string("pob_run_batch: ") + e.what()); }
    return POB_OK;
}

int pob_pow_grind(int device, const uint64_t start_key[4], const uint64_t reveal_amount[4], const uint64_t burn_extra_commitment[4],
                  uint32_t zero_bytes, uint64_t max_tries, uint64_t found_key[4], uint64_t *tries) {
    if (!start_key || !reveal_amount || !burn_extra_commitment || !found_key || zero_bytes > 8 || max_tries == 0)
        return fail(POB_E_BAD_ARG, "pob_pow_grind: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(POB_E_NO_DEVICE, "pob_pow_grind: no such CUDA device");
    unsigned long long *d_hit = nullptr;
    try {
        CU(cudaSetDevice(device));
        GrindArgs ga; memcpy(ga.start, start_key, 32); ga.zero_bytes = zero_bytes;
        uint8_t msg[72];                                     // bytes 32..103 of the message
        for (int i = 0; i < 32; i++) { msg[i] = (uint8_t)(reveal_amount[3 - i / 8] >> (8 * (7 - i % 8))); msg[32 + i] = (uint8_t)(burn_extra_commitment[3 - i / 8] >> (8 * (7 - i % 8))); }
        memcpy(msg + 64, "EIP-7503", 8);
        memcpy(ga.lanes_tail, msg, 72);                      // little-endian lanes of a little-endian host
        CU(cudaMalloc(&d_hit, 8));
        ga.hit = d_hit;
        const uint64_t WINDOW = 1ull << 22;
        for (uint64_t first = 0; first < max_tries; first += WINDOW) {
            unsigned long long none = ~0ull, hit = ~0ull;
            CU(cudaMemcpy(d_hit, &none, 8, cudaMemcpyHostToDevice));
            ga.first = first; ga.count = std::min<uint64_t>(WINDOW, max_tries - first);
            k_pow_grind<<<(unsigned)((ga.count + 255) / 256), 256>>>(ga);
            CU(cudaGetLastError());
            CU(cudaMemcpy(&hit, d_hit, 8, cudaMemcpyDeviceToHost));
            if (hit != ~0ull) {
                uint64_t add = first + hit, k[4] = {start_key[0], start_key[1], start_key[2], start_key[3]};
                k[0] += add; if (k[0] < add) { if (++k[1] == 0) { if (++k[2] == 0) ++k[3]; } }
                memcpy(found_key, k, 32); if (tries) *tries = add + 1;
                cudaFree(d_hit); return POB_OK;
            }
        }
        cudaFree(d_hit);
    } catch (const std::exception &e) { if (d_hit) cudaFree(d_hit); return fail(POB_E_CUDA, std::string("pob_pow_grind: ") + e.what()); }
    if (tries) *tries = max_tries;
    return fail(POB_E_RANGE, "pob_pow_grind: no key in the search window satisfies the proof-of-work check");
}

int pob_last_timing(const pob_handle *h, pob_timing *out) {
    if (!h || !out) return fail(POB_E_BAD_ARG, "pob_last_timing: null argument");
    *out = h->timing; return POB_OK;
}

static int resident_slot(pob_handle *h, uint32_t index, uint64_t **slot) {
    if (!h->last_expanded || index >= h->last_n) return fail(POB_E_RANGE, "witness index not in the last expanded batch");
    const uint32_t nslots = (uint32_t)h->slots.size();
    if ((uint64_t)index + nslots < h->last_n) return fail(POB_E_RANGE, "witness slot already overwritten by a later instance");
    *slot = h->slots[index % nslots]; return POB_OK;
}

int pob_selfcheck_keccak(pob_handle *h, uint32_t index, uint64_t *n_blocks, uint64_t *n_bad) {
    if (!h || !n_blocks || !n_bad) return fail(POB_E_BAD_ARG, "pob_selfcheck_keccak: null argument");
    uint64_t *s = nullptr; int rc = resident_slot(h, index, &s); if (rc) return rc;
    unsigned long long *d_bad = nullptr, bad = 0;
    try {
        CU(cudaSetDevice(h->device));
        CU(cudaMalloc(&d_bad, 8)); CU(cudaMemset(d_bad, 0, 8));
        if (h->n_blocks) k_check_rounds<<<(h->n_blocks * 32 + 255) / 256, 256>>>(s, h->d_block_base, h->n_blocks, d_bad);
        CU(cudaGetLastError());
        CU(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
        cudaFree(d_bad);
    } catch (const std::exception &e) { if (d_bad) cudaFree(d_bad); return fail(POB_E_CUDA, std::string("pob_selfcheck_keccak: ") + e.what()); }
    *n_blocks = h->n_blocks; *n_bad = bad;
    return POB_OK;
}

int pob_debug_poke_witness(pob_handle *h, uint32_t index, uint64_t signal, const uint64_t value[4]) {
    if (!h || !value) return fail(POB_E_BAD_ARG, "pob_debug_poke_witness: null argument");
    uint64_t *s = nullptr; int rc = resident_slot(h, index, &s); if (rc) return rc;
    if (signal >= h->P.n_signals) return fail(POB_E_RANGE, "pob_debug_poke_witness: signal index out of range");
    if (cudaSetDevice(h->device) != cudaSuccess || cudaMemcpy(s + 4 * signal, value, 32, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(POB_E_CUDA, "pob_debug_poke_witness: cudaMemcpy failed");
    return POB_OK;
}

int pob_witness_device_ptr(pob_handle *h, uint32_t index, void **dptr) {
    if (!h || !dptr) return fail(POB_E_BAD_ARG, "pob_witness_device_ptr: null argument");
    uint64_t *s = nullptr; int rc = resident_slot(h, index, &s); if (rc) return rc;
    *dptr = s; return POB_OK;
}

int pob_copy_witness(pob_handle *h, uint32_t index, uint64_t first_signal, uint64_t n_signals, uint64_t *dst_host) {
    if (!h || !dst_host) return fail(POB_E_BAD_ARG, "pob_copy_witness: null argument");
    uint64_t *s = nullptr; int rc = resident_slot(h, index, &s); if (rc) return rc;
    if (first_signal + n_signals > h->P.n_signals) return fail(POB_E_RANGE, "pob_copy_witness: range exceeds the witness");
    if (cudaSetDevice(h->device) != cudaSuccess || cudaMemcpy(dst_host, s + 4 * first_signal, (size_t)n_signals * 32, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(POB_E_CUDA, "pob_copy_witness: cudaMemcpy failed");
    return POB_OK;
}

int pob_write_wtns(pob_handle *h, uint32_t index, const char *path) {
    if (!h || !path) return fail(POB_E_BAD_ARG, "pob_write_wtns: null argument");
    uint64_t *s = nullptr; int rc = resident_slot(h, index, &s); if (rc) return rc;
    FILE *f = fopen(path, "wb"); if (!f) return fail(POB_E_IO, std::string("cannot open ") + path);
    // iden3 binary witness format, version 2 (what the circom runtime's writeBinWitness emits)
    const uint64_t n = h->P.n_signals; uint32_t u32; uint64_t u64; Fr p = fr_p();
    bool ok = fwrite("wtns", 1, 4, f) == 4;
    u32 = 2; ok &= fwrite(&u32, 4, 1, f) == 1; u32 = 2; ok &= fwrite(&u32, 4, 1, f) == 1;
    u32 = 1; ok &= fwrite(&u32, 4, 1, f) == 1; u64 = 40; ok &= fwrite(&u64, 8, 1, f) == 1;
    u32 = 32; ok &= fwrite(&u32, 4, 1, f) == 1; ok &= fwrite(p.l, 4, 8, f) == 8; u32 = (uint32_t)n; ok &= fwrite(&u32, 4, 1, f) == 1;
    u32 = 2; ok &= fwrite(&u32, 4, 1, f) == 1; u64 = 32ull * n; ok &= fwrite(&u64, 8, 1, f) == 1;
    // double-buffered: the D2H copy of chunk i+1 (pinned, own stream) overlaps the fwrite of chunk i
    const uint64_t CH = 1ull << 21;     // 2 Mi entries = 64 MiB per hop
    void *buf[2] = {nullptr, nullptr}; cudaStream_t cs = nullptr; cudaEvent_t ev[2] = {nullptr, nullptr};
    bool cuda_ok = cudaSetDevice(h->device) == cudaSuccess && cudaMallocHost(&buf[0], CH * 32) == cudaSuccess && cudaMallocHost(&buf[1], CH * 32) == cudaSuccess &&
                   cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&ev[0]) == cudaSuccess && cudaEventCreate(&ev[1]) == cudaSuccess;
    const uint64_t nchunk = (n + CH - 1) / CH;
    auto issue = [&](uint64_t c) {
        const uint64_t off = c * CH, cnt = std::min(CH, n - off);
        return cudaMemcpyAsync(buf[c & 1], s + 4 * off, (size_t)cnt * 32, cudaMemcpyDeviceToHost, cs) == cudaSuccess && cudaEventRecord(ev[c & 1], cs) == cudaSuccess;
    };
    if (cuda_ok && nchunk) cuda_ok = issue(0);
    for (uint64_t c = 0; cuda_ok && ok && c < nchunk; c++) {
        cuda_ok = cudaEventSynchronize(ev[c & 1]) == cudaSuccess;
        if (cuda_ok && c + 1 < nchunk) cuda_ok = issue(c + 1);
        const uint64_t cnt = std::min(CH, n - c * CH);
        if (cuda_ok) ok &= fwrite(buf[c & 1], 32, (size_t)cnt, f) == cnt;
    }
    if (cs) cudaStreamSynchronize(cs);
    for (int i = 0; i < 2; i++) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    if (cs) cudaStreamDestroy(cs);
    if (!cuda_ok) { fclose(f); return fail(POB_E_CUDA, "pob_write_wtns: CUDA copy failed"); }
    ok &= fclose(f) == 0;
    return ok ? POB_OK : fail(POB_E_IO, std::string("short write to ") + path);
}

}  // extern "C"
