// expand_probe.cu -- isolates what limits the KeccakfRound path of k_expand: decode + L1 loads vs pure stores.
// Builds against the real layout compiler to get the real 64-signal group descriptors.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "compiler.h"
using namespace pob;

__device__ __forceinline__ void st256(uint64_t *p, uint64_t a) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(0ull), "l"(0ull), "l"(0ull));
}
__device__ __forceinline__ void decode(const uint2 d, uint32_t k, uint32_t &w, uint32_t &b) {
    const uint32_t tt = k & 63, mode = d.y >> 16;
    w = d.x & 0xffffu; b = tt;
    if (mode) { const uint32_t sidx = (mode - 1) * 64 + tt, g = sidx / 3, m = sidx - 3 * g; b = g; w = (m == 0) ? (d.x & 0xffffu) : (m == 1) ? (d.x >> 16) : (d.y & 0xffffu); }
}
static const uint32_t TILE = 8192, GROUPS = 1604, RSIG = 102656;
// every CTA handles one 8192-entry tile of "round" (blockIdx.x % 13)-th tile of a round block
__device__ __forceinline__ void tile_of(uint32_t bx, uint32_t &goff, uint32_t &n) { uint32_t t = bx % 13; goff = t * 128; n = t == 12 ? RSIG - 12 * TILE : TILE; }

// P0: pure stores
__global__ void __launch_bounds__(256) p0(uint64_t *w, const uint2 *, const uint64_t *) {
    uint64_t *W = w + (uint64_t)blockIdx.x * TILE * 4; uint32_t goff, n; tile_of(blockIdx.x, goff, n);
#pragma unroll 4
    for (uint32_t k = threadIdx.x; k < n; k += 256) st256(W + 4ull * k, k & 1);
}
// P2: the k_expand round path (global/L1 descriptor + word loads), UN entries in flight
template <int UN, int T, int MINB> __global__ void __launch_bounds__(T, MINB) p2(uint64_t *w, const uint2 *D0, const uint64_t *words) {
    uint64_t *W = w + (uint64_t)blockIdx.x * TILE * 4; uint32_t goff, n; tile_of(blockIdx.x, goff, n);
    const uint2 *D = D0 + goff; const uint64_t *Ub = words + (blockIdx.x / 13 % 64) * 263;
    for (uint32_t base = threadIdx.x; base < n; base += T * UN) {
        uint64_t word[UN]; uint32_t bit[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) { const uint32_t k = base + T * u; word[u] = 0; bit[u] = 0; if (k < n) { uint32_t ww, b; decode(__ldg(D + (k >> 6)), k, ww, b); word[u] = Ub[ww]; bit[u] = b; } }
#pragma unroll
        for (int u = 0; u < UN; u++) { const uint32_t k = base + T * u; if (k < n) st256(W + 4ull * k, (word[u] >> bit[u]) & 1ull); }
    }
}
// P3: descriptors + words staged in shared memory
template <int UN> __global__ void __launch_bounds__(256) p3(uint64_t *w, const uint2 *D0, const uint64_t *words) {
    __shared__ uint2 sD[128]; __shared__ uint64_t sW[264];
    uint64_t *W = w + (uint64_t)blockIdx.x * TILE * 4; uint32_t goff, n; tile_of(blockIdx.x, goff, n);
    const uint64_t *Ub = words + (blockIdx.x / 13 % 64) * 263;
    if (threadIdx.x < 128 && goff + threadIdx.x < GROUPS) sD[threadIdx.x] = D0[goff + threadIdx.x];
    for (uint32_t i = threadIdx.x; i < 263; i += 256) sW[i] = Ub[i];
    __syncthreads();
    for (uint32_t base = threadIdx.x; base < n; base += 256 * UN) {
#pragma unroll
        for (int u = 0; u < UN; u++) { const uint32_t k = base + 256 * u; if (k < n) { uint32_t ww, b; decode(sD[k >> 6], k, ww, b); st256(W + 4ull * k, (sW[ww] >> b) & 1ull); } }
    }
}
// P5: P3 + one expanded BYTE table in shared memory: pre-expand the tile's bits once (1 byte per entry), then stream
__global__ void __launch_bounds__(256) p5(uint64_t *w, const uint2 *D0, const uint64_t *words) {
    __shared__ uint2 sD[128]; __shared__ uint64_t sW[264]; __shared__ uint8_t sB[TILE];
    uint64_t *W = w + (uint64_t)blockIdx.x * TILE * 4; uint32_t goff, n; tile_of(blockIdx.x, goff, n);
    const uint64_t *Ub = words + (blockIdx.x / 13 % 64) * 263;
    if (threadIdx.x < 128 && goff + threadIdx.x < GROUPS) sD[threadIdx.x] = D0[goff + threadIdx.x];
    for (uint32_t i = threadIdx.x; i < 263; i += 256) sW[i] = Ub[i];
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n; k += 256) { uint32_t ww, b; decode(sD[k >> 6], k, ww, b); sB[k] = (uint8_t)((sW[ww] >> b) & 1ull); }
    __syncthreads();
#pragma unroll 8
    for (uint32_t k = threadIdx.x; k < n; k += 256) st256(W + 4ull * k, sB[k]);
}
// P6: decode into shared memory as full 32-byte entries, TMA bulk store of 32 KiB chunks (double buffered)
__global__ void __launch_bounds__(256) p6(uint64_t *w, const uint2 *D0, const uint64_t *words) {
    extern __shared__ __align__(128) uint64_t sm[];      // 2 x 32 KiB staging + tables
    uint2 *sD = reinterpret_cast<uint2 *>(sm + 2 * 4096); uint64_t *sW = sm + 2 * 4096 + 128;
    uint64_t *W = w + (uint64_t)blockIdx.x * TILE * 4; uint32_t goff, n; tile_of(blockIdx.x, goff, n);
    const uint64_t *Ub = words + (blockIdx.x / 13 % 64) * 263;
    if (threadIdx.x < 128 && goff + threadIdx.x < GROUPS) sD[threadIdx.x] = D0[goff + threadIdx.x];
    for (uint32_t i = threadIdx.x; i < 263; i += 256) sW[i] = Ub[i];
    __syncthreads();
    const uint32_t CH = 1024; int buf = 0;
    for (uint32_t base = 0; base < n; base += CH, buf ^= 1) {
        uint64_t *s = sm + buf * 4096;
        if (base >= 2 * CH) { if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); __syncthreads(); }
        const uint32_t cnt = min(CH, n - base);
        for (uint32_t j = threadIdx.x; j < cnt; j += 256) {
            uint32_t ww, b; decode(sD[(base + j) >> 6], base + j, ww, b);
            ulonglong2 lo = make_ulonglong2((sW[ww] >> b) & 1ull, 0), hi = make_ulonglong2(0, 0);
            *reinterpret_cast<ulonglong2 *>(s + 4 * j) = lo; *reinterpret_cast<ulonglong2 *>(s + 4 * j + 2) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(W + 4ull * base), "r"(saddr), "r"(cnt * 32) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> static float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    std::vector<Fr> ps(1); ps[0] = fr_from_u64(1);
    Program P = compile_circuit("KeccakBytes", ps, false);
    const uint32_t NT = 13 * 64 * 128;                      // tiles: 8192 round blocks worth
    uint64_t bytes = 0; for (uint32_t t = 0; t < NT; t++) bytes += 32ull * ((t % 13 == 12) ? RSIG - 12 * TILE : TILE);
    uint64_t *w; cudaMalloc(&w, (uint64_t)NT * TILE * 32);
    uint2 *D; cudaMalloc(&D, GROUPS * 8 + 1024); cudaMemcpy(D, P.round_desc.data(), GROUPS * 8, cudaMemcpyHostToDevice);
    std::vector<uint64_t> hw(64 * 263); for (size_t i = 0; i < hw.size(); i++) hw[i] = 0x9E3779B97F4A7C15ull * (i + 1);
    uint64_t *words; cudaMalloc(&words, hw.size() * 8); cudaMemcpy(words, hw.data(), hw.size() * 8, cudaMemcpyHostToDevice);
    auto report = [&](const char *name, float ms) { printf("%-56s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
    report("P0 pure stores T=256", timeit([&] { p0<<<NT, 256>>>(w, D, words); }, 5));
    report("P2 k_expand path UN=8 T=256 (48 regs)", timeit([&] { p2<8, 256, 5><<<NT, 256>>>(w, D, words); }, 5));
    report("P2 UN=4 T=256 minb 8", timeit([&] { p2<4, 256, 8><<<NT, 256>>>(w, D, words); }, 5));
    report("P2 UN=2 T=256 minb 8", timeit([&] { p2<2, 256, 8><<<NT, 256>>>(w, D, words); }, 5));
    report("P2 UN=4 T=512 minb 4", timeit([&] { p2<4, 512, 4><<<NT, 512>>>(w, D, words); }, 5));
    report("P2 UN=2 T=1024 minb 2", timeit([&] { p2<2, 1024, 2><<<NT, 1024>>>(w, D, words); }, 5));
    report("P3 smem tables UN=4", timeit([&] { p3<4><<<NT, 256>>>(w, D, words); }, 5));
    report("P3 smem tables UN=8", timeit([&] { p3<8><<<NT, 256>>>(w, D, words); }, 5));
    report("P5 smem tables + byte staging", timeit([&] { p5<<<NT, 256>>>(w, D, words); }, 5));
    cudaFuncSetAttribute(p6, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768 + 1024 + 264 * 8);
    report("P6 smem decode + TMA bulk store", timeit([&] { p6<<<NT, 256, 2 * 32768 + 1024 + 264 * 8>>>(w, D, words); }, 5));
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
