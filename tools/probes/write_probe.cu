// write_probe.cu -- what is the write-only HBM ceiling on this part, and which store shape reaches it?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_probe write_probe.cu ; ./write_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cuda/barrier>

__device__ __forceinline__ void st256(uint64_t *p, uint64_t a) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(0ull), "l"(0ull), "l"(0ull) : "memory");
}
__device__ __forceinline__ void st128(uint64_t *p, uint64_t a) {
    asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(0ull) : "memory");
}
// A: one CTA per contiguous tile of `tile` 32-byte entries, thread k writes entries k, k+T, ...
template <int T> __global__ void __launch_bounds__(T) k_tile256(uint64_t *w, uint32_t tile, uint64_t v) {
    uint64_t *W = w + (uint64_t)blockIdx.x * tile * 4;
#pragma unroll 4
    for (uint32_t k = threadIdx.x; k < tile; k += T) st256(W + 4ull * k, v + k);
}
// B: 128-bit stores, two per entry
template <int T> __global__ void __launch_bounds__(T) k_tile128(uint64_t *w, uint32_t tile, uint64_t v) {
    uint64_t *W = w + (uint64_t)blockIdx.x * tile * 4;
#pragma unroll 4
    for (uint32_t k = threadIdx.x; k < 2 * tile; k += T) st128(W + 2ull * k, (k & 1) ? 0 : v + k);
}
// C: persistent grid, grid-stride over tiles
template <int T> __global__ void __launch_bounds__(T) k_persist(uint64_t *w, uint32_t tile, uint32_t ntiles, uint64_t v) {
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint64_t *W = w + (uint64_t)t * tile * 4;
#pragma unroll 4
        for (uint32_t k = threadIdx.x; k < tile; k += T) st256(W + 4ull * k, v + k);
    }
}
// D: stage a 32 KiB tile in shared memory, one thread issues a TMA bulk store (cp.async.bulk.global.shared::cta)
__global__ void __launch_bounds__(256) k_bulk(uint64_t *w, uint32_t tile, uint64_t v) {
    extern __shared__ __align__(128) uint64_t sm[];      // 2 x 32 KiB
    const uint32_t CH = 1024;                             // entries per chunk (32 KiB)
    uint64_t *W = w + (uint64_t)blockIdx.x * tile * 4;
    int buf = 0;
    for (uint32_t base = 0; base < tile; base += CH, buf ^= 1) {
        uint64_t *s = sm + buf * CH * 4;
        if (base >= 2 * CH) { if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); __syncthreads(); }
        for (uint32_t k = threadIdx.x; k < CH; k += 256) { s[4 * k] = v + base + k; s[4 * k + 1] = 0; s[4 * k + 2] = 0; s[4 * k + 3] = 0; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(W + 4ull * base), "r"(saddr), "r"(CH * 32) : "memory");
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
    const uint64_t BYTES = 32ull << 30; uint64_t *w; if (cudaMalloc(&w, BYTES) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    auto report = [&](const char *name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, BYTES / (ms * 1e-3) / 1e9); fflush(stdout); };
    for (uint32_t tile : {2048u, 8192u, 32768u}) {
        uint32_t nt = (uint32_t)(BYTES / 32 / tile); char nm[96];
        snprintf(nm, 96, "tile256 T=256 tile=%u", tile); report(nm, timeit([&] { k_tile256<256><<<nt, 256>>>(w, tile, 1); }, 5));
        snprintf(nm, 96, "tile256 T=512 tile=%u", tile); report(nm, timeit([&] { k_tile256<512><<<nt, 512>>>(w, tile, 1); }, 5));
        snprintf(nm, 96, "tile256 T=1024 tile=%u", tile); report(nm, timeit([&] { k_tile256<1024><<<nt, 1024>>>(w, tile, 1); }, 5));
        snprintf(nm, 96, "tile128 T=256 tile=%u", tile); report(nm, timeit([&] { k_tile128<256><<<nt, 256>>>(w, tile, 1); }, 5));
    }
    for (int mult : {2, 4, 8}) {
        uint32_t tile = 8192, nt = (uint32_t)(BYTES / 32 / tile); char nm[96];
        snprintf(nm, 96, "persistent 148x%d CTAs T=256 tile=8192", mult); report(nm, timeit([&] { k_persist<256><<<148 * mult, 256>>>(w, tile, nt, 1); }, 5));
    }
    { uint32_t tile = 8192, nt = (uint32_t)(BYTES / 32 / tile);
      cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      report("TMA bulk store 32 KiB chunks, tile=8192", timeit([&] { k_bulk<<<nt, 256, 65536>>>(w, tile, 1); }, 5)); }
    { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e)); }
    report("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(w, 0, BYTES); }, 3));
    return 0;
}
