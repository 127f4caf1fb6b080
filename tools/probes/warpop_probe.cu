// warpop_probe.cu -- latency of k_eval's building blocks on an otherwise idle SM (one warp, clock64 inside the kernel):
//   one Keccak absorb (24 rounds, 238 words stored per round), one Montgomery product in a dependent chain (PTX form and the
//   portable form, -DPOB_PORTABLE_MONT), one iteration of the deferred-inverse state machine.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I proof-of-burn_b200/csrc -I include -o warpop_probe tools/probes/warpop_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "compiler.h"
#include "kernels.cuh"

__global__ void k_absorb(uint64_t *W, AbsorbOp op, long long *cyc) {
    const long long t0 = clock64();
    absorb_warp(W, op);
    __syncwarp();
    if (threadIdx.x == 0) cyc[0] = clock64() - t0;
}
__global__ void k_mont_chain(Fr *v, int n, long long *cyc) {
    Fr a = v[threadIdx.x], b = v[32 + threadIdx.x];
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) a = fr_mont(a, b);
    const long long t1 = clock64();
    v[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / n;
}
__global__ void k_inv_steps(InvChain *c, int n, long long *cyc) {
    InvChain x = c[threadIdx.x];
    const long long t0 = clock64();
    inv_chain_steps(x, (uint32_t)n);
    const long long t1 = clock64();
    c[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / n;
}

int main() {
    uint64_t *W; cudaMalloc(&W, (size_t)(64 + ABSORB_WORDS + 64) * 8); cudaMemset(W, 0x3c, (size_t)(64 + ABSORB_WORDS + 64) * 8);
    long long *cyc; cudaMallocManaged(&cyc, 64);
    AbsorbOp op{NONE_IDX, 0, 64, 0};
    for (int r = 0; r < 3; r++) { k_absorb<<<1, 32>>>(W, op, cyc); cudaDeviceSynchronize(); }
    printf("absorb (one warp, idle SM): %lld cycles = %lld per round\n", cyc[0], cyc[0] / 24);
    Fr *v; cudaMallocManaged(&v, 64 * sizeof(Fr));
    for (int i = 0; i < 64; i++) for (int k = 0; k < 8; k++) v[i].l[k] = (k == 7) ? 0x0fffffffu : 0x9e3779b9u * (i * 8 + k + 1);
    for (int r = 0; r < 3; r++) { k_mont_chain<<<1, 32>>>(v, 512, cyc); cudaDeviceSynchronize(); }
    printf("fr_mont, dependent chain (one warp): %lld cycles per product\n", cyc[1]);
    InvChain *c; cudaMallocManaged(&c, 32 * sizeof(InvChain));
    for (int i = 0; i < 32; i++) inv_chain_init(c[i], v[i]);
    for (int r = 0; r < 1; r++) { k_inv_steps<<<1, 32>>>(c, 200, cyc); cudaDeviceSynchronize(); }
    printf("deferred-inverse iteration (one warp): %lld cycles\n", cyc[2]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
