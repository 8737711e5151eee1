// L2 fetch granularity vs random gather / RED throughput over a 4 GiB table (roofline microbenchmark)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1;} } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
template <int MODE> __global__ void k(uint32_t* table, uint32_t mask, int iters, uint64_t seed, uint32_t* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint64_t h = mix64(seed + tid * 1315423911ull + it);
        uint32_t v[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = mix64(h + u); uint32_t a = (uint32_t)g & mask, b = (uint32_t)(g >> 32) & mask;
            if (MODE == 0) { v[2*u] = __ldg(table + a); v[2*u+1] = __ldg(table + b); }
            else if (MODE == 1) { v[2*u] = table[a]; v[2*u+1] = table[b]; }
            else if (MODE == 2) { v[2*u] = __ldcg(table + a); v[2*u+1] = __ldcg(table + b); }
            else { atomicAdd(table + a, 1u); atomicAdd(table + b, 1u); v[2*u] = v[2*u+1] = 0; }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += v[u];
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main() {
    uint32_t* table; CK(cudaMalloc(&table, 4ull << 30)); CK(cudaMemset(table, 0, 4ull << 30)); uint32_t* out; CK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 16, threads = 256, iters = 64; const double nops = (double)blocks * threads * iters * 16;
    for (int gran : {128, 64, 32}) {
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        for (int mode = 0; mode < 4; ++mode) {
            float best = 1e30f;
            for (int r = 0; r < 3; ++r) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, threads>>>(table, (1u << 30) - 1, iters, 7 + r, out);
                if (mode == 1) k<1><<<blocks, threads>>>(table, (1u << 30) - 1, iters, 7 + r, out);
                if (mode == 2) k<2><<<blocks, threads>>>(table, (1u << 30) - 1, iters, 7 + r, out);
                if (mode == 3) k<3><<<blocks, threads>>>(table, (1u << 30) - 1, iters, 7 + r, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const char* names[4] = {"gather_ldg_nc", "gather_ld_ca", "gather_ld_cg", "red_add"};
            printf("{\"bench\": \"l2fetch\", \"granularity_set\": %d, \"granularity_got\": %zu, \"op\": \"%s\", \"ms\": %.3f, \"Gops\": %.2f}\n", gran, got, names[mode], best, nops / best / 1e6);
        }
    }
    return 0;
}
