// Roofline microbenchmarks for the 15-mer table passes (not product code).
// Measures, on the box it runs on, the denominators DESIGN.md / bench.py quote:
//   * uniform-random red.global.add.u32 throughput vs table footprint (L2-resident .. 4 GiB)
//   * uniform-random 4-byte gather throughput vs footprint
//   * shared-memory atomic throughput (spread / contended)
//   * pinned H2D / D2H bandwidth, 4 GiB memset
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/ubench tools/ubench_roofline.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// each thread issues `per_thread` random REDs, 16 per hash pair (unrolled) into table[0..mask]
template <bool HALF_BLOCKS>
__global__ void k_red(uint32_t* __restrict__ table, uint32_t mask, int iters, uint64_t seed) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint64_t h = mix64(seed + tid * 1315423911ull + it);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = mix64(h + u);
            uint32_t a = (uint32_t)g & mask, b = (uint32_t)(g >> 32) & mask;
            if (HALF_BLOCKS) { a &= ~0x8000u; b &= ~0x8000u; }  // only bit15==0 entries (canonical half)
            atomicAdd(table + a, 1u);
            atomicAdd(table + b, 1u);
        }
    }
}

__global__ void k_gather(const uint32_t* __restrict__ table, uint32_t mask, int iters, uint64_t seed, uint32_t* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint64_t h = mix64(seed + tid * 1315423911ull + it);
        uint32_t v[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = mix64(h + u);
            v[2 * u] = __ldg(table + ((uint32_t)g & mask));
            v[2 * u + 1] = __ldg(table + ((uint32_t)(g >> 32) & mask));
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += v[u];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// shared-memory atomics: nbins bins per CTA, each thread does iters*16 atomics
__global__ void k_smem_atomic(int nbins_mask, int iters, uint64_t seed, uint32_t* out) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i <= nbins_mask; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint64_t h = mix64(seed + tid * 1315423911ull + it);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = mix64(h + u);
            atomicAdd(&sh[(uint32_t)g & nbins_mask], 1u);
            atomicAdd(&sh[(uint32_t)(g >> 32) & nbins_mask], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh[0] == 0xFFFFFFFFu) out[0] = 1;
}

// private (per-lane) byte histogram increments: non-atomic LDS/ADD/STS, conflict-free layout
__global__ void k_smem_private(int nbins_mask, int iters, uint64_t seed, uint32_t* out) {
    extern __shared__ uint8_t sh8[];
    // layout: word (bin>>2)*blockDim + tid, byte bin&3
    int nwords = ((nbins_mask + 1) >> 2) * blockDim.x;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) ((uint32_t*)sh8)[i] = 0;
    __syncthreads();
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint64_t h = mix64(seed + tid * 1315423911ull + it);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t g = mix64(h + u);
            uint32_t a = (uint32_t)g & nbins_mask, b = (uint32_t)(g >> 32) & nbins_mask;
            sh8[((a >> 2) * blockDim.x + threadIdx.x) * 4 + (a & 3)]++;
            sh8[((b >> 2) * blockDim.x + threadIdx.x) * 4 + (b & 3)]++;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh8[0] == 0xFF && sh8[1] == 0xFE) out[0] = 1;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d}\n", p.name, p.multiProcessorCount, p.l2CacheSize);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint32_t* table; size_t tbytes = 4ull << 30;
    CK(cudaMalloc(&table, tbytes));
    uint32_t* out; CK(cudaMalloc(&out, 64));
    // memset 4 GiB
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(table, 0, tbytes)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf("{\"bench\": \"memset_4GiB\", \"ms\": %.3f, \"GBps\": %.1f}\n", time_ms(e0, e1), tbytes / time_ms(e0, e1) / 1e6);
    }
    const int threads = 256;
    const int blocks = p.multiProcessorCount * 8;
    // RED sweep
    for (int half = 0; half < 2; ++half)
    for (int lg = 22; lg <= 30; ++lg) {   // entries: 4M (16 MB) .. 1G (4 GiB)
        uint32_t mask = (1u << lg) - 1;
        int iters = 64;
        double nops = (double)blocks * threads * iters * 16;
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0));
            if (half) k_red<true><<<blocks, threads>>>(table, mask, iters, 1234 + r);
            else k_red<false><<<blocks, threads>>>(table, mask, iters, 1234 + r);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms = time_ms(e0, e1); if (ms < best) best = ms;
        }
        printf("{\"bench\": \"red_u32_random\", \"half_blocks\": %d, \"footprint_MiB\": %.0f, \"ops\": %.3g, \"ms\": %.3f, \"Gops\": %.2f}\n",
               half, (double)(1ull << lg) * 4 / (1 << 20) / (half ? 2 : 1), nops, best, nops / best / 1e6);
        fflush(stdout);
    }
    // bigger RED run at 4 GiB to get steady state
    {
        uint32_t mask = (1u << 30) - 1; int iters = 512;
        double nops = (double)blocks * threads * iters * 16;
        CK(cudaEventRecord(e0)); k_red<false><<<blocks, threads>>>(table, mask, iters, 99); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("{\"bench\": \"red_u32_random_long\", \"footprint_MiB\": 4096, \"ops\": %.3g, \"ms\": %.3f, \"Gops\": %.2f}\n", nops, ms, nops / ms / 1e6);
        CK(cudaEventRecord(e0)); k_red<true><<<blocks, threads>>>(table, mask, iters, 99); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        ms = time_ms(e0, e1);
        printf("{\"bench\": \"red_u32_random_long\", \"half_blocks\": 1, \"footprint_MiB\": 2048, \"ops\": %.3g, \"ms\": %.3f, \"Gops\": %.2f}\n", nops, ms, nops / ms / 1e6);
    }
    // gather sweep
    for (int lg = 22; lg <= 30; ++lg) {
        uint32_t mask = (1u << lg) - 1;
        int iters = 64;
        double nops = (double)blocks * threads * iters * 16;
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0));
            k_gather<<<blocks, threads>>>(table, mask, iters, 4321 + r, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms = time_ms(e0, e1); if (ms < best) best = ms;
        }
        printf("{\"bench\": \"gather_u32_random\", \"footprint_MiB\": %.0f, \"ops\": %.3g, \"ms\": %.3f, \"Gops\": %.2f, \"sectorGBps\": %.1f}\n",
               (double)(1ull << lg) * 4 / (1 << 20), nops, best, nops / best / 1e6, nops * 32 / best / 1e6);
        fflush(stdout);
    }
    // occupancy variants for gather at 4 GiB: more blocks
    for (int mult = 4; mult <= 32; mult *= 2) {
        uint32_t mask = (1u << 30) - 1; int iters = 64; int b2 = p.multiProcessorCount * mult;
        double nops = (double)b2 * threads * iters * 16;
        CK(cudaEventRecord(e0)); k_gather<<<b2, threads>>>(table, mask, iters, 777, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("{\"bench\": \"gather_u32_random_4GiB\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"Gops\": %.2f}\n", mult, ms, nops / ms / 1e6);
        CK(cudaEventRecord(e0)); k_red<false><<<b2, threads>>>(table, mask, iters, 778); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        ms = time_ms(e0, e1);
        printf("{\"bench\": \"red_u32_random_4GiB\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"Gops\": %.2f}\n", mult, ms, nops / ms / 1e6);
    }
    // smem atomics
    for (int lg = 5; lg <= 15; lg += 2) {
        int mask = (1 << lg) - 1; int iters = 256; size_t sh = (size_t)(mask + 1) * 4;
        CK(cudaFuncSetAttribute(k_smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int b2 = p.multiProcessorCount * (sh > 100 * 1024 ? 1 : 2);
        double nops = (double)b2 * threads * iters * 16;
        CK(cudaEventRecord(e0)); k_smem_atomic<<<b2, threads, sh>>>(mask, iters, 5, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms = time_ms(e0, e1);
        printf("{\"bench\": \"smem_atomic_random\", \"bins\": %d, \"ctas\": %d, \"ms\": %.3f, \"Gops\": %.2f}\n", mask + 1, b2, ms, nops / ms / 1e6);
    }
    for (int lg = 5; lg <= 9; lg += 2) {
        int mask = (1 << lg) - 1; int iters = 8;  // <=255 increments per byte: 8*16=128 ok
        size_t sh = (size_t)((mask + 1) / 4) * threads * 4;
        CK(cudaFuncSetAttribute(k_smem_private, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (sh > 200 * 1024) continue;
        int b2 = p.multiProcessorCount * 64;
        double nops = (double)b2 * threads * iters * 16;
        CK(cudaEventRecord(e0)); k_smem_private<<<b2, threads, sh>>>(mask, iters, 5, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms = time_ms(e0, e1);
        printf("{\"bench\": \"smem_private_u8\", \"bins\": %d, \"smem\": %zu, \"ms\": %.3f, \"Gops\": %.2f}\n", mask + 1, sh, ms, nops / ms / 1e6);
    }
    // PCIe
    {
        size_t n = 1ull << 30; void* h; CK(cudaMallocHost(&h, n));
        for (int r = 0; r < 2; ++r) {
            CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(table, h, n, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            printf("{\"bench\": \"h2d_pinned_1GiB\", \"ms\": %.3f, \"GBps\": %.1f}\n", time_ms(e0, e1), n / time_ms(e0, e1) / 1e6);
            CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(h, table, n, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            printf("{\"bench\": \"d2h_pinned_1GiB\", \"ms\": %.3f, \"GBps\": %.1f}\n", time_ms(e0, e1), n / time_ms(e0, e1) / 1e6);
        }
        CK(cudaFreeHost(h));
    }
    return 0;
}
