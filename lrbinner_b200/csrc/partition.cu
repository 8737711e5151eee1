// partition.cu — L2-resident 15-mer table passes (count and search) via one key-partition of the windows.
//
// Why (measured on this pool's B200, profiles/r01_ubench_roofline.jsonl): uniform-random RED.ADD.U32 runs at
// 20 G/s over a 4 GiB table (every update misses L2: 32 B sector in, 32 B out) but at ~190 G/s when the
// touched slice is <= 64 MiB; random 4 B gathers: 38-48 G/s vs 288 G/s.  The direct kernels
// (kernels.cu: k_count15 / k_search15) sit exactly on the DRAM-random numbers.  Here the valid windows are
// first partitioned by the high bits of their bit-15-clear key into buckets whose table slice (2^25 keys =
// 128 MiB of addresses, 64 MiB touched because bit 15 is clear) fits the 126 MB L2; then, bucket by bucket,
// the keys are streamed back (coalesced) and applied to the resident slice:
//
//   k_bucket_hist   one scan: windows per bucket                                   (sizes the regions exactly)
//   k_partition     one scan: (key[, read id]) -> bucket regions; per 8192-slot CTA chunk the entries are
//                   ranked with shared-memory atomics, staged in shared memory and copied out as contiguous
//                   runs, so global writes are coalesced
//   k_count_keys    per bucket: RED.ADD.U32 table[key]                              (L2-resident atomics)
//   k_search_keys   per bucket: count = table[key] (L2 hit, right after the bucket was counted), bucket rule,
//                   run-length aggregation of equal (read, bin) neighbours, RED into hist[read][bin] / sums
//
// Results are bit-identical to the direct kernels: the same windows, the same keys, integer sums.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/lrbinner_b200.h"
#include "common.h"
#include "lane_core.cuh"

using namespace lrb;

namespace {

constexpr int kPartThreads = 256;          // one 32-slot block per thread -> 8192 slots per CTA chunk
constexpr int kChunkSlots = kPartThreads * 32;
constexpr int kMaxBuckets = 64;

__global__ void __launch_bounds__(256) k_fill_blk_read(lrb_reads_view R, uint32_t* __restrict__ blk_read) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R.n_reads) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t b1 = R.read_blk[r + 1];
    for (uint32_t b = R.read_blk[r] + lane; b < b1; b += 32) blk_read[b] = (uint32_t)r;
}

struct BlockWindows {
    uint32_t m, pw, w0, w1;
};

__device__ __forceinline__ BlockWindows load_block(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid,
                                                   uint64_t gb) {
    BlockWindows b;
    const uint32_t v = __ldg(valid + gb);
    const uint32_t pv = gb ? __ldg(valid + gb - 1) : 0u;
    b.m = window15_mask(pv, v);
    b.pw = b.w0 = b.w1 = 0;
    if (b.m) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(codes) + gb);
        b.w0 = w.x;
        b.w1 = w.y;
        b.pw = gb ? __ldg(codes + 2 * gb - 1) : 0u;
    }
    return b;
}

// windows per bucket (bucket = key >> shift, numbered from bucket0 = key_lo >> shift)
__global__ void __launch_bounds__(256)
k_bucket_hist(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid, uint64_t blk_lo, uint64_t blk_hi,
              uint32_t key_lo, uint32_t key_hi, int shift, unsigned long long* __restrict__ counts) {
    __shared__ uint32_t s_cnt[kMaxBuckets];
    if (threadIdx.x < kMaxBuckets) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t bucket0 = key_lo >> shift;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t gb = blk_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < blk_hi; gb += stride) {
        const BlockWindows b = load_block(codes, valid, gb);
        if (!b.m) continue;
        canon15_block(b.pw, b.w0, b.w1, b.m, [&](uint32_t key) {
            if (key >= key_lo && key < key_hi) atomicAdd(&s_cnt[(key >> shift) - bucket0], 1u);
        });
    }
    __syncthreads();
    if (threadIdx.x < kMaxBuckets && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// (key[, read]) of every valid window -> its bucket's region, coalesced through a shared-memory stage
template <bool WITH_RID>
__global__ void __launch_bounds__(kPartThreads, 2)
k_partition(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ blk_read,
            uint64_t blk_lo, uint64_t blk_hi, uint32_t key_lo, uint32_t key_hi, int shift, int nb,
            const unsigned long long* __restrict__ offsets, unsigned long long* __restrict__ cursor,
            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ rids_out) {
    extern __shared__ uint32_t s_stage[];  // keys[kChunkSlots] (+ rids[kChunkSlots])
    __shared__ uint32_t s_cnt[kMaxBuckets], s_base[kMaxBuckets + 1];
    __shared__ unsigned long long s_gbase[kMaxBuckets];
    uint32_t* stage_key = s_stage;
    uint32_t* stage_rid = s_stage + kChunkSlots;
    const uint32_t bucket0 = key_lo >> shift;
    const int tid = threadIdx.x;
    const uint64_t n_chunks = (blk_hi - blk_lo + kPartThreads - 1) / kPartThreads;

    for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        if (tid < kMaxBuckets) s_cnt[tid] = 0;
        __syncthreads();
        const uint64_t gb = blk_lo + chunk * kPartThreads + tid;
        uint32_t key[32];
        uint32_t posw[16];  // rank inside the CTA's bucket run, two 16-bit values per word
        uint32_t m = 0, rid = 0;
        if (gb < blk_hi) {
            const BlockWindows b = load_block(codes, valid, gb);
            m = b.m;
            if (m) {
                if (WITH_RID) rid = __ldg(blk_read + gb);
                const uint32_t r0 = rc16(b.w1), r1 = rc16(b.w0), r2 = rc16(b.pw);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if ((j & 1) == 0) posw[j >> 1] = 0;
                    key[j] = 0;
                    if ((m >> j) & 1u) {
                        const uint32_t kk = canonical15(kmer_ending_at<15>(b.pw, b.w0, b.w1, j), rc15_ending_at(r0, r1, r2, j));
                        if (kk >= key_lo && kk < key_hi) {
                            key[j] = kk;
                            const uint32_t p = atomicAdd(&s_cnt[(kk >> shift) - bucket0], 1u);
                            posw[j >> 1] |= p << (16 * (j & 1));
                        } else {
                            m &= ~(1u << j);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid < 32) {  // exclusive scan of the bucket counts (nb <= 64: two per lane) + global reservation
            const uint32_t c0 = (tid < nb) ? s_cnt[tid] : 0u;
            const uint32_t c1 = (tid + 32 < nb) ? s_cnt[tid + 32] : 0u;
            uint32_t x0 = c0, x1 = c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y0 = __shfl_up_sync(0xFFFFFFFFu, x0, d), y1 = __shfl_up_sync(0xFFFFFFFFu, x1, d);
                if (tid >= d) { x0 += y0; x1 += y1; }
            }
            const uint32_t tot0 = __shfl_sync(0xFFFFFFFFu, x0, 31);
            s_base[tid] = x0 - c0;
            s_base[tid + 32] = tot0 + x1 - c1;
            if (tid == 31) s_base[kMaxBuckets] = tot0 + x1;
            if (tid < nb && c0) s_gbase[tid] = offsets[tid] + atomicAdd(&cursor[tid], (unsigned long long)c0);
            if (tid + 32 < nb && c1) s_gbase[tid + 32] = offsets[tid + 32] + atomicAdd(&cursor[tid + 32], (unsigned long long)c1);
        }
        __syncthreads();
        if (m) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if ((m >> j) & 1u) {
                    const uint32_t idx = s_base[(key[j] >> shift) - bucket0] + ((posw[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                    stage_key[idx] = key[j];
                    if (WITH_RID) stage_rid[idx] = rid;
                }
            }
        }
        __syncthreads();
        const uint32_t total = s_base[kMaxBuckets];
        for (uint32_t i = tid; i < total; i += kPartThreads) {
            const uint32_t kk = stage_key[i];
            const uint32_t b = (kk >> shift) - bucket0;
            const unsigned long long dst = s_gbase[b] + (i - s_base[b]);
            __stcs(keys_out + dst, kk);
            if (WITH_RID) __stcs(rids_out + dst, stage_rid[i]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_count_keys(const uint32_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint32_t k0 = __ldcs(keys + i), k1 = __ldcs(keys + i + stride), k2 = __ldcs(keys + i + 2 * stride),
                       k3 = __ldcs(keys + i + 3 * stride);
        atomicAdd(table + k0, 1u);
        atomicAdd(table + k1, 1u);
        atomicAdd(table + k2, 1u);
        atomicAdd(table + k3, 1u);
    }
    for (; i < n; i += stride) atomicAdd(table + __ldcs(keys + i), 1u);
}

// entries of one bucket: hist[read][bin(table[key])] += 1, sums[read] += 1, aggregated over runs of equal neighbours
__global__ void __launch_bounds__(256)
k_search_keys(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rids, uint64_t n,
              const uint32_t* __restrict__ table, uint32_t S32, uint64_t magic, uint32_t B, uint32_t* __restrict__ hist,
              uint32_t* __restrict__ sums) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += stride) {
        const bool act = i < n;
        const uint32_t key = act ? __ldcs(keys + i) : 0u;
        const uint32_t rid = act ? __ldcs(rids + i) : 0xFFFFFFFFu;
        const uint32_t cnt = act ? table[key] : 0u;  // written by k_count_keys just before: an L2 hit
        const uint32_t bin = coverage_bin(cnt, S32, magic, B);
        const unsigned long long tag = act ? (((unsigned long long)rid << 12) | bin) : ~0ull;
        const unsigned long long ptag = __shfl_up_sync(0xFFFFFFFFu, tag, 1);
        const uint32_t prid = __shfl_up_sync(0xFFFFFFFFu, rid, 1);
        const bool head = (lane == 0) || (tag != ptag);
        const bool head_r = (lane == 0) || (rid != prid);
        const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
        const uint32_t heads_r = __ballot_sync(0xFFFFFFFFu, head_r);
        const uint32_t above = ~((2u << lane) - 1u);  // lanes above this one (lane 31: none)
        if (act && head) {
            const uint32_t nx = heads & above;
            const uint32_t run = (nx ? (uint32_t)__ffs(nx) - 1u : 32u) - lane;
            atomicAdd(hist + (size_t)rid * B + bin, run);
        }
        if (act && head_r) {
            const uint32_t nx = heads_r & above;
            const uint32_t run = (nx ? (uint32_t)__ffs(nx) - 1u : 32u) - lane;
            atomicAdd(sums + rid, run);
        }
    }
}

int sms() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

}  // namespace

extern "C" int lrb_dev_fill_blk_read(const lrb_reads_view* dev, uint32_t* blk_read, void* stream) {
    if (!dev || !blk_read) return lrb_set_error(LRB_EINVAL, "lrb_dev_fill_blk_read: null argument");
    if (!dev->n_reads) return LRB_OK;
    const uint64_t threads = dev->n_reads * 32;
    k_fill_blk_read<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*dev, blk_read);
    LRB_CUDA(cudaGetLastError());
    return LRB_OK;
}

// ---- host side: build the partition once, then count and/or search bucket by bucket ---------------------------
extern "C" int lrb_dev_partition_build(const lrb_reads_view* dev, const uint32_t* blk_read, int with_rids, uint64_t blk_lo,
                                       uint64_t blk_hi, uint32_t key_lo, uint32_t key_hi, int log2_bucket_keys,
                                       lrb_partition* part, void* stream) {
    if (!dev || !part || !part->keys || !part->small) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_build: null argument");
    if (with_rids && (!part->rids || !blk_read)) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_build: read ids need part->rids and blk_read");
    if (key_hi > kTableEntries) key_hi = kTableEntries;
    if (blk_hi > dev->n_blocks) blk_hi = dev->n_blocks;
    const int shift = log2_bucket_keys;
    if (shift < 20 || shift > 30) return lrb_set_error(LRB_EINVAL, "log2_bucket_keys must be in [20, 30]");
    const uint32_t bsz = 1u << shift;
    if (key_lo >= key_hi || (key_lo & (bsz - 1)) || (key_hi & (bsz - 1)))
        return lrb_set_error(LRB_EINVAL, "key range must be non-empty and aligned to the bucket size 2^%d", shift);
    const int nb = (int)((key_hi - key_lo) >> shift);
    if (nb < 1 || nb > kMaxBuckets) return lrb_set_error(LRB_EINVAL, "key range spans %d buckets (max %d): raise log2_bucket_keys", nb, kMaxBuckets);
    part->n_buckets = nb;
    part->shift = shift;
    part->key_lo = key_lo;
    part->has_rids = with_rids ? 1 : 0;
    for (int b = 0; b <= kMaxBuckets; ++b) part->offset[b] = 0;
    for (int b = 0; b < kMaxBuckets; ++b) part->count[b] = 0;
    if (blk_lo >= blk_hi) return LRB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_counts = part->small;                    // [64]
    unsigned long long* d_offsets = part->small + kMaxBuckets;     // [64]
    unsigned long long* d_cursor = part->small + 2 * kMaxBuckets;  // [64]
    LRB_CUDA(cudaMemsetAsync(part->small, 0, sizeof(unsigned long long) * 3 * kMaxBuckets, st));
    const uint64_t nblk = blk_hi - blk_lo;
    const int nsm = sms();
    {
        const uint64_t want = (nblk + 255) / 256;
        const unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)nsm * 8);
        k_bucket_hist<<<grid, 256, 0, st>>>(dev->codes, dev->valid, blk_lo, blk_hi, key_lo, key_hi, shift, d_counts);
        LRB_CUDA(cudaGetLastError());
    }
    LRB_CUDA(cudaMemcpyAsync(part->count, d_counts, sizeof(unsigned long long) * kMaxBuckets, cudaMemcpyDeviceToHost, st));
    LRB_CUDA(cudaStreamSynchronize(st));  // the one host round trip: region sizes
    for (int b = 0; b < kMaxBuckets; ++b) part->offset[b + 1] = part->offset[b] + (b < nb ? part->count[b] : 0ull);
    const unsigned long long total = part->offset[nb];
    if (total > part->capacity)
        return lrb_set_error(LRB_ENOMEM, "partition workspace too small: %llu entries needed, %llu available", total, (unsigned long long)part->capacity);
    LRB_CUDA(cudaMemcpyAsync(d_offsets, part->offset, sizeof(unsigned long long) * kMaxBuckets, cudaMemcpyHostToDevice, st));
    const uint64_t n_chunks = (nblk + kPartThreads - 1) / kPartThreads;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_chunks, (uint64_t)nsm * 16);
    if (with_rids) {
        const size_t smem = sizeof(uint32_t) * 2 * kChunkSlots;
        LRB_CUDA(cudaFuncSetAttribute(k_partition<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_partition<true><<<grid, kPartThreads, smem, st>>>(dev->codes, dev->valid, blk_read, blk_lo, blk_hi, key_lo, key_hi, shift, nb,
                                                              d_offsets, d_cursor, part->keys, part->rids);
    } else {
        const size_t smem = sizeof(uint32_t) * kChunkSlots;
        k_partition<false><<<grid, kPartThreads, smem, st>>>(dev->codes, dev->valid, blk_read, blk_lo, blk_hi, key_lo, key_hi, shift, nb,
                                                               d_offsets, d_cursor, part->keys, part->rids);
    }
    LRB_CUDA(cudaGetLastError());
    return LRB_OK;
}

// mode bit 0: count (table[key] += 1), bit 1: search (hist/sums through table[key]); both = per bucket count then search
extern "C" int lrb_dev_partition_apply(const lrb_partition* part, int mode, uint32_t* table, long bin_size, int bins,
                                       uint32_t* hist, uint32_t* sums, void* stream) {
    if (!part || !table) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_apply: null argument");
    const bool do_count = mode & 1, do_search = mode & 2;
    if (do_search) {
        if (!hist || !sums) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_apply: search needs hist and sums");
        if (!part->has_rids) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_apply: partition was built without read ids");
        if (bin_size <= 0) return lrb_set_error(LRB_EINVAL, "bin_size must be >= 1 (the reference divides by it)");
        if (bins <= 0 || bins > LRB_MAX_BINS) return lrb_set_error(LRB_EINVAL, "bins must be in [1, %d]", LRB_MAX_BINS);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t S32 = bin_size > 0xFFFFFFFFl ? 0xFFFFFFFFu : (uint32_t)(bin_size > 0 ? bin_size : 1);
    const uint64_t magic = coverage_magic(S32);
    const int nsm = sms();
    for (int b = 0; b < part->n_buckets; ++b) {
        const uint64_t n = part->count[b];
        if (!n) continue;
        const uint32_t* kb = part->keys + part->offset[b];
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 1023) / 1024, (uint64_t)nsm * 8);
        if (do_count) k_count_keys<<<grid, 256, 0, st>>>(kb, n, table);
        if (do_search) k_search_keys<<<grid, 256, 0, st>>>(kb, part->rids + part->offset[b], n, table, S32, magic, (uint32_t)bins, hist, sums);
    }
    LRB_CUDA(cudaGetLastError());
    return LRB_OK;
}
