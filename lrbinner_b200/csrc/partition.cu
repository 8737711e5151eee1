// partition.cu — L2-resident 15-mer table passes (count and search) via one key-partition of the windows.
//
// Why (measured on this pool's B200, profiles/r01_ubench_roofline.jsonl): uniform-random RED.ADD.U32 runs at
// 20 G/s over a 4 GiB table (every update misses L2: a sector pair in, a sector out) but at ~190 G/s when
// the touched slice is <= 64 MiB; random 4 B gathers: 38-48 G/s vs 288 G/s.  The direct kernels
// (kernels.cu: k_count15 / k_search15) sit exactly on the DRAM-random numbers (profiles/r01_bench_n1_v0_*).
// Here the valid windows are first partitioned by the high bits of their bit-15-clear key into buckets whose
// table slice (2^24 keys = 64 MiB of addresses, 32 MiB touched because bit 15 is clear) fits the 126 MB L2
// together with the histogram rows; then, bucket by bucket, the lists are streamed back (coalesced) and
// applied to the resident slice:
//
//   k_bucket_hist   one scan of a chunk of the stream: windows per bucket           (sizes the regions exactly)
//   k_chunk_scan    device-side exclusive scan -> region offsets of the chunk       (no host round trip)
//   k_partition     second scan: (key[, read id]) -> bucket regions; per 8192-slot CTA step the entries are
//                   ranked with shared-memory atomics, staged in shared memory and copied out bucket by
//                   bucket as contiguous runs, so global writes are coalesced
//   k_count_keys    per bucket: RED.ADD.U32 table[key]                              (L2-resident atomics)
//   k_search_keys   per bucket: count = table[key] (an L2 hit), bucket rule, run-length aggregation of equal
//                   (read, bin) neighbours, RED into hist[read][bin] / sums[read]
//
// The stream can be added in several chunks (lrb_dev_partition_add), each with its own regions, so the
// partition of chunk i overlaps the host-to-device copy of chunk i+1; everything is asynchronous on the
// caller's stream.  Results are bit-identical to the direct kernels: same windows, same keys, integer sums.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/lrbinner_b200.h"
#include "common.h"
#include "lane_core.cuh"

using namespace lrb;

namespace {

constexpr int kPartThreads = 256;          // one 32-slot block per thread -> 8192 slots per CTA step
constexpr int kChunkSlots = kPartThreads * 32;
constexpr int kMaxBuckets = LRB_PART_MAX_BUCKETS;
constexpr int kMaxChunks = LRB_PART_MAX_CHUNKS;
typedef unsigned long long ull;

// device-side bookkeeping, lives in lrb_partition.small (LRB_PART_SMALL_U64 u64)
struct PartMeta {
    ull counts[kMaxChunks][kMaxBuckets];   // entries of (chunk, bucket)
    ull offsets[kMaxChunks][kMaxBuckets];  // first entry of (chunk, bucket) in keys[] / rids[]
    ull cursor[kMaxChunks][kMaxBuckets];   // fill cursors used by k_partition
    ull chunk_base[kMaxChunks + 1];        // first entry of each chunk's region
    ull needed;                            // entries the chunks added so far need in total
    ull overflow;                          // != 0: capacity exceeded, lists are incomplete (apply does nothing)
};
static_assert(sizeof(PartMeta) <= sizeof(ull) * LRB_PART_SMALL_U64, "lrb_partition.small too small");

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
              uint32_t key_lo, uint32_t key_hi, int shift, ull* __restrict__ counts) {
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
    if (threadIdx.x < kMaxBuckets && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (ull)s_cnt[threadIdx.x]);
}

// one warp: offsets of chunk c = chunk_base[c] + exclusive scan of its bucket counts; guards the capacity
__global__ void k_chunk_scan(PartMeta* __restrict__ m, int c, int nb, ull capacity) {
    const int lane = threadIdx.x;
    const ull c0 = (lane < nb) ? m->counts[c][lane] : 0ull;
    const ull c1 = (lane + 32 < nb) ? m->counts[c][lane + 32] : 0ull;
    ull x0 = c0, x1 = c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const ull y0 = __shfl_up_sync(0xFFFFFFFFu, x0, d), y1 = __shfl_up_sync(0xFFFFFFFFu, x1, d);
        if (lane >= d) { x0 += y0; x1 += y1; }
    }
    const ull tot0 = __shfl_sync(0xFFFFFFFFu, x0, 31);
    const ull total = tot0 + __shfl_sync(0xFFFFFFFFu, x1, 31);
    const ull base = m->chunk_base[c];
    const bool fits = (m->overflow == 0) && (base + total <= capacity);
    __syncwarp();
    m->offsets[c][lane] = base + x0 - c0;
    m->offsets[c][lane + 32] = base + tot0 + x1 - c1;
    if (!fits) {  // keep the lists consistent (this chunk contributes nothing) and remember how much was needed
        m->counts[c][lane] = 0;
        m->counts[c][lane + 32] = 0;
    }
    if (lane == 0) {
        m->needed += total;
        m->chunk_base[c + 1] = fits ? base + total : base;
        if (!fits) m->overflow = 1;
    }
}

// (key[, read]) of every valid window of the chunk -> its bucket's region, coalesced through a shared-memory stage
template <bool WITH_RID, bool FULL>
__global__ void __launch_bounds__(kPartThreads, 3)
k_partition(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ blk_read,
            uint64_t blk_lo, uint64_t blk_hi, uint32_t key_lo, uint32_t key_hi, int shift, int nb, PartMeta* __restrict__ meta,
            int c, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ rids_out) {
    __shared__ uint32_t s_key[kChunkSlots];   // staged keys, grouped by bucket
    __shared__ uint8_t s_tid[kChunkSlots];    // which thread (hence which read) staged the entry
    __shared__ uint32_t s_rid[kPartThreads];
    __shared__ uint32_t s_cnt[kMaxBuckets], s_base[kMaxBuckets];
    __shared__ ull s_gbase[kMaxBuckets];
    if (meta->overflow) return;
    const ull* __restrict__ offsets = meta->offsets[c];
    ull* __restrict__ cursor = meta->cursor[c];
    const uint32_t bucket0 = key_lo >> shift;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t n_steps = (blk_hi - blk_lo + kPartThreads - 1) / kPartThreads;

    for (uint64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
        if (tid < kMaxBuckets) s_cnt[tid] = 0;
        __syncthreads();
        const uint64_t gb = blk_lo + step * kPartThreads + tid;
        uint32_t key[32];
        uint32_t posw[16];  // rank inside the CTA's bucket run, two 16-bit values per word
        uint32_t m = 0;
        if (gb < blk_hi) {
            const BlockWindows b = load_block(codes, valid, gb);
            m = b.m;
            if (m) {
                if (WITH_RID) s_rid[tid] = __ldg(blk_read + gb);
                const uint32_t r0 = rc16(b.w1), r1 = rc16(b.w0), r2 = rc16(b.pw);
                if (FULL && m == 0xFFFFFFFFu) {  // interior block of a read, whole key space: no predicates at all
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t kk = canonical15(kmer_ending_at<15>(b.pw, b.w0, b.w1, j), rc15_ending_at(r0, r1, r2, j));
                        key[j] = kk;
                        const uint32_t p = atomicAdd(&s_cnt[kk >> shift], 1u);
                        if (j & 1) posw[j >> 1] |= p << 16; else posw[j >> 1] = p;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((j & 1) == 0) posw[j >> 1] = 0;
                        key[j] = 0;
                        if ((m >> j) & 1u) {
                            const uint32_t kk = canonical15(kmer_ending_at<15>(b.pw, b.w0, b.w1, j), rc15_ending_at(r0, r1, r2, j));
                            if (FULL || (kk >= key_lo && kk < key_hi)) {
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
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the bucket counts (nb <= 64: two per lane) + global reservation
            const uint32_t c0 = (lane < nb) ? s_cnt[lane] : 0u;
            const uint32_t c1 = (lane + 32 < nb) ? s_cnt[lane + 32] : 0u;
            uint32_t x0 = c0, x1 = c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y0 = __shfl_up_sync(0xFFFFFFFFu, x0, d), y1 = __shfl_up_sync(0xFFFFFFFFu, x1, d);
                if (lane >= d) { x0 += y0; x1 += y1; }
            }
            const uint32_t tot0 = __shfl_sync(0xFFFFFFFFu, x0, 31);
            s_base[lane] = x0 - c0;
            s_base[lane + 32] = tot0 + x1 - c1;
            if (lane < nb && c0) s_gbase[lane] = offsets[lane] + atomicAdd(&cursor[lane], (ull)c0);
            if (lane + 32 < nb && c1) s_gbase[lane + 32] = offsets[lane + 32] + atomicAdd(&cursor[lane + 32], (ull)c1);
        }
        __syncthreads();
        if (m == 0xFFFFFFFFu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t idx = s_base[(key[j] >> shift) - bucket0] + ((j & 1) ? (posw[j >> 1] >> 16) : (posw[j >> 1] & 0xFFFFu));
                s_key[idx] = key[j];
                if (WITH_RID) s_tid[idx] = (uint8_t)tid;
            }
        } else if (m) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if ((m >> j) & 1u) {
                    const uint32_t idx = s_base[(key[j] >> shift) - bucket0] + ((posw[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                    s_key[idx] = key[j];
                    if (WITH_RID) s_tid[idx] = (uint8_t)tid;
                }
            }
        }
        __syncthreads();
        for (int b = warp; b < nb; b += kPartThreads / 32) {  // each warp copies whole bucket runs: contiguous both sides
            const uint32_t cnt = s_cnt[b];
            if (!cnt) continue;
            const uint32_t beg = s_base[b];
            const ull g = s_gbase[b];
            for (uint32_t i = lane; i < cnt; i += 32) {
                __stcs(keys_out + g + i, s_key[beg + i]);
                if (WITH_RID) __stcs(rids_out + g + i, s_rid[s_tid[beg + i]]);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_count_keys(const uint32_t* __restrict__ keys, const PartMeta* __restrict__ meta, int bucket, int n_chunks,
             uint32_t* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int c = 0; c < n_chunks; ++c) {
        const uint64_t n = meta->counts[c][bucket];
        const uint32_t* __restrict__ kk = keys + meta->offsets[c][bucket];
        uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            const uint32_t k0 = __ldcs(kk + i), k1 = __ldcs(kk + i + stride), k2 = __ldcs(kk + i + 2 * stride),
                           k3 = __ldcs(kk + i + 3 * stride);
            atomicAdd(table + k0, 1u);
            atomicAdd(table + k1, 1u);
            atomicAdd(table + k2, 1u);
            atomicAdd(table + k3, 1u);
        }
        for (; i < n; i += stride) atomicAdd(table + __ldcs(kk + i), 1u);
    }
}

// entries of one bucket: hist[read][bin(table[key])] += 1, sums[read] += 1, aggregated over runs of equal neighbours
__global__ void __launch_bounds__(256)
k_search_keys(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rids, const PartMeta* __restrict__ meta,
              int bucket, int n_chunks, const uint32_t* __restrict__ table, uint32_t S32, uint64_t magic, uint32_t B,
              uint32_t* __restrict__ hist, uint32_t* __restrict__ sums) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int c = 0; c < n_chunks; ++c) {
        const uint64_t n = meta->counts[c][bucket];
        const uint64_t off = meta->offsets[c][bucket];
        const uint32_t* __restrict__ kk = keys + off;
        const uint32_t* __restrict__ rr = rids + off;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += stride) {
            const bool act = i < n;
            const uint32_t key = act ? __ldcs(kk + i) : 0u;
            const uint32_t rid = act ? __ldcs(rr + i) : 0xFFFFFFFFu;
            const uint32_t cnt = act ? table[key] : 0u;  // slice resident in L2 (just counted, or warmed by earlier gathers)
            const uint32_t bin = act ? coverage_bin(cnt, S32, magic, B) : 0xFFFFu;
            // one RED per distinct (read, bin) and one per distinct read in the warp: consecutive list entries come from
            // the same few reads and bins, and the L1/LSU sector rate (gather + REDs) is what bounds this kernel
            const uint32_t m_r = __match_any_sync(0xFFFFFFFFu, rid);
            const uint32_t m_g = m_r & __match_any_sync(0xFFFFFFFFu, bin);
            if (act) {
                if ((uint32_t)__ffs(m_g) - 1u == lane) atomicAdd(hist + (size_t)rid * B + bin, (uint32_t)__popc(m_g));
                if ((uint32_t)__ffs(m_r) - 1u == lane) atomicAdd(sums + rid, (uint32_t)__popc(m_r));
            }
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

// ---- host side: begin, add chunk(s) of the stream, then count and/or search bucket by bucket -------------------
extern "C" int lrb_dev_partition_begin(lrb_partition* part, int with_rids, uint32_t key_lo, uint32_t key_hi,
                                       int log2_bucket_keys, void* stream) {
    if (!part || !part->keys || !part->small) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_begin: null argument");
    if (with_rids && !part->rids) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_begin: read ids need part->rids");
    if (key_hi > kTableEntries) key_hi = kTableEntries;
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
    part->key_hi = key_hi;
    part->has_rids = with_rids ? 1 : 0;
    part->n_chunks = 0;
    LRB_CUDA(cudaMemsetAsync(part->small, 0, sizeof(PartMeta), (cudaStream_t)stream));
    return LRB_OK;
}

extern "C" int lrb_dev_partition_add(const lrb_reads_view* dev, const uint32_t* blk_read, uint64_t blk_lo, uint64_t blk_hi,
                                     lrb_partition* part, void* stream) {
    if (!dev || !part || !part->keys || !part->small) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_add: null argument");
    if (part->has_rids && !blk_read) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_add: read ids need blk_read");
    if (blk_hi > dev->n_blocks) blk_hi = dev->n_blocks;
    if (blk_lo >= blk_hi) return LRB_OK;
    if (part->n_chunks >= kMaxChunks) return lrb_set_error(LRB_EINVAL, "too many chunks in one partition (max %d)", kMaxChunks);
    const int c = part->n_chunks++;
    cudaStream_t st = (cudaStream_t)stream;
    PartMeta* meta = reinterpret_cast<PartMeta*>(part->small);
    const uint64_t nblk = blk_hi - blk_lo;
    const int nsm = sms();
    const int nb = part->n_buckets, shift = part->shift;
    {
        const uint64_t want = (nblk + 255) / 256;
        const unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)nsm * 8);
        k_bucket_hist<<<grid, 256, 0, st>>>(dev->codes, dev->valid, blk_lo, blk_hi, part->key_lo, part->key_hi, shift, &meta->counts[c][0]);
        k_chunk_scan<<<1, 32, 0, st>>>(meta, c, nb, (ull)part->capacity);
    }
    const uint64_t n_steps = (nblk + kPartThreads - 1) / kPartThreads;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_steps, (uint64_t)nsm * 24);
    const bool full = part->key_lo == 0 && part->key_hi >= kTableEntries;
#define LRB_LAUNCH_PART(RID, FULLK)                                                                                              \
    k_partition<RID, FULLK><<<grid, kPartThreads, 0, st>>>(dev->codes, dev->valid, blk_read, blk_lo, blk_hi, part->key_lo, part->key_hi, \
                                                           shift, nb, meta, c, part->keys, part->rids)
    if (part->has_rids) { if (full) LRB_LAUNCH_PART(true, true); else LRB_LAUNCH_PART(true, false); }
    else { if (full) LRB_LAUNCH_PART(false, true); else LRB_LAUNCH_PART(false, false); }
#undef LRB_LAUNCH_PART
    LRB_CUDA(cudaGetLastError());
    return LRB_OK;
}

// One-shot convenience: begin + a single chunk.
extern "C" int lrb_dev_partition_build(const lrb_reads_view* dev, const uint32_t* blk_read, int with_rids, uint64_t blk_lo,
                                       uint64_t blk_hi, uint32_t key_lo, uint32_t key_hi, int log2_bucket_keys,
                                       lrb_partition* part, void* stream) {
    int rc = lrb_dev_partition_begin(part, with_rids, key_lo, key_hi, log2_bucket_keys, stream);
    if (rc) return rc;
    return lrb_dev_partition_add(dev, blk_read, blk_lo, blk_hi, part, stream);
}

// Synchronises the stream and reports whether the lists are complete; *needed = entries required in total.
extern "C" int lrb_dev_partition_check(const lrb_partition* part, uint64_t* needed, void* stream) {
    if (!part || !part->small) return lrb_set_error(LRB_EINVAL, "lrb_dev_partition_check: null argument");
    ull h[2] = {0, 0};
    const PartMeta* meta = reinterpret_cast<const PartMeta*>(part->small);
    LRB_CUDA(cudaMemcpyAsync(h, &meta->needed, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (needed) *needed = h[0];
    if (h[1])
        return lrb_set_error(LRB_ENOMEM, "partition workspace too small: %llu entries needed, %llu available", h[0], (ull)part->capacity);
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
    if (part->n_chunks == 0) return LRB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t S32 = bin_size > 0xFFFFFFFFl ? 0xFFFFFFFFu : (uint32_t)(bin_size > 0 ? bin_size : 1);
    const uint64_t magic = coverage_magic(S32);
    const PartMeta* meta = reinterpret_cast<const PartMeta*>(part->small);
    const unsigned grid = (unsigned)sms() * 8;
    for (int b = 0; b < part->n_buckets; ++b) {
        if (do_count) k_count_keys<<<grid, 256, 0, st>>>(part->keys, meta, b, part->n_chunks, table);
        if (do_search) k_search_keys<<<grid, 256, 0, st>>>(part->keys, part->rids, meta, b, part->n_chunks, table, S32, magic, (uint32_t)bins, hist, sums);
    }
    LRB_CUDA(cudaGetLastError());
    return LRB_OK;
}
