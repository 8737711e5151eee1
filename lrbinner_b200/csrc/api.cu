// api.cu — context / host-buffer / file level of the C ABI (see include/lrbinner_b200.h).
//
// lrb_ctx owns the device copy of one read set, the 4 GiB 15-mer table, the result buffers and the
// stream they are driven on.  lrb_profile_host is the seam-to-seam path: packed reads in (pinned) host
// memory -> H2D -> composition -> count -> mirror -> search -> D2H.  The file-level functions are the
// drop-ins for the reference's three executables (argv contracts in the header).
#include <cuda_runtime.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "lane_core.cuh"
#include "synth_core.h"

using namespace lrb;

// ---- host allocation -------------------------------------------------------------------------
void* lrb_host_alloc(size_t bytes, bool* pinned, bool want_pinned) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (want_pinned) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
            *pinned = true;
            return p;
        }
        cudaGetLastError();  // no device (host-only tests): plain memory; any later device call reports the real error
    }
    *pinned = false;
    if (posix_memalign(&p, 4096, bytes) != 0) return nullptr;
    return p;
}

void lrb_host_free(void* p, bool pinned) {
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}

extern "C" void* lrb_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable) != cudaSuccess) {
        lrb_set_error(LRB_ECUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" void lrb_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// ---- context -----------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return LRB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return lrb_set_error(LRB_ECUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return LRB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct lrb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t sync_ev[LRB_PART_MAX_CHUNKS + 2] = {};  // [0] index arrays, [1..] H2D chunks, [last] composition D2H
    DevBuf codes, valid, read_len, read_blk, tile_read, tile_blk;
    DevBuf table, comp, hist, sums, text;
    DevBuf part_keys, part_small, part_steps, part_sub, blk_read;  // L2-resident (partitioned) table passes
    DevBuf exc_blk, exc_valid;                            // validity exceptions (lrb_dev_fill_valid)
    lrb_partition part = {};
    bool table_ready = false;  // holds a complete (mirrored) table
    lrb_reads_view dview = {};
    float ms[7] = {0, 0, 0, 0, 0, 0, 0};
};

#define CTX_CUDA(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return lrb_set_error(LRB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" int lrb_ctx_create(int device, lrb_ctx** out) {
    if (!out) return lrb_set_error(LRB_EINVAL, "lrb_ctx_create: null out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return lrb_set_error(LRB_ECUDA, "no CUDA device available (%s): liblrb200 has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return lrb_set_error(LRB_EINVAL, "device %d out of range (have %d)", device, n);
    CTX_CUDA(cudaSetDevice(device));
    lrb_ctx* c = new lrb_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return lrb_set_error(LRB_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        lrb_ctx_destroy(c);
        return lrb_set_error(LRB_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    for (auto& ev : c->ev) cudaEventCreate(&ev);
    for (auto& ev : c->sync_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    *out = c;
    return LRB_OK;
}

extern "C" void lrb_ctx_destroy(lrb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->codes, &c->valid, &c->read_len, &c->read_blk, &c->tile_read, &c->tile_blk, &c->table, &c->comp,
                      &c->hist, &c->sums, &c->text, &c->part_keys, &c->part_small, &c->part_steps, &c->part_sub, &c->blk_read, &c->exc_blk, &c->exc_valid})
        b->release();
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->sync_ev) if (ev) cudaEventDestroy(ev);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int comp_width(int k) { return k == 3 ? 32 : k == 4 ? 136 : k == 5 ? 512 : 0; }

static int upload_reads(lrb_ctx* c, const lrb_reads* r) {
    const uint64_t nb = r->n_blocks, n = r->n_reads, nt = r->n_tiles;
    int rc;
    if ((rc = c->codes.reserve(sizeof(uint32_t) * (2 * nb + 2)))) return rc;
    if ((rc = c->valid.reserve(sizeof(uint32_t) * (nb + 1)))) return rc;
    if ((rc = c->read_len.reserve(sizeof(uint32_t) * (n + 1)))) return rc;
    if ((rc = c->read_blk.reserve(sizeof(uint32_t) * (n + 1)))) return rc;
    if ((rc = c->tile_read.reserve(sizeof(uint32_t) * (nt + 1)))) return rc;
    if ((rc = c->tile_blk.reserve(sizeof(uint32_t) * (nt + 1)))) return rc;
    cudaStream_t st = c->stream;
    CTX_CUDA(cudaMemcpyAsync(c->codes.p, r->codes, sizeof(uint32_t) * (2 * nb + 2), cudaMemcpyHostToDevice, st));
    CTX_CUDA(cudaMemcpyAsync(c->valid.p, r->valid, sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, st));
    if (n) CTX_CUDA(cudaMemcpyAsync(c->read_len.p, r->read_len, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    CTX_CUDA(cudaMemcpyAsync(c->read_blk.p, r->read_blk, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, st));
    if (nt) {
        CTX_CUDA(cudaMemcpyAsync(c->tile_read.p, r->tile_read, sizeof(uint32_t) * nt, cudaMemcpyHostToDevice, st));
        CTX_CUDA(cudaMemcpyAsync(c->tile_blk.p, r->tile_blk, sizeof(uint32_t) * nt, cudaMemcpyHostToDevice, st));
    }
    lrb_reads_view& v = c->dview;
    v.n_reads = n; v.n_blocks = nb; v.n_tiles = nt; v.total_bases = r->total_bases;
    v.codes = (const uint32_t*)c->codes.p; v.valid = (const uint32_t*)c->valid.p;
    v.read_len = (const uint32_t*)c->read_len.p; v.read_blk = (const uint32_t*)c->read_blk.p;
    v.tile_read = (const uint32_t*)c->tile_read.p; v.tile_blk = (const uint32_t*)c->tile_blk.p;
    return LRB_OK;
}

// Seam-to-seam pipeline.  Three streams: copy_in streams the packed reads in chunks (cut at read boundaries),
// `stream` runs composition + partition of chunk i as soon as it has landed (so the kernels that only need the
// reads overlap the PCIe transfer), then the table passes; copy_out returns the composition rows while the
// table passes run.  Phase timings (lrb_ctx_last_timings) therefore overlap; [6] is the wall of the whole call.
extern "C" int lrb_profile_host(lrb_ctx* c, const lrb_reads* r, int k, long bin_size, int bins, uint32_t* comp_counts,
                                uint32_t* cov_hist, uint32_t* cov_sums, uint32_t* table_host, int use_loaded_table) {
    if (!c || !r) return lrb_set_error(LRB_EINVAL, "lrb_profile_host: null argument");
    const int P = comp_width(k);
    if (comp_counts && !P) return lrb_set_error(LRB_EINVAL, "k must be 3, 4 or 5 (got %d)", k);
    const bool do_search = cov_hist || cov_sums;
    if (do_search && (!cov_hist || !cov_sums)) return lrb_set_error(LRB_EINVAL, "cov_hist and cov_sums go together");
    if (do_search && bin_size <= 0) return lrb_set_error(LRB_EINVAL, "bin_size must be >= 1 (the reference divides by it)");
    if (do_search && (bins <= 0 || bins > LRB_MAX_BINS)) return lrb_set_error(LRB_EINVAL, "bins must be in [1, %d]", LRB_MAX_BINS);
    if (use_loaded_table && !c->table_ready) return lrb_set_error(LRB_EINVAL, "no table loaded in this context");
    const bool do_count = !use_loaded_table && (do_search || table_host);
    CTX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream, sin = c->copy_in, sout = c->copy_out;
    const uint64_t n = r->n_reads, nb = r->n_blocks, nt = r->n_tiles;
    int rc;
    // ---- allocate before timing starts ------------------------------------------------------------------
    if ((rc = c->codes.reserve(sizeof(uint32_t) * (2 * nb + 2)))) return rc;
    if ((rc = c->valid.reserve(sizeof(uint32_t) * (nb + 1)))) return rc;
    if ((rc = c->read_len.reserve(sizeof(uint32_t) * (n + 1)))) return rc;
    if ((rc = c->read_blk.reserve(sizeof(uint32_t) * (n + 1)))) return rc;
    if ((rc = c->tile_read.reserve(sizeof(uint32_t) * (nt + 1)))) return rc;
    if ((rc = c->tile_blk.reserve(sizeof(uint32_t) * (nt + 1)))) return rc;
    if (comp_counts && (rc = c->comp.reserve(std::max<size_t>(16, sizeof(uint32_t) * n * P)))) return rc;
    if (do_search) {
        if ((rc = c->hist.reserve(std::max<size_t>(16, sizeof(uint32_t) * n * (size_t)bins)))) return rc;
        if ((rc = c->sums.reserve(std::max<size_t>(16, sizeof(uint32_t) * n)))) return rc;
    }
    if ((do_count || do_search) && (rc = c->table.reserve(sizeof(uint32_t) * (size_t)kTableEntries))) return rc;
    // validity bitmap: when the reads carry their exception list (blocks whose valid word is not implied by the read
    // length) and it is short, only codes + exceptions cross PCIe (0.25 instead of 0.375 B/base) and the bitmap is
    // rebuilt on the device; LRB_SHIP_VALID=1 forces the plain copy.
    bool ship_valid = !(r->exc_ready && r->n_exc <= nb / 16);
    {
        const char* e = getenv("LRB_SHIP_VALID");
        if (e && atoi(e) > 0) ship_valid = true;
    }
    if (!ship_valid) {
        if ((rc = c->exc_blk.reserve(sizeof(uint32_t) * (r->n_exc + 1)))) return rc;
        if ((rc = c->exc_valid.reserve(sizeof(uint32_t) * (r->n_exc + 1)))) return rc;
    }
    // table passes: key-partitioned + L2-resident (csrc/partition.cu) unless LRB_TABLE_PATH=direct or its
    // workspace (4 B per slot) does not fit; the direct kernels (one random HBM access per window) remain as
    // the small-memory GPU path.  Both are bit-identical.
    bool use_part = (do_count || do_search) && nb > 0;
    {
        const char* e = getenv("LRB_TABLE_PATH");
        if (e && !strcmp(e, "direct")) use_part = false;
    }
    int bucket_shift = 24;
    {
        const char* e = getenv("LRB_BUCKET_LOG2");
        if (e && atoi(e) >= 24 && atoi(e) <= 25) bucket_shift = atoi(e);
    }
    if (use_part) {
        const size_t cap = std::max<uint64_t>(nb * 32, 1);  // >= number of slots: the lists can never overflow
        const uint64_t step_cap = lrb_partition_step_capacity(nb, LRB_PART_MAX_CHUNKS);
        if (c->part_keys.reserve(sizeof(uint32_t) * cap) || c->part_steps.reserve(sizeof(uint32_t) * lrb_partition_steps_words(step_cap)) ||
            c->part_small.reserve(sizeof(unsigned long long) * LRB_PART_SMALL_U64) || c->blk_read.reserve(sizeof(uint32_t) * (nb + 1))) {
            cudaGetLastError();
            c->part_keys.release();
            c->part_steps.release();
            use_part = false;
        } else {
            c->part.keys = (uint32_t*)c->part_keys.p;
            c->part.steps = (uint32_t*)c->part_steps.p;
            c->part.small = (unsigned long long*)c->part_small.p;
            c->part.capacity = cap;
            c->part.step_capacity = step_cap;
            // second-level lists for the shared-memory count: worth it once the read set is large; optional
            c->part.sub = nullptr;
            c->part.sub_capacity = 0;
            const char* e = getenv("LRB_COUNT_PATH");
            const bool want_smem = e ? !strcmp(e, "smem") : nb >= (1u << 19);
            if (do_count && want_smem && !(e && !strcmp(e, "l2"))) {
                const size_t sub_cap = 2 * cap + (1u << 22);  // 2-byte entries in fixed-size segments: 2x headroom + fill counters and dump areas
                if (c->part_sub.reserve(sizeof(uint16_t) * sub_cap) == LRB_OK) {
                    c->part.sub = (uint16_t*)c->part_sub.p;
                    c->part.sub_capacity = sub_cap;
                } else {
                    cudaGetLastError();
                }
            }
        }
    }
    lrb_reads_view& v = c->dview;
    v.n_reads = n; v.n_blocks = nb; v.n_tiles = nt; v.total_bases = r->total_bases;
    v.codes = (const uint32_t*)c->codes.p; v.valid = (const uint32_t*)c->valid.p;
    v.read_len = (const uint32_t*)c->read_len.p; v.read_blk = (const uint32_t*)c->read_blk.p;
    v.tile_read = (const uint32_t*)c->tile_read.p; v.tile_blk = (const uint32_t*)c->tile_blk.p;

    // ---- chunk plan: cut at read boundaries, roughly equal numbers of blocks ---------------------------------
    int n_chunks = 1;
    {
        const char* e = getenv("LRB_H2D_CHUNKS");
        const int want = e && atoi(e) > 0 ? atoi(e) : 16;
        n_chunks = (int)std::min<uint64_t>((uint64_t)std::min(want, LRB_PART_MAX_CHUNKS), std::max<uint64_t>(1, nb >> 16));
    }
    std::vector<uint64_t> cr(n_chunks + 1), cb(n_chunks + 1), ct(n_chunks + 1);  // read / block / tile cut points
    cr[0] = cb[0] = ct[0] = 0;
    for (int i = 1; i < n_chunks; ++i) {
        const uint32_t target = (uint32_t)(nb * (uint64_t)i / n_chunks);
        uint64_t ri = (uint64_t)(std::upper_bound(r->read_blk, r->read_blk + n, target) - r->read_blk);
        if (ri > 0) --ri;
        ri = std::max(ri, cr[i - 1]);
        cr[i] = ri;
        cb[i] = r->read_blk[ri];
        ct[i] = (uint64_t)(std::lower_bound(r->tile_read, r->tile_read + nt, (uint32_t)ri) - r->tile_read);
    }
    cr[n_chunks] = n; cb[n_chunks] = nb; ct[n_chunks] = nt;

    // ---- go --------------------------------------------------------------------------------------------------
    CTX_CUDA(cudaEventRecord(c->ev[0], st));
    CTX_CUDA(cudaStreamWaitEvent(sin, c->ev[0], 0));
    CTX_CUDA(cudaStreamWaitEvent(sout, c->ev[0], 0));
    if (n) CTX_CUDA(cudaMemcpyAsync(c->read_len.p, r->read_len, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, sin));
    CTX_CUDA(cudaMemcpyAsync(c->read_blk.p, r->read_blk, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, sin));
    if (nt) {
        CTX_CUDA(cudaMemcpyAsync(c->tile_read.p, r->tile_read, sizeof(uint32_t) * nt, cudaMemcpyHostToDevice, sin));
        CTX_CUDA(cudaMemcpyAsync(c->tile_blk.p, r->tile_blk, sizeof(uint32_t) * nt, cudaMemcpyHostToDevice, sin));
    }
    if (!ship_valid && r->n_exc) {
        CTX_CUDA(cudaMemcpyAsync(c->exc_blk.p, r->exc_blk, sizeof(uint32_t) * r->n_exc, cudaMemcpyHostToDevice, sin));
        CTX_CUDA(cudaMemcpyAsync(c->exc_valid.p, r->exc_valid, sizeof(uint32_t) * r->n_exc, cudaMemcpyHostToDevice, sin));
    }
    CTX_CUDA(cudaEventRecord(c->sync_ev[0], sin));
    for (int i = 0; i < n_chunks; ++i) {
        const uint64_t b0 = cb[i], b1 = cb[i + 1];
        const uint64_t w0 = 2 * b0, w1 = (i == n_chunks - 1) ? 2 * nb + 2 : 2 * b1;
        const uint64_t v1 = (i == n_chunks - 1) ? nb + 1 : b1;
        CTX_CUDA(cudaMemcpyAsync((uint32_t*)c->codes.p + w0, r->codes + w0, sizeof(uint32_t) * (w1 - w0), cudaMemcpyHostToDevice, sin));
        if (ship_valid)
            CTX_CUDA(cudaMemcpyAsync((uint32_t*)c->valid.p + b0, r->valid + b0, sizeof(uint32_t) * (v1 - b0), cudaMemcpyHostToDevice, sin));
        CTX_CUDA(cudaEventRecord(c->sync_ev[1 + i], sin));
    }
    CTX_CUDA(cudaEventRecord(c->ev[1], sin));  // H2D done

    // compute stream: zero the outputs while the first chunk is in flight
    if (comp_counts && n) CTX_CUDA(cudaMemsetAsync(c->comp.p, 0, sizeof(uint32_t) * n * P, st));
    if (do_search && n) {
        CTX_CUDA(cudaMemsetAsync(c->hist.p, 0, sizeof(uint32_t) * n * (size_t)bins, st));
        CTX_CUDA(cudaMemsetAsync(c->sums.p, 0, sizeof(uint32_t) * n, st));
    }
    if (do_count) {
        c->table_ready = false;
        CTX_CUDA(cudaMemsetAsync(c->table.p, 0, sizeof(uint32_t) * (size_t)kTableEntries, st));
    }
    CTX_CUDA(cudaStreamWaitEvent(st, c->sync_ev[0], 0));  // index arrays (and validity exceptions) are on the device
    if (!ship_valid && (rc = lrb_dev_fill_valid(&v, (const uint32_t*)c->exc_blk.p, (const uint32_t*)c->exc_valid.p, r->n_exc, st))) return rc;
    if (use_part) {
        if (do_search && (rc = lrb_dev_fill_blk_read(&v, (uint32_t*)c->blk_read.p, st))) return rc;
        if ((rc = lrb_dev_partition_begin(&c->part, do_search ? 1 : 0, 0, kTableEntries, bucket_shift, st))) return rc;
    }
    for (int i = 0; i < n_chunks; ++i) {
        CTX_CUDA(cudaStreamWaitEvent(st, c->sync_ev[1 + i], 0));
        if (comp_counts && n && (rc = lrb_dev_composition(&v, k, (uint32_t*)c->comp.p, ct[i], ct[i + 1], st))) return rc;
        if (use_part) {
            if ((rc = lrb_dev_partition_add(&v, (const uint32_t*)c->blk_read.p, cb[i], cb[i + 1], &c->part, st))) return rc;
        } else if (do_count) {
            if ((rc = lrb_dev_count(&v, (uint32_t*)c->table.p, cb[i], cb[i + 1], 0, kTableEntries, st))) return rc;
        }
    }
    CTX_CUDA(cudaEventRecord(c->ev[2], st));  // composition (+ partition) of every chunk done
    if (comp_counts && n) {                   // composition rows go home while the table passes run
        CTX_CUDA(cudaStreamWaitEvent(sout, c->ev[2], 0));
        CTX_CUDA(cudaMemcpyAsync(comp_counts, c->comp.p, sizeof(uint32_t) * n * P, cudaMemcpyDeviceToHost, sout));
    }
    CTX_CUDA(cudaEventRecord(c->sync_ev[LRB_PART_MAX_CHUNKS + 1], sout));
    if (use_part) {
        // count every bucket, then search every bucket: each pass is ONE launch over all buckets when the second-level lists
        // exist (no per-bucket tails: -2.9 ms at config #2), which more than pays for reading the table slices twice
        const bool batched = do_count && c->part.sub;
        if (batched) {
            if ((rc = lrb_dev_partition_apply(&c->part, 1 | 4, (uint32_t*)c->table.p, bin_size, bins, nullptr, nullptr, st))) return rc;
            if (do_search && n && (rc = lrb_dev_partition_apply(&c->part, 2, (uint32_t*)c->table.p, bin_size, bins, (uint32_t*)c->hist.p,
                                                                (uint32_t*)c->sums.p, st)))
                return rc;
        } else {
            const int mode = (do_count ? 1 : 0) | (do_search && n ? 2 : 0);
            if (mode && (rc = lrb_dev_partition_apply(&c->part, mode, (uint32_t*)c->table.p, bin_size, bins, (uint32_t*)c->hist.p,
                                                      (uint32_t*)c->sums.p, st)))
                return rc;
        }
    }
    CTX_CUDA(cudaEventRecord(c->ev[3], st));  // table passes done
    if (do_count) {
        if ((rc = lrb_dev_mirror((uint32_t*)c->table.p, st))) return rc;
        c->table_ready = true;
    }
    CTX_CUDA(cudaEventRecord(c->ev[4], st));
    if (!use_part && do_search && n) {
        if ((rc = lrb_dev_search(&v, (const uint32_t*)c->table.p, bin_size, bins, (uint32_t*)c->hist.p, (uint32_t*)c->sums.p, 0, nt, 0,
                                 kTableEntries, st)))
            return rc;
    }
    CTX_CUDA(cudaEventRecord(c->ev[5], st));
    if (do_search && n) {
        CTX_CUDA(cudaMemcpyAsync(cov_hist, c->hist.p, sizeof(uint32_t) * n * (size_t)bins, cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaMemcpyAsync(cov_sums, c->sums.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    }
    if (table_host) CTX_CUDA(cudaMemcpyAsync(table_host, c->table.p, sizeof(uint32_t) * (size_t)kTableEntries, cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamWaitEvent(st, c->sync_ev[LRB_PART_MAX_CHUNKS + 1], 0));  // join the composition D2H
    CTX_CUDA(cudaEventRecord(c->ev[6], st));
    CTX_CUDA(cudaStreamSynchronize(st));
    CTX_CUDA(cudaStreamSynchronize(sin));
    CTX_CUDA(cudaStreamSynchronize(sout));
    CTX_CUDA(cudaGetLastError());
    // [0] h2d (copy stream) [1] composition+partition (until the last chunk is processed) [2] table passes
    // [3] mirror [4] direct search [5] result D2H tail [6] whole call
    cudaEventElapsedTime(&c->ms[0], c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->ms[1], c->ev[0], c->ev[2]);
    for (int i = 2; i < 6; ++i) cudaEventElapsedTime(&c->ms[i], c->ev[i], c->ev[i + 1]);
    cudaEventElapsedTime(&c->ms[6], c->ev[0], c->ev[6]);
    return LRB_OK;
}

extern "C" int lrb_ctx_last_timings(const lrb_ctx* c, float* ms7) {
    if (!c || !ms7) return lrb_set_error(LRB_EINVAL, "lrb_ctx_last_timings: null argument");
    memcpy(ms7, c->ms, sizeof c->ms);
    return LRB_OK;
}

// table <-> file through a pair of pinned staging buffers (the 4 GiB never sits in host RAM as a whole)
extern "C" int lrb_ctx_table_save(lrb_ctx* c, const char* path) {
    if (!c || !path) return lrb_set_error(LRB_EINVAL, "lrb_ctx_table_save: null argument");
    if (!c->table_ready) return lrb_set_error(LRB_EINVAL, "lrb_ctx_table_save: no table in this context");
    CTX_CUDA(cudaSetDevice(c->device));
    FILE* f = fopen(path, "wb");
    if (!f) return lrb_set_error(LRB_EIO, "cannot open %s for writing", path);
    const uint64_t size = kTableEntries;
    if (fwrite(&size, sizeof size, 1, f) != 1) { fclose(f); return lrb_set_error(LRB_EIO, "short write to %s", path); }
    const size_t chunk = 64u << 20;
    bool pin = false, pin2 = false;
    char* stage[2] = {(char*)lrb_host_alloc(chunk, &pin), (char*)lrb_host_alloc(chunk, &pin2)};
    int rc = LRB_OK;
    const size_t total = (size_t)size * 4;
    const char* dsrc = (const char*)c->table.p;
    size_t off = 0;
    int cur = 0;
    if (!stage[0] || !stage[1]) rc = lrb_set_error(LRB_ENOMEM, "out of memory (staging)");
    if (!rc && cudaMemcpyAsync(stage[0], dsrc, std::min(chunk, total), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = lrb_set_error(LRB_ECUDA, "D2H failed");
    while (!rc && off < total) {
        const size_t nbytes = std::min(chunk, total - off);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = lrb_set_error(LRB_ECUDA, "D2H failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        const size_t next = off + nbytes;
        if (next < total && cudaMemcpyAsync(stage[cur ^ 1], dsrc + next, std::min(chunk, total - next), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = lrb_set_error(LRB_ECUDA, "D2H failed"); break; }
        if (fwrite(stage[cur], 1, nbytes, f) != nbytes) { rc = lrb_set_error(LRB_EIO, "short write to %s", path); break; }
        off = next;
        cur ^= 1;
    }
    cudaStreamSynchronize(c->stream);
    lrb_host_free(stage[0], pin);
    lrb_host_free(stage[1], pin2);
    if (fclose(f) != 0 && !rc) rc = lrb_set_error(LRB_EIO, "close failed for %s", path);
    return rc;
}

extern "C" int lrb_ctx_table_load(lrb_ctx* c, const char* path) {
    if (!c || !path) return lrb_set_error(LRB_EINVAL, "lrb_ctx_table_load: null argument");
    CTX_CUDA(cudaSetDevice(c->device));
    FILE* f = fopen(path, "rb");
    if (!f) return lrb_set_error(LRB_EIO, "cannot open table file %s", path);
    uint64_t size = 0;
    if (fread(&size, sizeof size, 1, f) != 1 || size != kTableEntries) {
        fclose(f);
        return lrb_set_error(LRB_EFORMAT, "%s is not a 4^15-entry 15mers-counts file", path);
    }
    int rc = c->table.reserve(sizeof(uint32_t) * (size_t)kTableEntries);
    if (rc) { fclose(f); return rc; }
    c->table_ready = false;
    const size_t chunk = 64u << 20;
    bool pin = false, pin2 = false;
    char* stage[2] = {(char*)lrb_host_alloc(chunk, &pin), (char*)lrb_host_alloc(chunk, &pin2)};
    if (!stage[0] || !stage[1]) rc = lrb_set_error(LRB_ENOMEM, "out of memory (staging)");
    const size_t total = (size_t)size * 4;
    size_t off = 0;
    int cur = 0;
    cudaEvent_t done[2];
    cudaEventCreate(&done[0]);
    cudaEventCreate(&done[1]);
    bool used[2] = {false, false};
    while (!rc && off < total) {
        const size_t nbytes = std::min(chunk, total - off);
        if (used[cur]) cudaEventSynchronize(done[cur]);
        if (fread(stage[cur], 1, nbytes, f) != nbytes) { rc = lrb_set_error(LRB_EFORMAT, "%s is truncated", path); break; }
        if (cudaMemcpyAsync((char*)c->table.p + off, stage[cur], nbytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = lrb_set_error(LRB_ECUDA, "H2D failed"); break; }
        cudaEventRecord(done[cur], c->stream);
        used[cur] = true;
        off += nbytes;
        cur ^= 1;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) rc = lrb_set_error(LRB_ECUDA, "H2D failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    lrb_host_free(stage[0], pin);
    lrb_host_free(stage[1], pin2);
    fclose(f);
    if (!rc) c->table_ready = true;
    return rc;
}

// ---- synthetic reads ---------------------------------------------------------------------------
namespace {

struct DevEmit {
    uint32_t* codes;  // first code word of the read
    uint32_t* valid;
    uint32_t w0 = 0, w1 = 0, v = 0;
    __device__ void operator()(uint32_t pos, char c) {
        const uint32_t j = pos & 31u;
        const uint32_t code = ((unsigned char)c >> 1) & 3u;
        if (j < 16) w0 |= code << (30 - 2 * j); else w1 |= code << (62 - 2 * j);
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') v |= 1u << j;
        if (j == 31u) {
            const uint32_t b = pos >> 5;
            codes[2 * (size_t)b] = w0; codes[2 * (size_t)b + 1] = w1; valid[b] = v;
            w0 = w1 = v = 0;
        }
    }
};

__global__ void __launch_bounds__(128)
k_synth(lrb_reads_view R, lrb_synth_params p, const uint32_t* __restrict__ glen, const uint32_t* __restrict__ meta,
        uint32_t* __restrict__ codes, uint32_t* __restrict__ valid) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n_reads) return;
    const uint32_t len = R.read_len[r], b0 = R.read_blk[r];
    DevEmit e{codes + 2 * (size_t)b0, valid + b0};
    synth_read(p, glen, r, meta[4 * r] % p.n_genomes, meta[4 * r + 1], meta[4 * r + 2], len, [&](uint32_t pos, char c) { e(pos, c); });
    const uint32_t b = len >> 5;  // trailing (partial or empty) block
    e.codes[2 * (size_t)b] = e.w0; e.codes[2 * (size_t)b + 1] = e.w1; e.valid[b] = e.v;
}

}  // namespace

extern "C" int lrb_dev_synth(const lrb_reads_view* dev, const lrb_synth_params* p, const uint32_t* glen,
                             const uint32_t* meta, void* stream) {
    if (!dev || !p || !glen || !meta || p->n_genomes == 0) return lrb_set_error(LRB_EINVAL, "lrb_dev_synth: bad argument");
    if (dev->n_reads == 0) return LRB_OK;
    const unsigned grid = (unsigned)((dev->n_reads + 127) / 128);
    k_synth<<<grid, 128, 0, (cudaStream_t)stream>>>(*dev, *p, glen, meta, const_cast<uint32_t*>(dev->codes), const_cast<uint32_t*>(dev->valid));
    CTX_CUDA(cudaGetLastError());
    return LRB_OK;
}

extern "C" int lrb_synth_host(const lrb_synth_params* p, const uint32_t* glen, const uint32_t* meta,
                              const uint32_t* lengths, uint64_t n_reads, const uint64_t* offsets, char* bases) {
    if (!p || !glen || !meta || !lengths || !offsets || !bases || p->n_genomes == 0) return lrb_set_error(LRB_EINVAL, "lrb_synth_host: bad argument");
    for (uint64_t r = 0; r < n_reads; ++r) {
        char* dst = bases + offsets[r];
        synth_read(*p, glen, r, meta[4 * r] % p->n_genomes, meta[4 * r + 1], meta[4 * r + 2], lengths[r],
                   [&](uint32_t pos, char c) { dst[pos] = c; });
    }
    return LRB_OK;
}

// ---- file level (the three tools + the fused stage) ----------------------------------------------
namespace {

struct CtxGuard {
    lrb_ctx* c = nullptr;
    ~CtxGuard() { lrb_ctx_destroy(c); }
};
struct ReadsGuard {
    lrb_reads* r = nullptr;
    ~ReadsGuard() { lrb_reads_free(r); }
};

int default_device() {
    const char* e = getenv("LRB_DEVICE");
    if (e && *e) return atoi(e);
    const char* lr = getenv("LOCAL_RANK");
    if (lr && *lr) return atoi(lr);
    return 0;
}

int truncate_file(const char* path) {  // the tools create/truncate their output before reading (count-kmers.cpp:210)
    FILE* f = fopen(path, "wb");
    if (!f) return lrb_set_error(LRB_EIO, "cannot open %s for writing", path);
    fclose(f);
    return LRB_OK;
}

}  // namespace

extern "C" int lrb_count_kmers(const char* reads_path, const char* out_txt, int k, int threads) {
    if (!reads_path || !out_txt) return lrb_set_error(LRB_EINVAL, "lrb_count_kmers: null path");
    const int P = comp_width(k);
    if (!P) return lrb_set_error(LRB_EINVAL, "k must be 3, 4 or 5 (got %d)", k);
    int rc;
    if ((rc = truncate_file(out_txt))) return rc;
    CtxGuard cg;
    if ((rc = lrb_ctx_create(default_device(), &cg.c))) return rc;
    ReadsGuard rg;
    if ((rc = lrb_reads_from_file(reads_path, threads, &rg.r))) return rc;
    std::vector<uint32_t> counts((size_t)rg.r->n_reads * P + 1);
    if ((rc = lrb_profile_host(cg.c, rg.r, k, 1, 1, counts.data(), nullptr, nullptr, nullptr, 0))) return rc;
    return lrb_write_composition_txt(out_txt, counts.data(), rg.r->read_len, rg.r->n_reads, k, threads);
}

extern "C" int lrb_count_15mers(const char* reads_path, const char* out_table, int threads) {
    if (!reads_path || !out_table) return lrb_set_error(LRB_EINVAL, "lrb_count_15mers: null path");
    int rc;
    CtxGuard cg;
    if ((rc = lrb_ctx_create(default_device(), &cg.c))) return rc;
    ReadsGuard rg;
    if ((rc = lrb_reads_from_file(reads_path, threads, &rg.r))) return rc;
    lrb_ctx* c = cg.c;
    if ((rc = c->table.reserve(sizeof(uint32_t) * (size_t)kTableEntries))) return rc;
    if ((rc = upload_reads(c, rg.r))) return rc;
    CTX_CUDA(cudaMemsetAsync(c->table.p, 0, sizeof(uint32_t) * (size_t)kTableEntries, c->stream));
    if ((rc = lrb_dev_count(&c->dview, (uint32_t*)c->table.p, 0, rg.r->n_blocks, 0, kTableEntries, c->stream))) return rc;
    if ((rc = lrb_dev_mirror((uint32_t*)c->table.p, c->stream))) return rc;
    CTX_CUDA(cudaStreamSynchronize(c->stream));
    c->table_ready = true;
    return lrb_ctx_table_save(c, out_table);
}

extern "C" int lrb_search_15mers(const char* table_path, const char* reads_path, const char* out_txt, long bin_size,
                                 int bins, int threads) {
    if (!table_path || !reads_path || !out_txt) return lrb_set_error(LRB_EINVAL, "lrb_search_15mers: null path");
    if (bin_size <= 0) return lrb_set_error(LRB_EINVAL, "bin_size must be >= 1 (the reference divides by it)");
    if (bins <= 0 || bins > LRB_MAX_BINS) return lrb_set_error(LRB_EINVAL, "bins must be in [1, %d]", LRB_MAX_BINS);
    int rc;
    CtxGuard cg;
    if ((rc = lrb_ctx_create(default_device(), &cg.c))) return rc;
    if ((rc = lrb_ctx_table_load(cg.c, table_path))) return rc;
    if ((rc = truncate_file(out_txt))) return rc;
    ReadsGuard rg;
    if ((rc = lrb_reads_from_file(reads_path, threads, &rg.r))) return rc;
    const uint64_t n = rg.r->n_reads;
    std::vector<uint32_t> hist((size_t)n * bins + 1), sums(n + 1);
    if ((rc = lrb_profile_host(cg.c, rg.r, 0, bin_size, bins, nullptr, hist.data(), sums.data(), nullptr, 1))) return rc;
    return lrb_write_coverage_txt(out_txt, hist.data(), sums.data(), n, bins, threads);
}

extern "C" int lrb_profile(const char* reads_path, const char* out_dir, int k, long bin_size, int bins, int threads,
                           int write_table, int write_npy) {
    if (!reads_path || !out_dir) return lrb_set_error(LRB_EINVAL, "lrb_profile: null path");
    const int P = comp_width(k);
    if (!P) return lrb_set_error(LRB_EINVAL, "k must be 3, 4 or 5 (got %d)", k);
    if (bin_size <= 0) return lrb_set_error(LRB_EINVAL, "bin_size must be >= 1 (the reference divides by it)");
    if (bins <= 0 || bins > LRB_MAX_BINS) return lrb_set_error(LRB_EINVAL, "bins must be in [1, %d]", LRB_MAX_BINS);
    const std::string prof = std::string(out_dir) + "/profiles";
    mkdir(out_dir, 0755);
    mkdir(prof.c_str(), 0755);
    int rc;
    CtxGuard cg;
    if ((rc = lrb_ctx_create(default_device(), &cg.c))) return rc;
    ReadsGuard rg;
    if ((rc = lrb_reads_from_file(reads_path, threads, &rg.r))) return rc;
    const uint64_t n = rg.r->n_reads;
    std::vector<uint32_t> counts((size_t)n * P + 1), hist((size_t)n * bins + 1), sums(n + 1);
    if ((rc = lrb_profile_host(cg.c, rg.r, k, bin_size, bins, counts.data(), hist.data(), sums.data(), nullptr, 0))) return rc;
    if ((rc = lrb_write_composition_txt((prof + "/com_profs").c_str(), counts.data(), rg.r->read_len, n, k, threads))) return rc;
    if ((rc = lrb_write_coverage_txt((prof + "/cov_profs").c_str(), hist.data(), sums.data(), n, bins, threads))) return rc;
    if (write_npy) {
        if ((rc = lrb_write_composition_npy((prof + "/com_profs.npy").c_str(), counts.data(), rg.r->read_len, n, k, threads))) return rc;
        if ((rc = lrb_write_coverage_npy((prof + "/cov_profs.npy").c_str(), hist.data(), sums.data(), n, bins, threads))) return rc;
    }
    if (write_table && (rc = lrb_ctx_table_save(cg.c, (prof + "/15mers-counts").c_str()))) return rc;
    return LRB_OK;
}
