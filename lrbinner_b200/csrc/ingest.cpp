// ingest.cpp — host side of the read stream: FASTA/FASTQ(+gzip) records -> 2-bit packed blocks.
//
// Replaces SeqReader (mbcclr_utils/io_utils.h:133-165) + kseq_read (mbcclr_utils/kseq.h:177-218).
// The file is parsed ONCE; the three tools of the reference each re-parse it.  Record semantics are the
// ones the tools observe (see parse_records below); packing follows csrc/lane_core.cuh.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.h"
#include "lane_core.cuh"

namespace {

// ---- record scanner --------------------------------------------------------------------------
// Works on the whole (decompressed) file image.  Rules, as exercised by the tools:
//  * with no pending header char, skip forward to the next '>' or '@' ANYWHERE (kseq.h:181-185);
//  * the name runs to the first isspace(); unless that was '\n' the rest of the line is a comment;
//    end of data right after the header char ends the stream without a record (kseq.h:187);
//  * sequence lines follow until a line STARTS with '>', '@' or '+'; empty lines are skipped; after a
//    line is appended a single trailing '\r' is dropped when the accumulated sequence is longer than
//    one byte (kseq.h:141,193-197) — not when the line was cut short by end of data with nothing read;
//  * '+' starts a FASTQ quality block: rest of that line ignored; quality lines are appended (same CR
//    rule) while shorter than the sequence; then the next header is searched afresh.  Missing '\n'
//    after '+' or a length mismatch ends the stream WITHOUT emitting the record (kseq.h:209-214);
//  * the tools copy the sequence as a C string (io_utils.h:159): it is cut at the first NUL byte.
struct Record {
    uint64_t off;  // into `pool`
    uint32_t len;
};

static inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

struct Parsed {
    std::string pool;  // concatenated sequences
    std::vector<Record> recs;
};

static void parse_records(const unsigned char* buf, size_t n, Parsed& out) {
    size_t pos = 0;
    int pending = 0;  // header char already consumed
    std::string qual;
    while (true) {
        if (!pending) {
            while (pos < n && buf[pos] != '>' && buf[pos] != '@') ++pos;
            if (pos >= n) return;
            pending = buf[pos++];
        }
        if (pos >= n) return;  // nothing after the header char
        // name
        size_t i = pos;
        while (i < n && !is_space(buf[i])) ++i;
        int delim = 0;
        if (i < n) { delim = buf[i]; pos = i + 1; } else pos = n;
        if (delim != '\n') {  // comment to end of line
            const void* nl = pos < n ? memchr(buf + pos, '\n', n - pos) : nullptr;
            pos = nl ? (size_t)((const unsigned char*)nl - buf) + 1 : n;
        }
        // sequence
        const size_t seq_off = out.pool.size();
        int c = -1;
        while (true) {
            if (pos >= n) { c = -1; break; }
            c = buf[pos++];
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') continue;
            out.pool.push_back((char)c);
            if (pos < n) {
                const void* nl = memchr(buf + pos, '\n', n - pos);
                const size_t e = nl ? (size_t)((const unsigned char*)nl - buf) : n;
                out.pool.append((const char*)buf + pos, e - pos);
                pos = nl ? e + 1 : n;
                if (out.pool.size() - seq_off > 1 && out.pool.back() == '\r') out.pool.pop_back();
            }
        }
        if (c == '>' || c == '@') pending = c;
        size_t seq_len = out.pool.size() - seq_off;
        if (c == '+') {
            const void* nl = pos < n ? memchr(buf + pos, '\n', n - pos) : nullptr;
            if (!nl) { out.pool.resize(seq_off); return; }  // no quality block: stream ends, record dropped
            pos = (size_t)((const unsigned char*)nl - buf) + 1;
            qual.clear();
            while (pos < n) {
                const void* q = memchr(buf + pos, '\n', n - pos);
                const size_t e = q ? (size_t)((const unsigned char*)q - buf) : n;
                qual.append((const char*)buf + pos, e - pos);
                pos = q ? e + 1 : n;
                if (qual.size() > 1 && qual.back() == '\r') qual.pop_back();
                if (!(qual.size() < seq_len)) break;
            }
            pending = 0;
            if (qual.size() != seq_len) { out.pool.resize(seq_off); return; }
        }
        // C-string copy: cut at the first NUL
        if (seq_len) {
            const void* z = memchr(out.pool.data() + seq_off, 0, seq_len);
            if (z) {
                seq_len = (size_t)((const char*)z - (out.pool.data() + seq_off));
                out.pool.resize(seq_off + seq_len);
            }
        }
        if (seq_len > 0xFFFFFFFFull) { out.pool.resize(seq_off); return; }  // > 4 Gbase record: not representable
        out.recs.push_back(Record{seq_off, (uint32_t)seq_len});
    }
}

static bool read_whole_file(const char* path, std::vector<unsigned char>& data) {
    gzFile f = gzopen(path, "rb");  // transparent for uncompressed input, like io_utils.h:143
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    size_t n = 0;
    data.resize(1 << 22);
    for (;;) {
        if (data.size() - n < (1 << 20)) data.resize(data.size() * 2);
        const size_t room = std::min<size_t>(data.size() - n, 1u << 30);
        const int got = gzread(f, data.data() + n, (unsigned)room);
        if (got <= 0) break;  // EOF or a damaged stream: the tools stop quietly at that point too
        n += (size_t)got;
    }
    gzclose(f);
    data.resize(n);
    return true;
}

// ---- packing -----------------------------------------------------------------------------------
struct PackLut {
    uint8_t code[256];
    uint8_t ok[256];
    PackLut() {
        for (int c = 0; c < 256; ++c) {
            code[c] = (uint8_t)((c >> 1) & 3);  // count-kmers.cpp:77, kmer_utils.h:47
            ok[c] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? 1 : 0;  // kmer_utils.h:38
        }
    }
};
static const PackLut g_pack;

static void pack_read(const unsigned char* s, uint32_t len, uint32_t* codes, uint32_t* valid) {
    // codes/valid point at the read's first block; the read owns len/32 + 1 blocks (pre-zeroed)
    const uint32_t full = len / 32;
    for (uint32_t b = 0; b < full; ++b) {
        const unsigned char* p = s + (size_t)b * 32;
        uint32_t w0 = 0, w1 = 0, v = 0;
        for (int j = 0; j < 16; ++j) {
            w0 = (w0 << 2) | g_pack.code[p[j]];
            w1 = (w1 << 2) | g_pack.code[p[16 + j]];
            v |= (uint32_t)g_pack.ok[p[j]] << j;
            v |= (uint32_t)g_pack.ok[p[16 + j]] << (16 + j);
        }
        codes[2 * (size_t)b] = w0;
        codes[2 * (size_t)b + 1] = w1;
        valid[b] = v;
    }
    const uint32_t rem = len - full * 32;
    uint32_t w0 = 0, w1 = 0, v = 0;
    const unsigned char* p = s + (size_t)full * 32;
    for (uint32_t j = 0; j < rem; ++j) {
        const uint32_t code = g_pack.code[p[j]];
        if (j < 16) w0 |= code << (30 - 2 * j); else w1 |= code << (62 - 2 * j);
        v |= (uint32_t)g_pack.ok[p[j]] << j;
    }
    codes[2 * (size_t)full] = w0;
    codes[2 * (size_t)full + 1] = w1;
    valid[full] = v;
}

static void free_reads(lrb_reads* r) {
    if (!r) return;
    lrb_host_free(r->codes, r->pinned);
    lrb_host_free(r->valid, r->pinned);
    free(r->read_len);
    free(r->read_blk);
    free(r->tile_read);
    free(r->tile_blk);
    lrb_host_free(r->exc_blk, r->exc_pinned);
    lrb_host_free(r->exc_valid, r->exc_pinned);
    delete r;
}

// builds read_blk / tiles from read_len and allocates zeroed codes/valid
static int build_layout(lrb_reads* r) {
    const uint64_t n = r->n_reads;
    r->read_blk = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
    if (!r->read_blk) return lrb_set_error(LRB_ENOMEM, "out of memory (read_blk)");
    uint64_t blk = 0, tiles = 0, bases = 0;
    for (uint64_t i = 0; i < n; ++i) {
        r->read_blk[i] = (uint32_t)blk;
        const uint64_t nb = (uint64_t)r->read_len[i] / 32 + 1;
        blk += nb;
        tiles += (nb + LRB_TILE_BLOCKS - 1) / LRB_TILE_BLOCKS;
        bases += r->read_len[i];
        if (blk > 0xFFFFFFF0ull) return lrb_set_error(LRB_EINVAL, "read set too large: more than 2^32 blocks (137 Gbases)");
    }
    r->read_blk[n] = (uint32_t)blk;
    r->n_blocks = blk;
    r->n_tiles = tiles;
    r->total_bases = bases;
    r->tile_read = (uint32_t*)malloc(sizeof(uint32_t) * (tiles + 1));
    r->tile_blk = (uint32_t*)malloc(sizeof(uint32_t) * (tiles + 1));
    if (!r->tile_read || !r->tile_blk) return lrb_set_error(LRB_ENOMEM, "out of memory (tiles)");
    uint64_t t = 0;
    for (uint64_t i = 0; i < n; ++i) {
        for (uint32_t b = r->read_blk[i]; b < r->read_blk[i + 1]; b += LRB_TILE_BLOCKS) {
            r->tile_read[t] = (uint32_t)i;
            r->tile_blk[t] = b;
            ++t;
        }
    }
    r->codes = (uint32_t*)lrb_host_alloc(sizeof(uint32_t) * (2 * blk + 2), &r->pinned);
    bool pinned2 = r->pinned;
    r->valid = r->codes ? (uint32_t*)lrb_host_alloc(sizeof(uint32_t) * (blk + 1), &pinned2) : nullptr;
    if (!r->codes || !r->valid) return lrb_set_error(LRB_ENOMEM, "out of memory (packed stream, %llu blocks)", (unsigned long long)blk);
    if (pinned2 != r->pinned) {  // keep one allocation kind for both
        lrb_host_free(r->valid, pinned2);
        lrb_host_free(r->codes, r->pinned);
        r->pinned = false;
        r->codes = (uint32_t*)calloc(2 * blk + 2, 4);
        r->valid = (uint32_t*)calloc(blk + 1, 4);
        if (!r->codes || !r->valid) return lrb_set_error(LRB_ENOMEM, "out of memory (packed stream)");
    } else {
        memset(r->codes, 0, sizeof(uint32_t) * (2 * blk + 2));
        memset(r->valid, 0, sizeof(uint32_t) * (blk + 1));
    }
    return LRB_OK;
}

template <class GetSeq>
static void pack_all(lrb_reads* r, int threads, GetSeq get) {
    const uint64_t n = r->n_reads;
    threads = std::max(1, std::min(threads, 64));
    if (n < 64) threads = 1;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i)
            pack_read(get(i), r->read_len[i], r->codes + 2 * (size_t)r->read_blk[i], r->valid + r->read_blk[i]);
    };
    if (threads == 1) { work(0, n); return; }
    // split by blocks so threads get even byte counts
    std::vector<std::thread> pool;
    uint64_t lo = 0;
    for (int t = 0; t < threads; ++t) {
        const uint64_t target = r->n_blocks * (uint64_t)(t + 1) / threads;
        uint64_t hi = (t == threads - 1) ? n : (uint64_t)(std::upper_bound(r->read_blk, r->read_blk + n, (uint32_t)target) - r->read_blk);
        if (hi < lo) hi = lo;
        if (hi > n) hi = n;
        pool.emplace_back(work, lo, hi);
        lo = hi;
    }
    for (auto& th : pool) th.join();
}

// exception list = blocks whose valid word differs from default_valid (threads scan disjoint read ranges, in order)
static int index_valid(lrb_reads* r, int threads) {
    lrb_host_free(r->exc_blk, r->exc_pinned);
    lrb_host_free(r->exc_valid, r->exc_pinned);
    r->exc_blk = r->exc_valid = nullptr;
    r->n_exc = 0;
    r->exc_ready = false;
    const uint64_t n = r->n_reads;
    threads = std::max(1, std::min(threads, 64));
    if (n < 64) threads = 1;
    std::vector<std::vector<uint32_t>> found(threads);  // (block, word) pairs
    auto work = [&](int t, uint64_t lo, uint64_t hi) {
        std::vector<uint32_t>& f = found[t];
        for (uint64_t i = lo; i < hi; ++i) {
            const uint32_t b0 = r->read_blk[i], b1 = r->read_blk[i + 1], len = r->read_len[i];
            for (uint32_t b = b0; b < b1; ++b) {
                const uint32_t v = r->valid[b];
                if (v != lrb::default_valid_word(len, (uint64_t)(b - b0) * 32)) { f.push_back(b); f.push_back(v); }
            }
        }
    };
    if (threads == 1) work(0, 0, n);
    else {
        std::vector<std::thread> pool;
        uint64_t lo = 0;
        for (int t = 0; t < threads; ++t) {
            const uint64_t target = r->n_blocks * (uint64_t)(t + 1) / threads;
            uint64_t hi = (t == threads - 1) ? n : (uint64_t)(std::upper_bound(r->read_blk, r->read_blk + n, (uint32_t)target) - r->read_blk);
            hi = std::min(std::max(hi, lo), n);
            pool.emplace_back(work, t, lo, hi);
            lo = hi;
        }
        for (auto& th : pool) th.join();
    }
    uint64_t total = 0;
    for (auto& f : found) total += f.size() / 2;
    bool p1 = false, p2 = false;
    r->exc_blk = (uint32_t*)lrb_host_alloc(sizeof(uint32_t) * (total + 1), &p1);
    r->exc_valid = (uint32_t*)lrb_host_alloc(sizeof(uint32_t) * (total + 1), &p2);
    if (!r->exc_blk || !r->exc_valid || p1 != p2) {
        lrb_host_free(r->exc_blk, p1);
        lrb_host_free(r->exc_valid, p2);
        r->exc_blk = r->exc_valid = nullptr;
        return lrb_set_error(LRB_ENOMEM, "out of memory (validity exceptions)");
    }
    r->exc_pinned = p1;
    uint64_t k = 0;
    for (auto& f : found)
        for (size_t j = 0; j < f.size(); j += 2) { r->exc_blk[k] = f[j]; r->exc_valid[k] = f[j + 1]; ++k; }
    r->n_exc = total;
    r->exc_ready = true;
    return LRB_OK;
}

}  // namespace

extern "C" int lrb_reads_index_valid(lrb_reads* r, int threads, uint64_t* n_exceptions) {
    if (!r) return lrb_set_error(LRB_EINVAL, "lrb_reads_index_valid: null argument");
    const int rc = index_valid(r, threads);
    if (!rc && n_exceptions) *n_exceptions = r->n_exc;
    return rc;
}

extern "C" int lrb_reads_exceptions(const lrb_reads* r, const uint32_t** blk, const uint32_t** word, uint64_t* n) {
    if (!r || !blk || !word || !n) return lrb_set_error(LRB_EINVAL, "lrb_reads_exceptions: null argument");
    *blk = r->exc_ready ? r->exc_blk : nullptr;
    *word = r->exc_ready ? r->exc_valid : nullptr;
    *n = r->exc_ready ? r->n_exc : 0;
    return LRB_OK;
}

extern "C" int lrb_reads_from_file(const char* path, int threads, lrb_reads** out) {
    if (!path || !out) return lrb_set_error(LRB_EINVAL, "lrb_reads_from_file: null argument");
    *out = nullptr;
    std::vector<unsigned char> data;
    read_whole_file(path, data);  // unreadable file == empty stream (tools: exit 0, empty outputs)
    Parsed parsed;
    parse_records(data.data(), data.size(), parsed);
    std::vector<unsigned char>().swap(data);
    lrb_reads* r = new lrb_reads();
    r->n_reads = parsed.recs.size();
    r->read_len = (uint32_t*)malloc(sizeof(uint32_t) * (r->n_reads + 1));
    if (!r->read_len) { free_reads(r); return lrb_set_error(LRB_ENOMEM, "out of memory (read_len)"); }
    for (uint64_t i = 0; i < r->n_reads; ++i) r->read_len[i] = parsed.recs[i].len;
    int rc = build_layout(r);
    if (rc) { free_reads(r); return rc; }
    const unsigned char* pool = (const unsigned char*)parsed.pool.data();
    pack_all(r, threads, [&](uint64_t i) { return pool + parsed.recs[i].off; });
    if ((rc = index_valid(r, threads))) { free_reads(r); return rc; }
    *out = r;
    return LRB_OK;
}

extern "C" int lrb_reads_from_ascii(const char* bases, const uint64_t* offsets, uint64_t n_reads, int threads,
                                    lrb_reads** out) {
    if (!out || (n_reads && (!bases || !offsets))) return lrb_set_error(LRB_EINVAL, "lrb_reads_from_ascii: null argument");
    *out = nullptr;
    lrb_reads* r = new lrb_reads();
    r->n_reads = n_reads;
    r->read_len = (uint32_t*)malloc(sizeof(uint32_t) * (n_reads + 1));
    if (!r->read_len) { free_reads(r); return lrb_set_error(LRB_ENOMEM, "out of memory (read_len)"); }
    for (uint64_t i = 0; i < n_reads; ++i) {
        const uint64_t l = offsets[i + 1] - offsets[i];
        if (offsets[i + 1] < offsets[i] || l > 0xFFFFFFFFull) { free_reads(r); return lrb_set_error(LRB_EINVAL, "lrb_reads_from_ascii: bad offsets at read %llu", (unsigned long long)i); }
        r->read_len[i] = (uint32_t)l;
    }
    int rc = build_layout(r);
    if (rc) { free_reads(r); return rc; }
    pack_all(r, threads, [&](uint64_t i) { return (const unsigned char*)bases + offsets[i]; });
    if ((rc = index_valid(r, threads))) { free_reads(r); return rc; }
    *out = r;
    return LRB_OK;
}

extern "C" int lrb_reads_from_lengths(const uint32_t* lengths, uint64_t n_reads, lrb_reads** out) {
    if (!out || (n_reads && !lengths)) return lrb_set_error(LRB_EINVAL, "lrb_reads_from_lengths: null argument");
    *out = nullptr;
    lrb_reads* r = new lrb_reads();
    r->n_reads = n_reads;
    r->read_len = (uint32_t*)malloc(sizeof(uint32_t) * (n_reads + 1));
    if (!r->read_len) { free_reads(r); return lrb_set_error(LRB_ENOMEM, "out of memory (read_len)"); }
    if (n_reads) memcpy(r->read_len, lengths, sizeof(uint32_t) * n_reads);
    int rc = build_layout(r);
    if (rc) { free_reads(r); return rc; }
    *out = r;
    return LRB_OK;
}

extern "C" int lrb_reads_view_get(const lrb_reads* r, lrb_reads_view* v) {
    if (!r || !v) return lrb_set_error(LRB_EINVAL, "lrb_reads_view_get: null argument");
    v->n_reads = r->n_reads;
    v->n_blocks = r->n_blocks;
    v->n_tiles = r->n_tiles;
    v->total_bases = r->total_bases;
    v->codes = r->codes;
    v->valid = r->valid;
    v->read_len = r->read_len;
    v->read_blk = r->read_blk;
    v->tile_read = r->tile_read;
    v->tile_blk = r->tile_blk;
    return LRB_OK;
}

extern "C" int lrb_reads_unpack(const lrb_reads* r, uint64_t i, char* dst, uint64_t cap) {
    if (!r || !dst || i >= r->n_reads) return lrb_set_error(LRB_EINVAL, "lrb_reads_unpack: bad argument");
    const uint32_t len = r->read_len[i];
    if (cap < len) return lrb_set_error(LRB_EINVAL, "lrb_reads_unpack: buffer too small");
    static const char letters[4] = {'A', 'C', 'T', 'G'};
    const uint32_t* codes = r->codes + 2 * (size_t)r->read_blk[i];
    for (uint32_t p = 0; p < len; ++p) dst[p] = letters[(codes[p >> 4] >> (30 - 2 * (p & 15))) & 3u];
    return LRB_OK;
}

extern "C" void lrb_reads_free(lrb_reads* r) { free_reads(r); }
