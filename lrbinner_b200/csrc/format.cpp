// format.cpp — host epilogue: exact "%f" text rows, .npy arrays, the 15mers-counts table file,
// plus the small host utilities shared by the library (error string, canonical k-mer LUT).
//
// Replaces the serial std::to_string loops of the tools (count-kmers.cpp:110-118,
// search-15mers.cpp:35-48) and writeKmerFile/readKmerFile (kmer_utils.h:89-112).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.h"
#include "fixed6.h"

// ---- error plumbing ----------------------------------------------------------------------------
static thread_local char t_err[512] = "";

int lrb_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* lrb_last_error(void) { return t_err; }
extern "C" int lrb_version(void) { return 100; }

// ---- canonical k-mer index (compute_kmer_inds, count-kmers.cpp:38-64) ---------------------------
// index = rank of min(x, rc(x)) among the canonical values in ascending order; a k-mer met after its
// reverse complement shares that one's index.
extern "C" int lrb_kmer_lut(int k, uint16_t* lut) {
    if (k < 1 || k > 5 || !lut) return -1;
    const uint32_t n = 1u << (2 * k);
    uint32_t next = 0;
    for (uint32_t x = 0; x < n; ++x) {
        uint32_t rc = 0, t = x;
        for (int i = 0; i < k; ++i) { rc = (rc << 2) | ((t & 3u) ^ 2u); t >>= 2; }
        lut[x] = (rc < x) ? lut[rc] : (uint16_t)next++;
    }
    return (int)next;
}

// ---- "%f" of count/total -----------------------------------------------------------------------
extern "C" uint32_t lrb_fixed6(uint32_t num, uint32_t den, int coverage) { return lrb::fixed6(num, den, coverage != 0); }

namespace {

// two decimal digits at a time
struct Digits2 {
    char d[100][2];
    Digits2() { for (int i = 0; i < 100; ++i) { d[i][0] = (char)('0' + i / 10); d[i][1] = (char)('0' + i % 10); } }
};
const Digits2 kDigits2;

inline void put_fixed6(char* dst, uint32_t q) {  // 8 chars "d.dddddd"
    if (q == 0) { memcpy(dst, "0.000000", 8); return; }   // most composition / coverage cells
    const uint32_t ip = q / 1000000u, f = q - ip * 1000000u;
    const uint32_t a = f / 10000u, r = f - a * 10000u, b2 = r / 100u, c2 = r - b2 * 100u;
    dst[0] = (char)('0' + ip);
    dst[1] = '.';
    memcpy(dst + 2, kDigits2.d[a], 2);
    memcpy(dst + 4, kDigits2.d[b2], 2);
    memcpy(dst + 6, kDigits2.d[c2], 2);
}

int comp_width(int k) { return k == 3 ? 32 : k == 4 ? 136 : k == 5 ? 512 : 0; }

bool pwrite_all(int fd, const char* p, size_t n, uint64_t off) {
    while (n) {
        const ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w <= 0) return false;
        p += w;
        n -= (size_t)w;
        off += (uint64_t)w;
    }
    return true;
}

// Fixed-width rows: every thread formats blocks of its share of the rows into a private buffer and writes them at their
// final offset (pwrite), so neither the text nor the file copy is ever serial and nothing of file size sits in memory.
// row(i, dst) fills row i (row_bytes bytes).
template <class RowFn>
int write_fixed_rows(const char* path, const void* head, size_t head_bytes, uint64_t n_rows, size_t row_bytes, int threads, RowFn row) {
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return lrb_set_error(LRB_EIO, "cannot open %s for writing", path);
    bool ok = !head_bytes || pwrite_all(fd, (const char*)head, head_bytes, 0);
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n_rows < 256) threads = 1;
    const uint64_t block_rows = std::max<uint64_t>(1, (2ull << 20) / std::max<size_t>(row_bytes, 1));
    std::vector<char> failed((size_t)threads, 0);
    auto work = [&](int t) {
        const uint64_t a = n_rows * (uint64_t)t / (uint64_t)threads, b = n_rows * (uint64_t)(t + 1) / (uint64_t)threads;
        std::vector<char> buf((size_t)std::min<uint64_t>(block_rows, std::max<uint64_t>(b - a, 1)) * row_bytes);
        for (uint64_t lo = a; lo < b; lo += block_rows) {
            const uint64_t hi = std::min(b, lo + block_rows);
            for (uint64_t i = lo; i < hi; ++i) row(i, buf.data() + (size_t)(i - lo) * row_bytes);
            if (!pwrite_all(fd, buf.data(), (size_t)(hi - lo) * row_bytes, head_bytes + lo * row_bytes)) { failed[(size_t)t] = 1; return; }
        }
    };
    if (ok && n_rows) {
        if (threads == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
            for (auto& th : pool) th.join();
        }
        for (char f : failed) ok = ok && !f;
    }
    if (close(fd) != 0) ok = false;
    if (!ok) return lrb_set_error(LRB_EIO, "short write to %s", path);
    return LRB_OK;
}

template <class RowFn>
int write_rows(const char* path, uint64_t n_rows, size_t row_bytes, int threads, RowFn row) {
    return write_fixed_rows(path, nullptr, 0, n_rows, row_bytes, threads, row);
}

std::string npy_header(uint64_t rows, uint64_t cols) {
    char dict[160];
    snprintf(dict, sizeof dict, "{'descr': '<f8', 'fortran_order': False, 'shape': (%llu, %llu), }",
             (unsigned long long)rows, (unsigned long long)cols);
    std::string h("\x93NUMPY\x01\x00", 8);
    size_t len = strlen(dict) + 1;              // + '\n'
    size_t total = 10 + len;
    size_t pad = (64 - total % 64) % 64;
    uint16_t hlen = (uint16_t)(len + pad);
    h.push_back((char)(hlen & 0xFF));
    h.push_back((char)(hlen >> 8));
    h += dict;
    h.append(pad, ' ');
    h.push_back('\n');
    return h;
}

template <class RowFn>
int write_npy(const char* path, uint64_t n_rows, uint64_t cols, int threads, RowFn row) {
    const std::string h = npy_header(n_rows, cols);
    return write_fixed_rows(path, h.data(), h.size(), n_rows, (size_t)cols * 8, threads,
                            [&](uint64_t i, char* dst) { row(i, reinterpret_cast<double*>(dst)); });
}

inline uint32_t comp_total(uint32_t len, int k) { return len >= (uint32_t)k ? len - (uint32_t)k + 1u : 0u; }

}  // namespace

extern "C" int lrb_write_composition_txt(const char* path, const uint32_t* counts, const uint32_t* read_len,
                                         uint64_t n_reads, int k, int threads) {
    const int P = comp_width(k);
    if (!path || !P || (n_reads && (!counts || !read_len))) return lrb_set_error(LRB_EINVAL, "lrb_write_composition_txt: bad argument");
    const size_t row_bytes = (size_t)P * 9 + 1;  // "d.dddddd " x P, then '\n'  (count-kmers.cpp:110-118)
    return write_rows(path, n_reads, row_bytes, threads, [=](uint64_t i, char* dst) {
        const uint32_t total = comp_total(read_len[i], k);
        const uint32_t* c = counts + (size_t)i * P;
        for (int j = 0; j < P; ++j) {
            put_fixed6(dst + (size_t)j * 9, lrb::fixed6(c[j], total, false));
            dst[(size_t)j * 9 + 8] = ' ';
        }
        dst[row_bytes - 1] = '\n';
    });
}

extern "C" int lrb_write_coverage_txt(const char* path, const uint32_t* hist, const uint32_t* sums, uint64_t n_reads,
                                      int bins, int threads) {
    if (!path || bins <= 0 || (n_reads && (!hist || !sums))) return lrb_set_error(LRB_EINVAL, "lrb_write_coverage_txt: bad argument");
    const size_t row_bytes = (size_t)bins * 9;  // values separated by ' ', '\n' after the last (search-15mers.cpp:35-48)
    return write_rows(path, n_reads, row_bytes, threads, [=](uint64_t i, char* dst) {
        const uint32_t* c = hist + (size_t)i * bins;
        for (int j = 0; j < bins; ++j) {
            put_fixed6(dst + (size_t)j * 9, lrb::fixed6(c[j], sums[i], true));
            dst[(size_t)j * 9 + 8] = (j == bins - 1) ? '\n' : ' ';
        }
    });
}

extern "C" int lrb_write_composition_npy(const char* path, const uint32_t* counts, const uint32_t* read_len,
                                         uint64_t n_reads, int k, int threads) {
    const int P = comp_width(k);
    if (!path || !P || (n_reads && (!counts || !read_len))) return lrb_set_error(LRB_EINVAL, "lrb_write_composition_npy: bad argument");
    return write_npy(path, n_reads, (uint64_t)P, threads, [=](uint64_t i, double* dst) {
        const uint32_t total = comp_total(read_len[i], k);
        const uint32_t* c = counts + (size_t)i * P;
        for (int j = 0; j < P; ++j) dst[j] = (double)lrb::fixed6(c[j], total, false) / 1e6;  // == float("%f" text)
    });
}

extern "C" int lrb_write_coverage_npy(const char* path, const uint32_t* hist, const uint32_t* sums, uint64_t n_reads,
                                      int bins, int threads) {
    if (!path || bins <= 0 || (n_reads && (!hist || !sums))) return lrb_set_error(LRB_EINVAL, "lrb_write_coverage_npy: bad argument");
    return write_npy(path, n_reads, (uint64_t)bins, threads, [=](uint64_t i, double* dst) {
        const uint32_t* c = hist + (size_t)i * bins;
        for (int j = 0; j < bins; ++j) dst[j] = (double)lrb::fixed6(c[j], sums[i], true) / 1e6;
    });
}

// ---- 15mers-counts file (kmer_utils.h:89-112): u64 size (= 2^30), then size little-endian u32 -----
extern "C" int lrb_table_write_file(const char* path, const uint32_t* table) {
    if (!path || !table) return lrb_set_error(LRB_EINVAL, "lrb_table_write_file: null argument");
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return lrb_set_error(LRB_EIO, "cannot open %s for writing", path);
    const uint64_t size = LRB_TABLE_ENTRIES;
    if (write(fd, &size, sizeof size) != (ssize_t)sizeof size) { close(fd); return lrb_set_error(LRB_EIO, "short write to %s", path); }
    const char* p = (const char*)table;
    size_t left = (size_t)size * 4;
    while (left) {
        const ssize_t w = write(fd, p, std::min<size_t>(left, 1u << 30));
        if (w <= 0) { close(fd); return lrb_set_error(LRB_EIO, "short write to %s", path); }
        p += w;
        left -= (size_t)w;
    }
    if (close(fd) != 0) return lrb_set_error(LRB_EIO, "close failed for %s", path);
    return LRB_OK;
}

extern "C" int lrb_table_read_file(const char* path, uint32_t* table) {
    if (!path || !table) return lrb_set_error(LRB_EINVAL, "lrb_table_read_file: null argument");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return lrb_set_error(LRB_EIO, "cannot open table file %s", path);
    uint64_t size = 0;
    if (read(fd, &size, sizeof size) != (ssize_t)sizeof size || size != LRB_TABLE_ENTRIES) {
        close(fd);
        return lrb_set_error(LRB_EFORMAT, "%s is not a 4^15-entry 15mers-counts file", path);
    }
    char* p = (char*)table;
    size_t left = (size_t)size * 4;
    while (left) {
        const ssize_t r = read(fd, p, std::min<size_t>(left, 1u << 30));
        if (r <= 0) { close(fd); return lrb_set_error(LRB_EFORMAT, "%s is truncated", path); }
        p += r;
        left -= (size_t)r;
    }
    close(fd);
    return LRB_OK;
}
