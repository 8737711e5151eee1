/*
 * lrb_oracle.h — CPU restatement of LRBinner's profile stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lrbinner_b200/ may include, link, import or execute this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.  Parity status: PINNED — tests/test_oracle_pins.py compares every
 * function here against the unmodified reference tools compiled into oracle/_ref (outputs committed
 * as fixtures under tests/golden/ by tests/golden/make_golden.py).
 *
 * Each function cites the reference lines (relative to /root/reference) it restates.
 */
#ifndef LRB_ORACLE_H
#define LRB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mbcclr_utils/kmer_utils.h:10-22 and count-kmers.cpp:24-36 — reverse complement of a k-mer packed
 * 2 bits/base with A=0,C=1,T=2,G=3 (complement = XOR 2), first base in the most significant pair. */
uint64_t orc_revcomp(uint64_t x, int k);

/* count-kmers.cpp:38-64 — kmer -> dense canonical index. lut has 4^k entries. Returns the width
 * (32/136/512 for k=3/4/5). */
int orc_kmer_lut(int k, uint32_t* lut);

/* count-kmers.cpp:66-95 — rolling k-mer over EVERY byte (no validity reset).  raw[width] gets the
 * integer counts, *total the number of windows, profile[width] (optional) raw/max(1,total) in double. */
void orc_composition(const char* seq, size_t len, int k, const uint32_t* lut, int width,
                     uint64_t* raw, uint64_t* total, double* profile);

/* kmer_utils.h:114-156 — rolling 15-mer; any byte outside uppercase ACGT resets the window; every
 * valid window increments table[val] and table[revcomp(val)] (u32 wrap-around). table has 4^15 entries. */
void orc_count_15mers(const char* seq, size_t len, uint32_t* table);

/* kmer_utils.h:24-87 — per-read coverage histogram.  raw[bins] gets the integer bucket counts,
 * *sum the number of valid windows, vec[bins] (optional) the normalised doubles after the <1e-4 rule. */
void orc_coverage(const char* seq, size_t len, const uint32_t* table, long bin_size, int bins,
                  uint64_t* raw, uint64_t* sum, double* vec);

/* The same two functions with the k-mer length as a parameter (k = 15 is the reference; other odd k serve
 * the CPU tests of the multi-GPU plumbing). table has 4^k entries. */
void orc_count_kmers_k(const char* seq, size_t len, int k, uint32_t* table);
void orc_coverage_k(const char* seq, size_t len, int k, const uint32_t* table, long bin_size, int bins,
                    uint64_t* raw, uint64_t* sum, double* vec);

size_t orc_window_keys(const char* seq, size_t len, int k, uint32_t* out); /* forward key of each valid window */

/* Bucket rule alone (kmer_utils.h:54-69): global count -> histogram bin. */
int orc_bucket(uint32_t count, long bin_size, int bins);

/* io_utils.h:133-165 + kseq.h:93-144,177-218 — FASTA/FASTQ (plain or gzip) record stream.
 * orc_reads_load parses the whole file; sequences are returned as they reach the tools
 * (std::string built from a C string: cut at the first NUL byte). */
typedef struct {
    size_t n;          /* number of records */
    char** seq;        /* seq[i] is NUL-terminated */
    size_t* len;       /* strlen(seq[i]) */
    char** name;
} orc_reads_t;
int orc_reads_load(const char* path, orc_reads_t* out);          /* 0 ok (a missing file gives 0 records, like the tools) */
int orc_reads_parse(const unsigned char* buf, size_t n, orc_reads_t* out);
void orc_reads_free(orc_reads_t* r);

/* File-level drivers with the tools' argv contracts (count-kmers.cpp:189-218, count-15mers.cpp:97-123,
 * search-15mers.cpp:121-157); byte-identical outputs. Return 0. */
int orc_count_kmers_file(const char* reads, const char* out_txt, int k);
int orc_count_15mers_file(const char* reads, const char* out_table);
int orc_search_15mers_file(const char* table, const char* reads, const char* out_txt, long bin_size, int bins);

/* kmer_utils.h:89-112 — table file: u64 size, then size u32. */
int orc_table_write(const char* path, const uint32_t* table, uint64_t size);
uint32_t* orc_table_read(const char* path, uint64_t* size);
uint32_t* orc_table_alloc(void);   /* zeroed 4^15 u32 (lazy pages) */
void orc_table_free(uint32_t* t);

/* "%f" rendering used by std::to_string(double) (count-kmers.cpp:112, search-15mers.cpp:39). */
int orc_format_f(double v, char* dst /* >= 32 bytes */);

#ifdef __cplusplus
}
#endif
#endif
