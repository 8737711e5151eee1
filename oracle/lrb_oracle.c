/*
 * lrb_oracle.c — plain-C restatement of LRBinner's profile stage (see lrb_oracle.h).
 * TEST INFRASTRUCTURE ONLY: scalar, single-threaded, written for obviousness, not speed.
 * Parity: PINNED against the reference tools in oracle/_ref (tests/test_oracle_pins.py, tests/golden/).
 */
#define _GNU_SOURCE
#include "lrb_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>
#include <sys/mman.h>

#define ORC_TABLE_SIZE (1ull << 30) /* 4^15, count-15mers.cpp:99 */
#define ORC_MASK15 1073741823ull    /* kmer_utils.h:46,130 */

/* ---- k-mer arithmetic ---------------------------------------------------------------------- */

/* kmer_utils.h:10-22: reverse the 32 base pairs of a u64, complement (XOR 0b10 per base), and shift
 * the k bases that matter back down.  Stated here base by base. */
uint64_t orc_revcomp(uint64_t x, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; ++i) {
        uint64_t base = (x >> (2 * i)) & 3u; /* i-th base from the end */
        r = (r << 2) | (base ^ 2u);
    }
    return r;
}

/* count-kmers.cpp:38-64: walk k-mers in ascending order; a k-mer whose reverse complement already has
 * an index shares it, otherwise it takes the next free index. */
int orc_kmer_lut(int k, uint32_t* lut) {
    uint32_t n = 1u << (2 * k), next = 0;
    uint8_t* seen = (uint8_t*)calloc(n, 1);
    for (uint32_t kmer = 0; kmer < n; ++kmer) {
        uint32_t rc = (uint32_t)orc_revcomp(kmer, k);
        if (seen[rc]) {
            lut[kmer] = lut[rc];
        } else {
            lut[kmer] = next++;
        }
        seen[kmer] = 1;
    }
    free(seen);
    return (int)next;
}

/* count-kmers.cpp:66-95 */
void orc_composition(const char* seq, size_t len, int k, const uint32_t* lut, int width,
                     uint64_t* raw, uint64_t* total_out, double* profile) {
    uint64_t mask = (1ull << (2 * k)) - 1, val = 0, total = 0;
    long run = 0;
    for (int i = 0; i < width; ++i) raw[i] = 0;
    for (size_t i = 0; i < len; ++i) {
        val = (val << 2) & mask;
        val += (uint64_t)((seq[i] >> 1) & 3); /* :77 — every byte contributes, no ACGT check */
        run++;
        if (run == k) {
            run--;
            raw[lut[val]]++;
            total++;
        }
    }
    if (total_out) *total_out = total;
    if (profile) {
        double denom = total > 1 ? (double)total : 1.0; /* :91 max(1.0, total) */
        for (int i = 0; i < width; ++i) profile[i] = (double)raw[i] / denom;
    }
}

static int orc_is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

/* kmer_utils.h:114-156, with the k-mer length as a parameter (the reference fixes k = 15, mask 4^15-1;
 * other odd k are only used by tests of the multi-GPU plumbing, where a 4 GiB table would be wasteful). */
void orc_count_kmers_k(const char* seq, size_t len, int k, uint32_t* table) {
    const uint64_t mask = (1ull << (2 * k)) - 1;
    uint64_t val = 0;
    long run = 0;
    for (size_t i = 0; i < len; ++i) {
        if (!orc_is_acgt(seq[i])) { /* :122-127 */
            val = 0;
            run = 0;
            continue;
        }
        val = (val << 2) & mask;
        val += (uint64_t)((seq[i] >> 1) & 3);
        run++;
        if (run == k) {
            run--;
            table[val] += 1u;                   /* :139-145 CAS loop == +1 mod 2^32 */
            table[orc_revcomp(val, k)] += 1u;   /* :148-153 */
        }
    }
}

void orc_count_15mers(const char* seq, size_t len, uint32_t* table) { orc_count_kmers_k(seq, len, 15, table); }

/* The forward key of every valid window, in read order: the `val` the loops at kmer_utils.h:36-72 and
 * :120-155 use as table index.  Returns the number of windows (out may be NULL to just count). */
size_t orc_window_keys(const char* seq, size_t len, int k, uint32_t* out) {
    const uint64_t mask = (1ull << (2 * k)) - 1;
    uint64_t val = 0;
    long run = 0;
    size_t n = 0;
    for (size_t i = 0; i < len; ++i) {
        if (!orc_is_acgt(seq[i])) {
            val = 0;
            run = 0;
            continue;
        }
        val = (val << 2) & mask;
        val += (uint64_t)((seq[i] >> 1) & 3);
        run++;
        if (run == k) {
            run--;
            if (out) out[n] = (uint32_t)val;
            n++;
        }
    }
    return n;
}

/* kmer_utils.h:54-69 */
int orc_bucket(uint32_t count_u32, long bin_size, int bins) {
    long count = (long)count_u32;
    count = count < 2 ? 0 : count;
    long pos = (count / bin_size) - 1;
    if (count <= bin_size) return 0;
    if (pos < bins && pos > 0) return (int)pos;
    return bins - 1;
}

/* kmer_utils.h:24-87, k-mer length as a parameter (reference: 15) */
void orc_coverage_k(const char* seq, size_t len, int k, const uint32_t* table, long bin_size, int bins,
                    uint64_t* raw, uint64_t* sum_out, double* vec) {
    const uint64_t mask = (1ull << (2 * k)) - 1;
    uint64_t val = 0, sum = 0;
    long run = 0;
    for (int i = 0; i < bins; ++i) raw[i] = 0;
    for (size_t i = 0; i < len; ++i) {
        if (!orc_is_acgt(seq[i])) {
            val = 0;
            run = 0;
            continue;
        }
        val = (val << 2) & mask;
        val += (uint64_t)((seq[i] >> 1) & 3);
        run++;
        if (run == k) {
            run--;
            raw[orc_bucket(table[val], bin_size, bins)]++;
            sum++;
        }
    }
    if (sum_out) *sum_out = sum;
    if (vec) {
        for (int i = 0; i < bins; ++i) vec[i] = (double)raw[i];
        if (sum > 0) {
            for (int i = 0; i < bins; ++i) {
                vec[i] /= (double)(long)sum; /* :79 double / long */
                if (vec[i] < 1e-4) vec[i] = 0; /* :80-83 */
            }
        }
    }
}

void orc_coverage(const char* seq, size_t len, const uint32_t* table, long bin_size, int bins,
                  uint64_t* raw, uint64_t* sum_out, double* vec) {
    orc_coverage_k(seq, len, 15, table, bin_size, bins, raw, sum_out, vec);
}

/* ---- record stream (kseq.h restated over an in-memory byte array) ---------------------------- */

typedef struct {
    const unsigned char* p;
    size_t n, pos;
} orc_stream;

typedef struct {
    char* s;
    size_t l, m;
} orc_str;

static void str_reserve(orc_str* s, size_t need) {
    if (need + 1 > s->m) {
        size_t m = s->m ? s->m : 256;
        while (m < need + 1) m *= 2;
        s->s = (char*)realloc(s->s, m);
        s->m = m;
    }
}
static void str_push(orc_str* s, const unsigned char* src, size_t n) {
    str_reserve(s, s->l + n);
    memcpy(s->s + s->l, src, n);
    s->l += n;
    s->s[s->l] = 0;
}

static int st_getc(orc_stream* st) { return st->pos < st->n ? (int)st->p[st->pos++] : -1; }

/* kseq.h:93-144 (ks_getuntil2).  sep: 0 = any isspace(), 2 = '\n'.  Returns -1 when called at end of
 * data (nothing consumed, no CR strip), else the string length. *dret = delimiter hit or 0. */
static long st_getuntil(orc_stream* st, int sep, orc_str* str, int* dret, int append) {
    if (dret) *dret = 0;
    if (!append) str->l = 0;
    if (st->pos >= st->n) return -1; /* :139 !gotany && eof */
    size_t i = st->pos;
    if (sep == 2) {
        while (i < st->n && st->p[i] != '\n') ++i;
    } else {
        while (i < st->n && !isspace(st->p[i])) ++i;
    }
    str_push(str, st->p + st->pos, i - st->pos);
    if (i < st->n) {
        if (dret) *dret = st->p[i];
        st->pos = i + 1;
    } else {
        st->pos = i;
    }
    if (sep == 2 && str->l > 1 && str->s[str->l - 1] == '\r') { /* :143 */
        str->l--;
        str->s[str->l] = 0;
    }
    return (long)str->l;
}

typedef struct {
    orc_stream st;
    int last_char;
    orc_str name, comment, seq, qual;
} orc_kseq;

/* kseq.h:177-218 (kseq_read): >=0 length, -1 EOF, -2 bad quality */
static long orc_kseq_read(orc_kseq* ks) {
    int c;
    long r;
    if (ks->last_char == 0) { /* :181-185 jump to the next header char, wherever it is */
        while ((c = st_getc(&ks->st)) >= 0 && c != '>' && c != '@') {
        }
        if (c < 0) return c;
        ks->last_char = c;
    }
    ks->comment.l = ks->seq.l = ks->qual.l = 0;
    str_reserve(&ks->seq, 0);
    ks->seq.s[0] = 0;
    if ((r = st_getuntil(&ks->st, 0, &ks->name, &c, 0)) < 0) return r; /* :187 */
    if (c != '\n') st_getuntil(&ks->st, 2, &ks->comment, 0, 0);        /* :188 */
    while ((c = st_getc(&ks->st)) >= 0 && c != '>' && c != '+' && c != '@') { /* :193 */
        if (c == '\n') continue;
        unsigned char ch = (unsigned char)c;
        str_push(&ks->seq, &ch, 1);
        st_getuntil(&ks->st, 2, &ks->seq, 0, 1); /* rest of the line, appended */
    }
    if (c == '>' || c == '@') ks->last_char = c; /* :198 */
    if (c != '+') return (long)ks->seq.l;        /* FASTA */
    while ((c = st_getc(&ks->st)) >= 0 && c != '\n') { /* :209 rest of the '+' line */
    }
    if (c == -1) return -2;
    /* :211 — keep appending quality lines while the call succeeds and qual is shorter than seq */
    while (st_getuntil(&ks->st, 2, &ks->qual, 0, 1) >= 0 && ks->qual.l < ks->seq.l) {
    }
    ks->last_char = 0;
    if (ks->seq.l != ks->qual.l) return -2; /* :214 */
    return (long)ks->seq.l;
}

int orc_reads_parse(const unsigned char* buf, size_t n, orc_reads_t* out) {
    orc_kseq ks;
    memset(&ks, 0, sizeof ks);
    ks.st.p = buf;
    ks.st.n = n;
    size_t cap = 1024;
    out->n = 0;
    out->seq = (char**)malloc(cap * sizeof(char*));
    out->name = (char**)malloc(cap * sizeof(char*));
    out->len = (size_t*)malloc(cap * sizeof(size_t));
    while (orc_kseq_read(&ks) >= 0) { /* io_utils.h:153-164 */
        if (out->n == cap) {
            cap *= 2;
            out->seq = (char**)realloc(out->seq, cap * sizeof(char*));
            out->name = (char**)realloc(out->name, cap * sizeof(char*));
            out->len = (size_t*)realloc(out->len, cap * sizeof(size_t));
        }
        size_t l = strlen(ks.seq.s); /* string(ks->seq.s): cut at first NUL */
        out->seq[out->n] = (char*)malloc(l + 1);
        memcpy(out->seq[out->n], ks.seq.s, l + 1);
        out->len[out->n] = l;
        out->name[out->n] = strdup(ks.name.s ? ks.name.s : "");
        out->n++;
    }
    free(ks.name.s);
    free(ks.comment.s);
    free(ks.seq.s);
    free(ks.qual.s);
    return 0;
}

int orc_reads_load(const char* path, orc_reads_t* out) {
    gzFile f = gzopen(path, "r"); /* io_utils.h:143 — transparent for plain files */
    size_t n = 0, cap = 1 << 20;
    unsigned char* buf = (unsigned char*)malloc(cap);
    if (f) {
        for (;;) {
            if (cap - n < (1 << 16)) {
                cap *= 2;
                buf = (unsigned char*)realloc(buf, cap);
            }
            int got = gzread(f, buf + n, (unsigned)(cap - n > (1u << 30) ? (1u << 30) : cap - n));
            if (got <= 0) break; /* EOF or stream error: the tools stop quietly either way */
            n += (size_t)got;
        }
        gzclose(f);
    }
    int rc = orc_reads_parse(buf, n, out);
    free(buf);
    return rc;
}

void orc_reads_free(orc_reads_t* r) {
    for (size_t i = 0; i < r->n; ++i) {
        free(r->seq[i]);
        free(r->name[i]);
    }
    free(r->seq);
    free(r->name);
    free(r->len);
    memset(r, 0, sizeof *r);
}

/* ---- table + file drivers -------------------------------------------------------------------- */

uint32_t* orc_table_alloc(void) {
    void* p = mmap(NULL, ORC_TABLE_SIZE * 4, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    return p == MAP_FAILED ? NULL : (uint32_t*)p;
}
void orc_table_free(uint32_t* t) {
    if (t) munmap(t, ORC_TABLE_SIZE * 4);
}

int orc_table_write(const char* path, const uint32_t* table, uint64_t size) { /* kmer_utils.h:89-97 */
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    fwrite(&size, sizeof size, 1, f);
    size_t done = 0, total = (size_t)size;
    while (done < total) {
        size_t chunk = total - done > (1u << 26) ? (1u << 26) : total - done;
        if (fwrite(table + done, 4, chunk, f) != chunk) {
            fclose(f);
            return 2;
        }
        done += chunk;
    }
    fclose(f);
    return 0;
}

uint32_t* orc_table_read(const char* path, uint64_t* size_out) { /* kmer_utils.h:99-112 */
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    uint64_t size = 0;
    if (fread(&size, sizeof size, 1, f) != 1 || size != ORC_TABLE_SIZE) {
        fclose(f);
        return NULL;
    }
    uint32_t* t = orc_table_alloc();
    size_t done = 0;
    while (done < size) {
        size_t chunk = size - done > (1u << 26) ? (1u << 26) : size - done;
        size_t got = fread(t + done, 4, chunk, f);
        if (got == 0) break;
        done += got;
    }
    fclose(f);
    if (size_out) *size_out = size;
    return t;
}

int orc_format_f(double v, char* dst) { return snprintf(dst, 32, "%f", v); }

int orc_count_kmers_file(const char* reads, const char* out_txt, int k) {
    uint32_t lut[1024];
    int width = orc_kmer_lut(k, lut);
    orc_reads_t rs;
    FILE* f = fopen(out_txt, "wb"); /* count-kmers.cpp:210 truncates first */
    if (!f) return 1;
    orc_reads_load(reads, &rs);
    uint64_t* raw = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)width);
    double* prof = (double*)malloc(sizeof(double) * (size_t)width);
    char tmp[32];
    for (size_t i = 0; i < rs.n; ++i) {
        orc_composition(rs.seq[i], rs.len[i], k, lut, width, raw, NULL, prof);
        for (int j = 0; j < width; ++j) { /* :110-118 every value is followed by a space */
            orc_format_f(prof[j], tmp);
            fputs(tmp, f);
            fputc(' ', f);
        }
        fputc('\n', f);
    }
    fclose(f);
    free(raw);
    free(prof);
    orc_reads_free(&rs);
    return 0;
}

int orc_count_15mers_file(const char* reads, const char* out_table) {
    uint32_t* t = orc_table_alloc();
    if (!t) return 1;
    orc_reads_t rs;
    orc_reads_load(reads, &rs);
    for (size_t i = 0; i < rs.n; ++i) orc_count_15mers(rs.seq[i], rs.len[i], t);
    int rc = orc_table_write(out_table, t, ORC_TABLE_SIZE);
    orc_table_free(t);
    orc_reads_free(&rs);
    return rc;
}

int orc_search_15mers_file(const char* table, const char* reads, const char* out_txt, long bin_size, int bins) {
    uint64_t size = 0;
    uint32_t* t = orc_table_read(table, &size);
    if (!t) return 1;
    FILE* f = fopen(out_txt, "wb");
    if (!f) {
        orc_table_free(t);
        return 1;
    }
    orc_reads_t rs;
    orc_reads_load(reads, &rs);
    uint64_t* raw = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)bins);
    double* vec = (double*)malloc(sizeof(double) * (size_t)bins);
    char tmp[32];
    for (size_t i = 0; i < rs.n; ++i) {
        orc_coverage(rs.seq[i], rs.len[i], t, bin_size, bins, raw, NULL, vec);
        for (int j = 0; j < bins; ++j) { /* search-15mers.cpp:35-48 single spaces, none trailing */
            orc_format_f(vec[j], tmp);
            fputs(tmp, f);
            if (j < bins - 1) fputc(' ', f);
        }
        fputc('\n', f);
    }
    fclose(f);
    free(raw);
    free(vec);
    orc_table_free(t);
    orc_reads_free(&rs);
    return 0;
}
