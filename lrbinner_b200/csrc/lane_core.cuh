// lane_core.cuh — per-lane bit arithmetic of the profile kernels, usable from host and device.
//
// Everything a CUDA lane does to one 32-slot block of the packed read stream is written here as
// LRB_HD functions so that tests/host_emul.cpp can run the exact same code on the CPU against the
// oracle before any GPU time is spent.  The kernels in kernels.cu add only the data movement,
// shared-memory histograms and atomics around these.
//
// Packed stream layout (see DESIGN.md "Data layout in HBM"):
//   * slot  = one base position in the global stream; block = 32 consecutive slots.
//   * codes : u32 words, 16 slots each, FIRST slot in the MOST significant bit pair
//             (slot s of a word sits at bits [30-2s, 31-2s]).  code = (ascii >> 1) & 3, i.e.
//             A=0 C=1 T=2 G=3 — the reference's encoding (count-kmers.cpp:77, kmer_utils.h:47,131).
//   * valid : u32 words, 32 slots each, slot s at bit s; 1 = byte was uppercase A/C/G/T and lies
//             inside a read (kmer_utils.h:38,122).  Padding slots are 0.
//   * every read starts on a block boundary and is followed by at least one padding slot, so a
//     15-mer window can never span two reads.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LRB_HD __host__ __device__ __forceinline__
#else
#define LRB_HD inline
#endif

namespace lrb {

constexpr uint32_t kMask15 = 0x3FFFFFFFu;  // 4^15 - 1 (kmer_utils.h:46)
constexpr uint32_t kTableEntries = 1u << 30;
constexpr int kTileBlocks = 256;  // max blocks (of 32 slots) of ONE read handled by one warp

// ((hi:lo) >> s) & 0xffffffff for s in [0,31]
LRB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31));
#endif
}

LRB_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

LRB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

LRB_HD uint32_t umulhi64_lo(uint32_t a, uint64_t m) {  // low 32 bits of floor(a*m / 2^64)
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umul64hi((uint64_t)a, m);
#else
    return (uint32_t)(((unsigned __int128)a * m) >> 64);
#endif
}

// Reverse the 16 base pairs of a code word and complement every base (A<->T, C<->G == XOR 0b10):
// the word-level form of revComp (kmer_utils.h:10-22).
LRB_HD uint32_t rc16(uint32_t w) {
    uint32_t x = brev32(w);                                           // reverses bits, swapping each pair's halves
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);          // un-swap inside each pair
    return x ^ 0xAAAAAAAAu;
}

// k-mer (K <= 16) whose LAST base is slot j (0..31) of the block (w0 = slots 0..15, w1 = slots
// 16..31); pw = the code word just before the block (slots -16..-1).  First base most significant,
// like the reference's rolling value.
template <int K>
LRB_HD uint32_t kmer_ending_at(uint32_t pw, uint32_t w0, uint32_t w1, int j) {
    constexpr uint32_t mask = (K == 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
    uint32_t x = (j < 16) ? funnel_r(w0, pw, 30 - 2 * j) : funnel_r(w1, w0, 62 - 2 * j);
    return x & mask;
}

// Reverse complement of the 15-mer ending at slot j, read out of the reverse-complemented block
// r0 = rc16(w1), r1 = rc16(w0), r2 = rc16(pw)  (r0 most significant).
LRB_HD uint32_t rc15_ending_at(uint32_t r0, uint32_t r1, uint32_t r2, int j) {
    uint32_t x;
    if (j <= 13) x = funnel_r(r2, r1, 2 * j + 4);
    else if (j <= 29) x = funnel_r(r1, r0, 2 * j - 28);
    else x = r0 >> (2 * j - 60);
    return x & kMask15;
}

// bit j set <=> the 15 slots ending at slot j of the current block are all valid
// (pv = validity word of the previous block, v = of this block).
LRB_HD uint32_t window15_mask(uint32_t pv, uint32_t v) {
    uint64_t a = ((uint64_t)v << 32) | pv;
    a &= a >> 1;   // runs of 2 starting at p
    a &= a >> 2;   // runs of 4
    a &= a >> 4;   // runs of 8
    a &= a >> 7;   // runs of 15
    return (uint32_t)(a >> 18);  // run starting at p = 32 + j - 14 ends at slot j
}

// valid word implied by the read length alone (every in-read slot an uppercase ACGT) for the block whose slot 0
// is read position p0; blocks that differ are shipped as exceptions (ingest.cpp, lrb_dev_fill_valid).
LRB_HD uint32_t default_valid_word(uint32_t len, uint64_t p0) {
    const uint64_t n_in = len > p0 ? len - p0 : 0;
    return n_in >= 32 ? 0xFFFFFFFFu : ((1u << (uint32_t)n_in) - 1u);
}

// The one key of {x, rc(x)} with bit 15 clear.  The middle base of a 15-mer sits at bits [14,15];
// reverse complement maps it to its own complement (XOR 2), so exactly one of the pair has bit 15 == 0.
// Every occurrence increments BOTH strands in the reference (kmer_utils.h:139-153), hence
// T[x] == T[rc(x)] always and counting the bit-15-clear member + mirroring reproduces the table.
LRB_HD uint32_t canonical15(uint32_t val, uint32_t rc) { return (val & 0x8000u) ? rc : val; }

// Scalar reverse complement of a 30-bit key (used by the mirror pass and tests).
LRB_HD uint32_t revcomp15(uint32_t x) {
    // left-align the 15 bases in a 16-base word (pad pair last); after rc16 the pad pair is first and the
    // reverse-complemented 15 bases occupy the low 30 bits.
    return rc16(x << 2) & kMask15;
}

// Histogram bin of a global 15-mer count (kmer_utils.h:54-69):
//   count < 2 -> 0; count <= S -> bin 0; pos = count/S - 1; 0 < pos < B -> pos; else B-1
// (so counts in (S, 2S) land in the LAST bin).  S32 = min(S, 2^32-1); magic = floor(2^64/S)+1 (S>1).
LRB_HD uint32_t coverage_bin(uint32_t count, uint32_t S32, uint64_t magic, uint32_t B) {
    uint32_t c = count < 2u ? 0u : count;
    if (c <= S32) return 0u;
    uint32_t q = (S32 == 1u) ? c : umulhi64_lo(c, magic);
    uint32_t pos = q - 1u;  // q >= 1 because c > S
    return (pos > 0u && pos < B) ? pos : (B - 1u);
}

inline uint64_t coverage_magic(uint32_t S32) { return S32 <= 1u ? 0ull : (~0ull / S32) + 1ull; }

// ---- per-block drivers (one lane, one 32-slot block) -------------------------------------------

// composition (count-kmers.cpp:73-87): which slots of the block END a k-mer window.  p0 = read position
// of slot 0 (a multiple of 32), len = read length.  No validity test: every byte takes part.
LRB_HD uint32_t comp_block_mask(uint32_t p0, uint32_t len, int k) {
    const uint32_t n_in = (len - p0) < 32u ? (len - p0) : 32u;  // in-read slots of this block (p0 <= len)
    uint32_t m = (n_in >= 32u) ? 0xFFFFFFFFu : ((1u << n_in) - 1u);
    if (p0 == 0) m &= ~((1u << (k - 1)) - 1u);  // the first k-1 positions of a read end no window
    return m;
}

template <int K, class F>
LRB_HD void comp_block(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t m, F f) {
    if (m == 0xFFFFFFFFu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f(kmer_ending_at<K>(pw, w0, w1, j));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if ((m >> j) & 1u) f(kmer_ending_at<K>(pw, w0, w1, j));
    }
}

// 15-mer windows of a block: f(key) with key = the bit-15-clear member of {val, rc(val)}
template <class F>
LRB_HD void canon15_block(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t m, F f) {
    const uint32_t r0 = rc16(w1), r1 = rc16(w0), r2 = rc16(pw);
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if ((m >> j) & 1u) f(canonical15(kmer_ending_at<15>(pw, w0, w1, j), rc15_ending_at(r0, r1, r2, j)));
}

// mirror pass index algebra.  Table index x = (H:14 | M:2 | L:14); rc(x) = (rc7(L) | M^2 | rc7(H)).
// Tile t in [0, 2^17) = (hrest:8 | mlow:1 | lrest:8); destination H = (a:6 | hrest), L = (lrest | b:6),
// M = 2|mlow (bit 15 set).  The source tile holds rows sb = rc3(b) of 64 contiguous entries sa = rc3(a).
LRB_HD uint32_t rc_small(uint32_t x, int nbases) {  // reverse complement of nbases (<= 16) bases
    return rc16(x << (32 - 2 * nbases)) & ((1u << (2 * nbases)) - 1u);
}
LRB_HD uint32_t mirror_dst_index(uint32_t t, uint32_t a, uint32_t b) {
    const uint32_t lrest = t & 0xFFu, mlow = (t >> 8) & 1u, hrest = t >> 9;
    return (((a << 8) | hrest) << 16) | ((2u | mlow) << 14) | (lrest << 6) | b;
}
LRB_HD uint32_t mirror_src_index(uint32_t t, uint32_t sb, uint32_t sa) {
    const uint32_t lrest = t & 0xFFu, mlow = (t >> 8) & 1u, hrest = t >> 9;
    return (((sb << 8) | rc_small(lrest, 4)) << 16) | (mlow << 14) | (rc_small(hrest, 4) << 6) | sa;
}

}  // namespace lrb
