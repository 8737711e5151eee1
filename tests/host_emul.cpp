// host_emul.cpp — runs the kernels' per-lane code (csrc/lane_core.cuh) on the CPU.
//
// Test infrastructure: the CUDA kernels in csrc/kernels.cu are thin shells (loads, shared-memory
// histograms, atomics) around the LRB_HD functions of lane_core.cuh.  This file re-creates those shells
// as plain loops over tiles / lanes / blocks so the bit arithmetic (window extraction, validity masks,
// canonical keys, mirror index algebra, bucket rule, packing layout) is checked against the oracle in the
// CPU-only test tier, before GPU time is spent.  It is NOT a fallback: the product never links it.
#include <stdint.h>
#include <string.h>

#include "../include/lrbinner_b200.h"
#include "../lrbinner_b200/csrc/lane_core.cuh"
#include "../lrbinner_b200/csrc/fixed6.h"

using namespace lrb;

template <int K>
static void comp_emul(const lrb_reads_view* R, const uint16_t* lut, int P, uint32_t* out) {
    for (uint64_t tile = 0; tile < R->n_tiles; ++tile) {
        const uint32_t r = R->tile_read[tile], b0 = R->tile_blk[tile];
        const uint32_t rb0 = R->read_blk[r], rb1 = R->read_blk[r + 1], len = R->read_len[r];
        const uint32_t nblk = (rb1 - b0) < (uint32_t)kTileBlocks ? (rb1 - b0) : (uint32_t)kTileBlocks;
        for (uint32_t lane = 0; lane < 32; ++lane)
            for (uint32_t i = lane; i < nblk; i += 32) {
                const uint32_t gb = b0 + i;
                const uint32_t w0 = R->codes[2 * (size_t)gb], w1 = R->codes[2 * (size_t)gb + 1];
                const uint32_t p0 = (gb - rb0) * 32u;
                const uint32_t pw = (p0 != 0) ? R->codes[2 * (size_t)gb - 1] : 0u;
                const uint32_t m = comp_block_mask(p0, len, K);
                comp_block<K>(pw, w0, w1, m, [&](uint32_t kmer) { out[(size_t)r * P + lut[kmer]]++; });
            }
    }
}

extern "C" int emul_composition(const lrb_reads_view* R, int k, const uint16_t* lut, uint32_t* out) {
    if (k == 3) comp_emul<3>(R, lut, 32, out);
    else if (k == 4) comp_emul<4>(R, lut, 136, out);
    else if (k == 5) comp_emul<5>(R, lut, 512, out);
    else return 1;
    return 0;
}

extern "C" int emul_count(const lrb_reads_view* R, uint32_t* table, uint32_t key_lo, uint32_t key_hi) {
    for (uint64_t gb = 0; gb < R->n_blocks; ++gb) {
        const uint32_t v = R->valid[gb], pv = gb ? R->valid[gb - 1] : 0u;
        const uint32_t m = window15_mask(pv, v);
        if (!m) continue;
        const uint32_t w0 = R->codes[2 * gb], w1 = R->codes[2 * gb + 1], pw = gb ? R->codes[2 * gb - 1] : 0u;
        canon15_block(pw, w0, w1, m, [&](uint32_t key) {
            if (key >= key_lo && key < key_hi) table[key]++;
        });
    }
    return 0;
}

// table[rc(key)] = table[key] for the listed bit-15-clear keys (sparse stand-in for k_mirror)
extern "C" int emul_mirror_keys(uint32_t* table, const uint32_t* keys, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        if (keys[i] & 0x8000u) return 1;
        table[revcomp15(keys[i])] = table[keys[i]];
    }
    return 0;
}

// checks k_mirror's index algebra over EVERY destination: dst has bit 15 set, src == rc(dst), the smem
// coordinates the kernel reads are the ones it wrote, and destinations cover each bit-15-set index once.
extern "C" uint64_t emul_mirror_check(uint8_t* seen /* 2^30 bits = 128 MiB, zeroed */) {
    uint64_t bad = 0;
    for (uint32_t t = 0; t < (1u << 17); ++t)
        for (uint32_t a = 0; a < 64; ++a)
            for (uint32_t tx = 0; tx < 64; ++tx) {
                const uint32_t dst = mirror_dst_index(t, a, tx);
                const uint32_t sb = rc_small(tx, 3), sa = rc_small(a, 3);  // tile[sb][sa] as the kernel reads it
                const uint32_t src = mirror_src_index(t, sb, sa);          // ... and the address that filled it
                if (!(dst & 0x8000u) || (src & 0x8000u) || dst >= kTableEntries || src != revcomp15(dst)) ++bad;
                if (seen[dst >> 3] & (1u << (dst & 7))) ++bad;
                seen[dst >> 3] |= (uint8_t)(1u << (dst & 7));
            }
    return bad;
}

extern "C" int emul_search(const lrb_reads_view* R, const uint32_t* table, long bin_size, int bins, uint32_t* hist,
                           uint32_t* sums, uint32_t key_lo, uint32_t key_hi) {
    const uint32_t S32 = bin_size > 0xFFFFFFFFl ? 0xFFFFFFFFu : (uint32_t)bin_size;
    const uint64_t magic = coverage_magic(S32);
    const bool filter = !(key_lo == 0 && key_hi >= kTableEntries);
    for (uint64_t tile = 0; tile < R->n_tiles; ++tile) {
        const uint32_t r = R->tile_read[tile], b0 = R->tile_blk[tile], rb1 = R->read_blk[r + 1];
        const uint32_t nblk = (rb1 - b0) < (uint32_t)kTileBlocks ? (rb1 - b0) : (uint32_t)kTileBlocks;
        for (uint32_t lane = 0; lane < 32; ++lane)
            for (uint32_t i = lane; i < nblk; i += 32) {
                const uint32_t gb = b0 + i;
                const uint32_t v = R->valid[gb], pv = gb ? R->valid[gb - 1] : 0u;
                const uint32_t m = window15_mask(pv, v);
                if (!m) continue;
                const uint32_t w0 = R->codes[2 * (size_t)gb], w1 = R->codes[2 * (size_t)gb + 1];
                const uint32_t pw = gb ? R->codes[2 * (size_t)gb - 1] : 0u;
                const uint32_t r0 = rc16(w1), r1 = rc16(w0), r2 = rc16(pw);
                for (int j = 0; j < 32; ++j) {
                    if (!((m >> j) & 1u)) continue;
                    const uint32_t val = kmer_ending_at<15>(pw, w0, w1, j);
                    uint32_t cnt;
                    if (filter) {
                        const uint32_t key = canonical15(val, rc15_ending_at(r0, r1, r2, j));
                        if (!(key >= key_lo && key < key_hi)) continue;
                        cnt = table[key];
                    } else {
                        cnt = table[val];
                    }
                    hist[(size_t)r * bins + coverage_bin(cnt, S32, magic, (uint32_t)bins)]++;
                    sums[r]++;
                }
            }
    }
    return 0;
}

extern "C" uint32_t emul_coverage_bin(uint32_t count, long bin_size, int bins) {
    const uint32_t S32 = bin_size > 0xFFFFFFFFl ? 0xFFFFFFFFu : (uint32_t)bin_size;
    return coverage_bin(count, S32, coverage_magic(S32), (uint32_t)bins);
}
extern "C" uint32_t emul_revcomp15(uint32_t x) { return revcomp15(x); }
