// synth_core.h — deterministic synthetic long reads (bench / test INPUT generator, not part of the
// reference's algorithm).  Integer-only, counter-based hashing, so the CUDA generator (lrb_dev_synth)
// and the host generator (lrb_synth_host) emit byte-identical reads; read metadata (genome, start,
// strand, length) is drawn on the host by lrbinner_b200/synth.py with numpy.random.default_rng(seed)
// following SURVEY.md section 8(d).
#pragma once
#include <stdint.h>

#include "../../include/lrbinner_b200.h"
#include "lane_core.cuh"

namespace lrb {

LRB_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// base index 0..3 = A,C,G,T of genome g at position i
LRB_HD uint32_t genome_base(uint64_t seed, uint32_t g, uint32_t i) {
    return (uint32_t)(mix64(seed ^ ((uint64_t)(g + 1) << 40) ^ (uint64_t)i) >> 33) & 3u;
}

// Emits the `len` bases of read r one ASCII byte at a time through emit(pos, byte).
// Error model per emitted/consumed base (thresholds are cumulative fractions of 2^32):
//   deletion  : skip one genome base, emit nothing
//   insertion : emit a random base, stay on the same genome base
//   substitute: emit a base different from the genome's
// then with probability n_thr the emitted base is replaced by 'N'.  flags bit0 = reverse strand
// (walk the genome backwards, complemented), bit1 = whole read lowercase.
template <class Emit>
LRB_HD void synth_read(const lrb_synth_params& p, const uint32_t* glen, uint64_t r, uint32_t g, uint32_t start,
                       uint32_t flags, uint32_t len, Emit emit) {
    const char upper[4] = {'A', 'C', 'G', 'T'};
    const uint32_t gl = glen[g];
    const bool rev = flags & 1u, lower = flags & 2u;
    uint32_t gpos = start % gl;
    uint64_t ctr = 0;
    const uint64_t rkey = mix64(p.seed ^ 0xA5A5A5A5ull) ^ ((r + p.read_base) * 0x9E3779B97F4A7C15ull);
    const uint32_t t_del = p.del_thr, t_ins = t_del + p.ins_thr, t_sub = t_ins + p.sub_thr;
    for (uint32_t out = 0; out < len;) {
        const uint64_t h = mix64(rkey + ctr++);
        const uint32_t u = (uint32_t)h;
        const uint32_t rnd = (uint32_t)(h >> 32);
        uint32_t gb = genome_base(p.seed, g, gpos);
        if (rev) gb = 3u - gb;  // complement in A,C,G,T index space
        uint32_t b;
        bool advance = true;
        if (u < t_del) {  // deletion
            gpos = rev ? (gpos == 0 ? gl - 1 : gpos - 1) : (gpos + 1 == gl ? 0 : gpos + 1);
            continue;
        } else if (u < t_ins) {
            b = rnd & 3u;
            advance = false;
        } else if (u < t_sub) {
            b = (gb + 1u + (rnd % 3u)) & 3u;
        } else {
            b = gb;
        }
        if (advance) gpos = rev ? (gpos == 0 ? gl - 1 : gpos - 1) : (gpos + 1 == gl ? 0 : gpos + 1);
        char c = upper[b];
        if (p.n_thr && (uint32_t)(mix64(h) >> 32) < p.n_thr) c = 'N';
        if (lower) c = (char)(c | 0x20);
        emit(out, c);
        ++out;
    }
}

}  // namespace lrb
