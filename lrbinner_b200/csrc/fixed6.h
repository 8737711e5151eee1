// fixed6.h — the six decimals glibc's "%f" prints for count/total, computed exactly (host + device).
//
// The tools print std::to_string(double) == printf("%f") of  double(count)/max(1.0,total)
// (count-kmers.cpp:89-92) resp. counts[i] /= sum with the `< 1e-4 -> 0` rule (kmer_utils.h:75-84).
// printf rounds the EXACT binary value of that double to 6 decimals, ties to even.  For a double
// d = m * 2^-s (m < 2^53) the digits are round_half_even(m * 10^6 / 2^s), evaluated here in 128-bit
// integer arithmetic — no libc call, so the same code runs in the CUDA text epilogue.
#pragma once
#include <stdint.h>
#include <string.h>

#include "lane_core.cuh"

namespace lrb {

LRB_HD void mul64x64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    *hi = (uint64_t)(p >> 64);
#endif
}

// Returns q in [0, 10^6] such that printf("%f", v) == "<q/10^6>.<q%10^6 as 6 digits>", where
// v = double(num)/double(max(den,1)), set to 0 first when `coverage` and v < 1e-4.  Requires num <= max(den,1).
LRB_HD uint32_t fixed6(uint32_t num, uint32_t den, bool coverage) {
    if (num == 0) return 0;
    const double d = (double)num / (double)(den ? den : 1u);  // IEEE round-to-nearest on host and device
    if (coverage && d < 1e-4) return 0;
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(d);
#else
    memcpy(&bits, &d, sizeof bits);
#endif
    const int e = (int)((bits >> 52) & 0x7FF);
    const uint64_t m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    const int s = 1075 - e;  // d = m * 2^-s ; d in [2^-32, 1] => s in [52, 84]
    if (s < 1 || s > 126) return d >= 1.0 ? 1000000u : 0u;
    uint64_t hi, lo;
    mul64x64(m, 1000000ull, &hi, &lo);  // < 2^73
    uint64_t q, rem_hi, rem_lo, half_hi, half_lo;
    if (s < 64) {
        q = (lo >> s) | (hi << (64 - s));
        rem_hi = 0;
        rem_lo = lo & ((1ull << s) - 1);
        half_hi = 0;
        half_lo = 1ull << (s - 1);
    } else {
        q = (s == 64) ? hi : (hi >> (s - 64));
        rem_hi = (s == 64) ? 0 : (hi & ((1ull << (s - 64)) - 1));
        rem_lo = lo;
        half_hi = (s == 64) ? 0 : (1ull << (s - 65));
        half_lo = (s == 64) ? (1ull << 63) : 0;
    }
    const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = rem_hi == half_hi && rem_lo == half_lo;
    if (gt || (eq && (q & 1ull))) ++q;
    return (uint32_t)q;
}

}  // namespace lrb
