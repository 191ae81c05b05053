// Decoding of one 64-byte rank-directory leaf (layout.hpp).  Shared by the kernels and by
// the host-side layout self-check (rbg_selftest_layout), so both read lines identically.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RBG_HD __host__ __device__ __forceinline__
#else
#define RBG_HD inline
#endif

namespace rbg {

RBG_HD uint32_t rbg_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t) __popc(x);
#else
    return (uint32_t) __builtin_popcount(x);
#endif
}

// Number of occurrences of the leaf's symbol in leaf positions [0,q), and whether position q
// itself holds the symbol.  w = the 16 words of the line.
RBG_HD uint32_t leaf_count(const uint32_t (&w)[16], uint32_t q, bool& inside) {
    const uint32_t mode = (w[1] >> 8) & 0xF;
    uint32_t cnt = 0;
    bool in = false;
    if (mode == 0) {                       // RUNS: (len << 16 | start) x 14
#pragma unroll
        for (int e = 0; e < 14; ++e) {
            const uint32_t f = w[2 + e];
            const int32_t d = (int32_t) q - (int32_t) (f & 0xFFFFu);
            const uint32_t len = f >> 16;
            const uint32_t before = (uint32_t) (d > 0 ? d : 0);
            cnt += before < len ? before : len;
            in = in || ((uint32_t) d < len);
        }
    } else {                               // BITS: 256-bit bitmap in w[2..9]
        uint32_t hit = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int32_t rem = (int32_t) q - 32 * i;
            const uint32_t mask = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            const uint32_t bit = (uint32_t) rem < 32u ? (1u << rem) : 0u;      // position q itself
            cnt += rbg_popc(w[2 + i] & mask);
            hit |= w[2 + i] & bit;
        }
        in = hit != 0;
    }
    inside = in;
    return cnt;
}

RBG_HD uint64_t leaf_base_count(const uint32_t (&w)[16]) {
    return (uint64_t) w[0] | ((uint64_t) (w[1] & 0xFFu) << 32);
}

}  // namespace rbg

// ---------------------------------------------------------------------------------------------
// Mixed leaf (layout v2, MixDir in layout.hpp): one 64-byte line answers rank_c for ALL symbols
// over a fixed window of 2^g BWT positions.
//   w[0..3]   low 32 bits of F[c] + #c in BWT[0, leaf_start), c = A,C,G,T
//   w[4]      byte c = bits 32..39 of the same
//   w[5..15]  22 x u16 entries (head << 13 | start): every run intersecting the leaf, in BWT
//             order, start relative to the leaf (the first is 0); head 0..3 = A,C,G,T,
//             4 = terminator.  Unused entries carry start = leaf size (they cover nothing).
//   A leaf with more than 22 runs is SPLIT: entry 0 = 0x1FFF, w[6] = line index of its first
//   child, w[7] = k; 2^k children of 2^(g-k) positions each, same format, never split again.
namespace rbg {

constexpr int kMixEntries = 22;
constexpr uint32_t kMixSplit = 0x1FFFu;
constexpr uint32_t kMixStartMask = 0x1FFFu;
constexpr int kMixHeadShift = 13;

RBG_HD bool mix_is_split(const uint32_t (&w)[16]) { return (w[5] & kMixStartMask) == kMixSplit; }

RBG_HD uint64_t mix_base_count(const uint32_t (&w)[16], uint32_t c) {
    const uint32_t lo = c == 0 ? w[0] : c == 1 ? w[1] : c == 2 ? w[2] : w[3];
    return (uint64_t) lo | ((uint64_t) ((w[4] >> (8 * c)) & 0xFFu) << 32);
}

RBG_HD uint32_t mix_min(uint32_t a, uint32_t b) { return a < b ? a : b; }

// #c in leaf positions [0,qa) and [0,qb) (qa, qb <= leaf_size), one branch-free pass over the
// 22 entries: run e covers [s_e, s_{e+1}), so its share of [0,q) is min(q,s_{e+1}) - min(q,s_e).
// With THIRD also [0,qc) (the toehold test BWT[hi]==c needs rank at hi and hi+1).
template <bool SECOND, bool THIRD>
RBG_HD void mix_count(const uint32_t (&w)[16], uint32_t c, uint32_t leaf_size, uint32_t qa, uint32_t qb, uint32_t qc,
                      uint32_t& ra, uint32_t& rb, uint32_t& rc) {
    uint32_t pa = 0, pb = 0, pc = 0;                 // min(q, s_e); s_0 == 0
    uint32_t h = (w[5] >> kMixHeadShift) & 7u;
    ra = rb = rc = 0;
#pragma unroll
    for (int e = 0; e < kMixEntries; ++e) {
        uint32_t ns = leaf_size, nh = 0;
        if (e + 1 < kMixEntries) {
            const uint32_t f = (w[5 + ((e + 1) >> 1)] >> (16 * ((e + 1) & 1))) & 0xFFFFu;
            ns = f & kMixStartMask;
            nh = f >> kMixHeadShift;
        }
        const uint32_t na = mix_min(qa, ns);
        const uint32_t nb = SECOND ? mix_min(qb, ns) : 0u;
        const uint32_t nc = THIRD ? mix_min(qc, ns) : 0u;
        if (h == c) {
            ra += na - pa;
            if (SECOND) rb += nb - pb;
            if (THIRD) rc += nc - pc;
        }
        pa = na;
        pb = nb;
        pc = nc;
        h = nh;
    }
}

}  // namespace rbg
