// The 64-byte rank-directory leaf ("mixed leaf", layout.hpp) and its decode.  Shared by the
// kernels and by the host-side layout self-check (rbg_selftest_layout), so both read lines
// identically; on the device the helpers below are single sm_100a instructions
// (VIMNMX.U16x2, IDP.2A, SHF), on the host plain C.
//
// One line answers rank_c for ALL four symbols over a window of 2^g BWT positions
// (replaces rle_string::rank, include/rle_string.hpp:131-161, ~10 dependent probes):
//   w[0..3]   low 32 bits of F[c] + #c in BWT[0, line_start), c = A,C,G,T
//   w[4]      byte c = bits 32..39 of the same
//   w[5]      heads (2 bits: A,C,G,T) of entries 0..15, TRANSPOSED: entry e sits at bit
//             8*(e&3) + 2*(e>>2), so that (eq >> 2i) & 0x01010101 is the byte vector of
//             entries 4i..4i+3
//   w[6..14]  18 x u16 run starts, entry e in the (e&1) half of w[6 + e/2]: every run that
//             intersects the line, in BWT order, start relative to the 2^g window.  Unused
//             entries carry 0xFFFF (they cover nothing).
//   w[15]     bits 0-1 head of entry 16, bits 8-9 head of entry 17, bits 16-19 mode,
//             bits 20-21 symbol the terminator is counted as (mode TERM)
// rank: run e covers [s_e, s_{e+1}), so with m_e = min(q, s_e) and X_e = [head_e == c]
//   #c in [line_start, window_start + q) = sum_e X_e (m_{e+1} - m_e)
//                                        = sum_e X_{e-1} m_e  -  sum_e X_e m_e      (m_18 = q)
// i.e. 9 packed mins and 18 two-way dot products, no branches, no per-entry extraction.
//
// Modes.  NORMAL: the line is the window's only line.  SPLIT: the window holds more than 18
// runs; the line is an index: w[0] = line index of its first child, w[6..14] = window-relative
// starts of up to 18 children (first 0, unused 0xFFFF).  Children are NORMAL/TERM lines in
// window coordinates (their first start is the child's start, their counts are taken there).
// TERM: the line covers a terminator (byte 1), which has no 2-bit code: its position is merged
// into the neighbouring run and rank of that symbol is corrected from the (<= 8) terminator
// positions kept in the kernel parameters.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RBG_HD __host__ __device__ __forceinline__
#else
#define RBG_HD inline
#endif

namespace rbg {

constexpr int kLeafEntries = 18;
constexpr uint32_t kLeafPad = 0xFFFFu;
constexpr uint32_t kModeShift = 16, kModeMask = 0xFu << kModeShift;
constexpr uint32_t kModeSplit = 1u << kModeShift, kModeTerm = 2u << kModeShift;
constexpr uint32_t kTermSymShift = 20;

RBG_HD uint32_t rbg_vminu2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    const uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
#endif
}
RBG_HD uint32_t rbg_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xFFFFu) * (b & 0xFFu) + (a >> 16) * ((b >> 8) & 0xFFu);
#endif
}
RBG_HD uint32_t rbg_dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_hi(a, b, c);
#else
    return c + (a & 0xFFFFu) * ((b >> 16) & 0xFFu) + (a >> 16) * (b >> 24);
#endif
}
RBG_HD uint32_t rbg_funnel_l8(uint32_t lo, uint32_t hi) {      // (hi:lo) << 8, upper word
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 8);
#else
    return (hi << 8) | (lo >> 24);
#endif
}

// Bit position of entry e's head in w[5] (e < 16).
RBG_HD uint32_t leaf_head_bit(uint32_t e) { return 8u * (e & 3u) + 2u * (e >> 2); }

// Per-symbol state of one LF step: the pattern that turns head fields equal to c into zero.
RBG_HD uint32_t leaf_cpat(uint32_t c) { return c * 0x55555555u; }

RBG_HD uint64_t leaf_base_count(const uint32_t (&w)[16], uint32_t c) {
    const uint32_t lo = c == 0 ? w[0] : c == 1 ? w[1] : c == 2 ? w[2] : w[3];
    return (uint64_t) lo | ((uint64_t) ((w[4] >> (8 * c)) & 0xFFu) << 32);
}

// Byte vectors X (entries 4i..4i+3 -> byte of xb[i], 1 where the head equals c).
RBG_HD void leaf_match(const uint32_t (&w)[16], uint32_t cpat, uint32_t (&xb)[5]) {
    const uint32_t x = w[5] ^ cpat;
    const uint32_t eq = ~(x | (x >> 1));
    const uint32_t y = w[15] ^ cpat;
    xb[0] = eq & 0x01010101u;
    xb[1] = (eq >> 2) & 0x01010101u;
    xb[2] = (eq >> 4) & 0x01010101u;
    xb[3] = (eq >> 6) & 0x01010101u;
    xb[4] = ~(y | (y >> 1)) & 0x00000101u;
}

// #c in [count point of the line, window_start + q) given the match vectors, q <= 2^g.
RBG_HD uint32_t leaf_rank_x(const uint32_t (&w)[16], const uint32_t (&xb)[5], const uint32_t (&xs)[5], uint32_t q) {
    const uint32_t qq = q | (q << 16);
    uint32_t plus0 = 0, plus1 = 0, minus0 = 0, minus1 = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint32_t m = rbg_vminu2(qq, w[6 + j]);
        if (j & 1) {
            plus1 = rbg_dp2a_hi(m, xs[j >> 1], plus1);
            minus1 = rbg_dp2a_hi(m, xb[j >> 1], minus1);
        } else {
            plus0 = rbg_dp2a_lo(m, xs[j >> 1], plus0);
            minus0 = rbg_dp2a_lo(m, xb[j >> 1], minus0);
        }
    }
    return plus0 + plus1 + (xb[4] >> 8) * q - minus0 - minus1;        // + X_17 * m_18, m_18 = q
}

RBG_HD void leaf_shift_match(const uint32_t (&xb)[5], uint32_t (&xs)[5]) {   // xs byte of entry e = X_{e-1}
    xs[0] = xb[0] << 8;
#pragma unroll
    for (int i = 1; i < 5; ++i) xs[i] = rbg_funnel_l8(xb[i - 1], xb[i]);
}

RBG_HD uint32_t leaf_rank(const uint32_t (&w)[16], uint32_t cpat, uint32_t q) {
    uint32_t xb[5], xs[5];
    leaf_match(w, cpat, xb);
    leaf_shift_match(xb, xs);
    return leaf_rank_x(w, xb, xs, q);
}

// SPLIT line: which child holds window position q (number of child starts <= q, minus one).
RBG_HD uint32_t leaf_child_of(const uint32_t (&w)[16], uint32_t q) {
    uint32_t idx = 0;
    for (int e = 1; e < kLeafEntries; ++e) {
        const uint32_t s = (w[6 + (e >> 1)] >> (16 * (e & 1))) & 0xFFFFu;
        idx += s <= q ? 1u : 0u;
    }
    return idx;
}

// Window-relative position at which a line's counts are taken (0 for a direct line).
RBG_HD uint32_t leaf_first_start(const uint32_t (&w)[16]) { return w[6] & 0xFFFFu; }

}  // namespace rbg
