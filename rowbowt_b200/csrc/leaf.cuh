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
