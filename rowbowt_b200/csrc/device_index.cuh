// Device-side views of the index (plain structs passed to kernels by value) and the
// per-step device functions: rank-pair / LF, toehold resolution, phi, marker windows.
#pragma once
#include <cstdint>

#include "leaf.cuh"

namespace rbg {

constexpr int kDevMaxTerm = 8;

struct DevRankDir {
    const uint32_t* lines;      // 64-byte leaves, 64-byte aligned
    const uint32_t* table;      // [4][n_buckets]
    uint64_t n_buckets;
    uint64_t n;
    uint32_t s;
    uint32_t n_term;
    uint64_t term_pos[kDevMaxTerm];
};

// Layout v2 (mixed leaves, leaf.cuh): line of BWT position p is p >> g; no table.
struct DevMixDir {
    const uint32_t* lines;      // 64-byte mixed leaves: [n_direct] direct, then split children
    uint64_t n;
    uint32_t g;
    uint32_t n_term;
    uint64_t term_pos[kDevMaxTerm];
};

struct DevPredTable {
    const uint64_t* keys;
    const uint32_t* table;
    uint64_t n_keys;
    uint32_t shift;
};

struct DevToehold {
    DevPredTable rows;
    const uint64_t* sample;
    uint64_t toehold0;
};

struct DevPhi {
    DevPredTable pred;
    const uint64_t* prev;
    uint64_t n;
};

struct DevMarkers {
    const uint64_t *starts, *ends, *idxs, *arr;
    uint64_t n_starts, n_ends, n_idxs;
    uint64_t size_starts, size_ends, size_idxs;
};

#if defined(__CUDACC__)

// One 64-byte line as two 256-bit read-only loads (LDG.E.256 on sm_100a); lines are touched
// once per step and would only evict the table / read words from L1, so no L1 allocation.
__device__ __forceinline__ void load_line(const uint32_t* p, uint32_t (&w)[16]) {
    asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
        : "l"(p));
    asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
        : "l"(p + 8));
}

// Address of the leaf of symbol c that covers BWT position pos, and pos's offset inside it.
__device__ __forceinline__ uint32_t leaf_of(const DevRankDir& D, uint32_t entry, uint64_t pos, uint32_t& q) {
    const uint32_t k = entry & 15u;
    const uint32_t g = D.s - k;
    q = (uint32_t) pos & ((1u << g) - 1u);
    return (entry >> 4) + ((uint32_t) (pos >> g) & ((1u << k) - 1u));
}

// RowBowt::LF(range,c) (include/rowbowt.hpp:74-88) for c in {A,C,G,T} (code 0..3, known present):
// rank_c(lo) and rank_c(hi+1) from one or two leaves.  Also reports BWT[hi]==c (LF_w_loc's
// trivial-case test, include/rowbowt.hpp:559).  Returns false when the new range is empty.
__device__ __forceinline__ bool lf_step(const DevRankDir& D, uint32_t c, uint64_t& lo, uint64_t& hi,
                                        bool& hi_is_c, uint32_t& lines_touched) {
    const uint32_t* tb = D.table + (uint64_t) c * D.n_buckets;
    const uint64_t blo = lo >> D.s, bhi = hi >> D.s;
    const uint32_t elo = __ldg(tb + blo);
    const uint32_t ehi = bhi == blo ? elo : __ldg(tb + bhi);
    uint32_t qlo, qhi;
    const uint32_t leaf_lo = leaf_of(D, elo, lo, qlo);
    const uint32_t leaf_hi = leaf_of(D, ehi, hi, qhi);
    uint32_t A[16], B[16];
    load_line(D.lines + (uint64_t) leaf_lo * 16, A);
    if (leaf_hi != leaf_lo) {
        load_line(D.lines + (uint64_t) leaf_hi * 16, B);
        lines_touched += 2;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) B[i] = A[i];
        lines_touched += 1;
    }
    bool in_lo, in_hi;
    // leaf headers carry F[c] + #c before the leaf, so these are already rows of the F column
    const uint64_t new_lo = leaf_base_count(A) + leaf_count(A, qlo, in_lo);                       // F[c] + #c in [0,lo)
    const uint64_t new_end = leaf_base_count(B) + leaf_count(B, qhi, in_hi) + (in_hi ? 1u : 0u);  // F[c] + #c in [0,hi]
    hi_is_c = in_hi;
    if (new_end == new_lo) return false;
    lo = new_lo;
    hi = new_end - 1;
    return true;
}

// Mixed-leaf line holding BWT position pos (a split leaf costs one more dependent load), pos's
// offset inside it and the window size of that line.  Returns the line index.
__device__ __forceinline__ uint64_t mix_fetch(const DevMixDir& D, uint64_t pos, uint32_t (&w)[16], uint32_t& q, uint32_t& size) {
    uint64_t idx = pos >> D.g;
    load_line(D.lines + idx * 16, w);
    size = 1u << D.g;
    if (mix_is_split(w)) {
        const uint32_t k = w[7], cg = D.g - k;
        idx = (uint64_t) w[6] + ((uint32_t) (pos >> cg) & ((1u << k) - 1u));
        load_line(D.lines + idx * 16, w);
        size = 1u << cg;
    }
    q = (uint32_t) pos & (size - 1u);
    return idx;
}

// RowBowt::LF(range,c) (include/rowbowt.hpp:74-88) on layout v2: both ranks from ONE line when lo
// and hi share a leaf (the common case once the range is narrow), else from two.
template <bool TOEHOLD>
__device__ __forceinline__ bool lf_step_mix(const DevMixDir& D, uint32_t c, uint64_t& lo, uint64_t& hi,
                                            bool& hi_is_c, uint32_t& lines_touched) {
    uint32_t A[16];
    uint32_t qa, sa, ra, rb, rc;
    mix_fetch(D, lo, A, qa, sa);
    uint64_t new_lo, new_end;
    if (((hi ^ lo) & ~(uint64_t) (sa - 1u)) == 0) {      // hi in the same line (same child when split)
        const uint32_t qb = ((uint32_t) hi & (sa - 1u)) + 1u;
        mix_count<true, TOEHOLD>(A, c, sa, qa, qb, qb - 1u, ra, rb, rc);
        const uint64_t base = mix_base_count(A, c);
        new_lo = base + ra;
        new_end = base + rb;
        lines_touched += 1;
    } else {
        uint32_t B[16];
        uint32_t qb, sb, unused;
        mix_fetch(D, hi, B, qb, sb);
        mix_count<false, false>(A, c, sa, qa, 0, 0, ra, unused, unused);
        mix_count<TOEHOLD, false>(B, c, sb, qb + 1u, qb, 0, rb, rc, unused);
        new_lo = mix_base_count(A, c) + ra;
        new_end = mix_base_count(B, c) + rb;
        lines_touched += 2;
    }
    hi_is_c = TOEHOLD ? (rb != rc) : false;          // rank(hi+1) - rank(hi) == 1  <=>  BWT[hi] == c
    if (new_end == new_lo) return false;
    lo = new_lo;
    hi = new_end - 1;
    return true;
}

// One entry point for both layouts.
template <bool TOEHOLD>
__device__ __forceinline__ bool lf_any(const DevRankDir& D, uint32_t c, uint64_t& lo, uint64_t& hi, bool& hi_is_c, uint32_t& lines) {
    return lf_step(D, c, lo, hi, hi_is_c, lines);
}
template <bool TOEHOLD>
__device__ __forceinline__ bool lf_any(const DevMixDir& D, uint32_t c, uint64_t& lo, uint64_t& hi, bool& hi_is_c, uint32_t& lines) {
    return lf_step_mix<TOEHOLD>(D, c, lo, hi, hi_is_c, lines);
}

// Same for the terminator (byte 1) as a query symbol: rank over the sorted term_pos list, F[1] = 0.
template <class Dir>
__device__ __forceinline__ bool lf_step_term(const Dir& D, uint64_t& lo, uint64_t& hi, bool& hi_is_c) {
    uint64_t before = 0, upto = 0;
    hi_is_c = false;
    for (uint32_t t = 0; t < D.n_term; ++t) {
        before += D.term_pos[t] < lo;
        upto += D.term_pos[t] <= hi;
        hi_is_c = hi_is_c || D.term_pos[t] == hi;
    }
    if (upto == before) return false;
    lo = before;
    hi = upto - 1;
    return true;
}

// #keys < x
__device__ __forceinline__ uint64_t pred_rank(const DevPredTable& T, uint64_t x) {
    const uint64_t b = x >> T.shift;
    uint64_t a = __ldg(T.table + b), z = __ldg(T.table + b + 1);
    while (a < z) {
        const uint64_t mid = (a + z) >> 1;
        if (__ldg(T.keys + mid) < x) a = mid + 1; else z = mid;
    }
    return a;
}

// Toehold after a non-trivial step: `row` is the LF image of a run end (see ToeholdDir).
__device__ __forceinline__ uint64_t toehold_at_row(const DevToehold& T, uint64_t row) {
    return __ldg(T.sample + pred_rank(T.rows, row));
}

// ToeholdSA::phi, include/toehold_sa.hpp:56-72
__device__ __forceinline__ uint64_t phi_step(const DevPhi& P, uint64_t i) {
    const uint64_t rk = pred_rank(P.pred, i);
    const uint64_t jr = rk == 0 ? P.pred.n_keys - 1 : rk - 1;     // predecessor_rank_circular
    const uint64_t j = __ldg(P.pred.keys + jr);
    const uint64_t delta = j < i ? i - j : i + 1;
    uint64_t v = __ldg(P.prev + jr) + delta;                       // < 2n
    return v >= P.n ? v - P.n : v;
}

__device__ __forceinline__ uint64_t dev_lower_bound(const uint64_t* a, uint64_t m, uint64_t x) {
    uint64_t lo = 0, hi = m;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// rle_window_arr::at_range (pfbwt-f/include/rle_window_array.hpp:130-154) reduced to the index
// range [first,last) of windows it would append, with its clamped rank/select helpers (:202-232).
__device__ __forceinline__ void marker_windows(const DevMarkers& M, uint64_t s, uint64_t e, uint64_t& first, uint64_t& last) {
    first = last = 0;
    // run_starts_rank(e): #starts <= e, clamped at the bit-vector size
    const uint64_t e_rs_rank = e + 1 >= M.size_starts ? M.n_starts : dev_lower_bound(M.starts, M.n_starts, e + 1);
    if (e_rs_rank == 0) return;
    const uint64_t e_rs_pos = e_rs_rank > M.n_starts ? M.size_starts : __ldg(M.starts + e_rs_rank - 1);
    if (e_rs_pos <= s) {
        const uint64_t e_re_pos = e_rs_rank > M.n_ends ? M.size_ends : __ldg(M.ends + e_rs_rank - 1);
        if (e_re_pos >= s) { first = e_rs_rank - 1; last = e_rs_rank; }
        return;
    }
    uint64_t s_rs_rank = s + 1 >= M.size_starts ? M.n_starts : dev_lower_bound(M.starts, M.n_starts, s + 1);
    s_rs_rank = s_rs_rank ? s_rs_rank : 1;
    const uint64_t s_rs_pos = s_rs_rank > M.n_starts ? M.size_starts : __ldg(M.starts + s_rs_rank - 1);
    // run_ends_rank(s): #ends < s, clamped to size-1
    uint64_t s_re_rank = s > M.size_ends - 1 ? dev_lower_bound(M.ends, M.n_ends, M.size_ends - 1)
                                             : dev_lower_bound(M.ends, M.n_ends, s);
    s_re_rank = s_re_rank ? s_re_rank : 1;
    const uint64_t s_re_pos = s_re_rank > M.n_ends ? M.size_ends : __ldg(M.ends + s_re_rank - 1);
    first = s_rs_pos > s_re_pos ? s_rs_rank - 1 : s_rs_rank;
    last = e_rs_rank;
    if (first > last) first = last;
}

// arr_idxs_select (1-based, clamped), rle_window_array.hpp:227-232
__device__ __forceinline__ uint64_t marker_sel(const DevMarkers& M, uint64_t i) {
    return i > M.n_idxs ? M.size_idxs : __ldg(M.idxs + i - 1);
}

#endif  // __CUDACC__

}  // namespace rbg
