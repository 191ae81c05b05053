// sm_100a kernels of rb_markers' greedy-seeding marker genotyping:
//   RowBowt::get_markers_greedy_seeding   include/rowbowt.hpp:406-482
//   the worker around it                  src/rb_markers.cpp:347-415 (both strands, sort + unique per seed)
//
// One thread walks one (read, strand).  The reverse strand is never materialised: the backward
// search of revcomp(S) consumes comp(S[0]), comp(S[1]), ... so the thread reads the same 2-bit
// stream forwards and flips the code (A0 C1 G2 T3: complement = code ^ 3).  Forward items fill the
// first half of the grid-stride space and reverse items the second, so that a warp holds one kind:
// forward strands of genuine reads extend for the whole read, reverse strands fail every ~log4(n)
// bases -- mixed in one warp they would serialise each other.
//
// Output sizes are data dependent (seeds per strand, marker words per seed), so the walk runs
// twice: a counting pass, one scan over (read, strand) items for seeds and one for words, and an
// emitting pass that writes every seed record and gathers its marker words at their final
// offsets.  A third kernel sorts / uniques each seed's words in place (they are few).
#include "kernels.cuh"

namespace rbg {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// 2k bits of the packed stream starting at base x0 (k <= 13)
__device__ __forceinline__ uint64_t packed_span(const uint64_t* packed, uint64_t x0, uint32_t k) {
    const uint32_t sh = 2u * (uint32_t) (x0 & 31);
    uint64_t v = __ldg(packed + (x0 >> 5)) >> sh;
    if (sh + 2u * k > 64u) v |= __ldg(packed + (x0 >> 5) + 1) << (64u - sh);
    return v & ((1ull << (2u * k)) - 1);
}
// k bits of the bad-base plane starting at base x0
__device__ __forceinline__ uint32_t bad_span(const uint32_t* bad, uint64_t x0, uint32_t k) {
    const uint32_t sh = (uint32_t) (x0 & 31);
    uint32_t v = __ldg(bad + (x0 >> 5)) >> sh;
    if (sh + k > 32u) v |= __ldg(bad + (x0 >> 5) + 1) << (32u - sh);
    return v & ((1u << k) - 1);
}

// One strand of one read: the state of get_markers_greedy_seeding plus the output cursors.
template <bool EMIT>
struct Walk {
    const DevLeafDir& D;
    const DevFtab& ft;
    const DevMarkers& M;
    const DevBatch& b;
    const GreedyParams& P;
    const DevSeedOut& o;
    uint64_t beg, m;
    bool rev;
    // cached words of the 2-bit stream / bad plane
    uint64_t cw = ~0ull, word = 0;
    uint32_t badw = 0;
    // outputs
    uint64_t seed_base = 0, word_base = 0;      // EMIT: first seed / word slot of this item
    uint64_t word_room = 0;                     // EMIT: word slots of this item
    uint64_t n_seeds = 0, words_done = 0;       // seeds emitted; raw words of finished seeds
    uint64_t seed_words = 0;                    // |mbuf| of the current seed
    unsigned long long steps = 0;
    uint32_t touched = 0;

    __device__ __forceinline__ Walk(const DevLeafDir& D_, const DevFtab& ft_, const DevMarkers& M_, const DevBatch& b_,
                                    const GreedyParams& P_, const DevSeedOut& o_, uint64_t beg_, uint64_t m_, bool rev_)
        : D(D_), ft(ft_), M(M_), b(b_), P(P_), o(o_), beg(beg_), m(m_), rev(rev_) {}

    // Q[j] of the strand string: its 2-bit code, false when the base has none ('N' after seq_ntoa_table)
    __device__ __forceinline__ bool base_at(uint64_t j, uint32_t& c) {
        const uint64_t x = beg + (rev ? m - 1 - j : j);
        if ((x >> 5) != cw) {
            cw = x >> 5;
            word = __ldg(b.packed + cw);
            badw = __ldg(b.bad + cw);
        }
        const uint32_t s = (uint32_t) x & 31u;
        c = ((uint32_t) (word >> (2u * s)) & 3u) ^ (rev ? 3u : 0u);
        return !((badw >> s) & 1u);
    }

    // search_ftab(Q.substr(a, k)), include/rowbowt.hpp:745-758
    __device__ __forceinline__ bool kmer_at(uint64_t a, uint64_t& lo, uint64_t& hi) {
        const uint32_t k = P.k;
        const uint64_t x0 = beg + (rev ? m - a - k : a);
        if (bad_span(b.bad, x0, k)) return false;
        uint64_t key = packed_span(b.packed, x0, k);
        if (rev) {          // complement, then reverse the order of the k 2-bit fields
            uint64_t y = __brevll(~key & ((1ull << (2u * k)) - 1));
            y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
            key = y >> (64u - 2u * k);
        }
        const ulonglong2 e = __ldg(ft.range + key);
        if (e.x > e.y) return false;
        lo = e.x;
        hi = e.y;
        return true;
    }

    // update_mbuf, include/rowbowt.hpp:437-441: append at_range(r) when the range is narrow enough
    __device__ __forceinline__ void update(uint64_t lo, uint64_t hi) {
        if (hi - lo + 1 > P.max_range) return;
        uint64_t first, last;
        marker_windows(M, lo, hi, first, last);
        if (last <= first) return;
        const uint64_t a = marker_sel(M, first + 1), z = marker_sel(M, last + 1);
        if (z <= a) return;
        if (EMIT) {
            // Words of a seed that min_range will drop were not counted: they may only land in slots this item
            // owns (a later seed of the item overwrites them), never beyond them in a neighbour's.
            const uint64_t at = words_done + seed_words;
            uint64_t* dst = o.words + word_base + at;
            for (uint64_t t = 0; t < z - a && at + t < word_room; ++t) dst[t] = __ldg(M.arr + a + t);
        }
        seed_words += z - a;
    }

    // out_fn, src/rb_markers.cpp:356-373, then mbuf.clear()
    __device__ __forceinline__ void emit(uint64_t lo, uint64_t hi, uint64_t qfirst, uint64_t qlast) {
        if (!(hi < lo)) {
            uint64_t raw = seed_words;
            if (!(hi - lo + 1 >= P.min_range && raw)) raw = 0;
            if (EMIT) {
                DevSeed s;
                s.lo = lo;
                s.hi = hi;
                s.mk_off = word_base + words_done;
                s.query_start = (uint32_t) (rev ? m - qfirst - 1 : qfirst);
                s.query_len = (uint32_t) (qlast - qfirst + 1);
                s.mk_raw = (uint32_t) raw;
                s.mk_cnt = (uint32_t) raw;
                o.seeds[seed_base + n_seeds] = s;
            }
            ++n_seeds;
            words_done += raw;
        }
        seed_words = 0;
    }

    __device__ __forceinline__ void run() {
        const uint64_t flo = 0, fhi = D.n - 1;              // full_range
        const uint32_t k = P.k;
        uint64_t plo = flo, phi = fhi, lo = flo, hi = fhi, i = 0;
        if (k) {                                            // :430-433
            if (kmer_at(m - k, lo, hi)) i = k;
            plo = lo;
            phi = hi;
        }
        uint64_t window_ei = m, seed_ei = m;
        for (; i < m; ++i) {
            uint32_t c;
            bool hi_is_c;
            ++steps;
            const bool ok = base_at(m - i - 1, c) && lf_step<false>(D, c, lo, hi, hi_is_c, touched);
            if (!ok) {                                      // :444 the seed fails
                if (seed_ei - (m - i) >= P.wsize) update(plo, phi);
                emit(plo, phi, m - i, seed_ei - 1);
                plo = flo;
                phi = fhi;
                seed_ei = window_ei = m - i - 1;
                lo = flo;
                hi = fhi;
                if (k && m - i - 1 >= k) {
                    // :454-464.  search_ftab answers a miss with the (non-empty) full range, so the reference's
                    // "shift left until a k-mer matches" loop always stops at its first k-mer: a hit continues
                    // from the table entry, a miss skips the k bases and continues from the full range.
                    seed_ei = window_ei = m - i - 1;
                    kmer_at(m - i - 1 - k, lo, hi);
                    i += k;
                    plo = lo;
                    phi = hi;
                }
            } else {                                        // :468-474 window checkpoint
                if (window_ei - (m - i - 1) >= P.wsize) {
                    update(lo, hi);
                    window_ei = m - i - 1;
                }
                plo = lo;
                phi = hi;
            }
        }
        if (hi >= lo && seed_ei - (m - i) >= P.wsize) update(lo, hi);       // :478-480
        emit(lo, hi, m - i, seed_ei - 1);
    }
};

template <bool EMIT>
__global__ void __launch_bounds__(kBlock) greedy_kernel(DevLeafDir D, DevFtab ft, DevMarkers M, DevBatch b, GreedyParams P,
                                                         DevSeedOut o, DevCounters* ctr) {
    unsigned long long steps = 0, lines = 0, words = 0;
    const uint64_t n = b.r1 - b.r0;
    for (uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += (uint64_t) gridDim.x * blockDim.x) {
        const bool rev = t >= n;
        const uint64_t read = b.r0 + (rev ? t - n : t);
        const uint64_t slot = 2 * read + (rev ? 1 : 0);
        const uint64_t beg = b.offs[read];
        Walk<EMIT> w(D, ft, M, b, P, o, beg, b.offs[read + 1] - beg, rev);
        if (EMIT) {
            w.seed_base = o.seed_off[slot];
            w.word_base = o.word_off[slot];
            w.word_room = o.word_off[slot + 1] - w.word_base;
        }
        w.run();
        if (!EMIT) {
            o.item_seeds[slot] = w.n_seeds;
            o.item_words[slot] = w.words_done;
        }
        steps += w.steps;
        lines += w.touched;
        words += w.words_done;
    }
    steps = warp_sum(steps);
    lines = warp_sum(lines);
    words = warp_sum(words);
    if ((threadIdx.x & 31) == 0 && steps) {
        atomicAdd(&ctr->lf_steps, steps);
        atomicAdd(&ctr->lf_lines, lines);
        if (EMIT) atomicAdd(&ctr->marker_words, words);
    }
}

// marker_cmp (src/rb_markers.cpp:243-251) orders by (seq, pos, allele) = the fields of the word
// (pfbwt-f/include/marker.hpp: allele[63:60] seq[59:46] pos[43:0]); ties fall back to the whole word.
__device__ __forceinline__ uint64_t marker_key(uint64_t w) {
    return (((w & 0x0FFFF00000000000ull) >> 46) << 48) | ((w & 0x00000FFFFFFFFFFFull) << 4) | (w >> 60);
}
__device__ __forceinline__ bool marker_before(uint64_t a, uint64_t b) {
    const uint64_t ka = marker_key(a), kb = marker_key(b);
    return ka < kb || (ka == kb && a < b);
}

__device__ void sift_down(uint64_t* w, uint64_t root, uint64_t end) {      // max-heap on marker_before
    for (;;) {
        uint64_t child = 2 * root + 1;
        if (child >= end) return;
        if (child + 1 < end && marker_before(w[child], w[child + 1])) ++child;
        if (!marker_before(w[root], w[child])) return;
        const uint64_t t = w[root];
        w[root] = w[child];
        w[child] = t;
        root = child;
    }
}

// One seed per thread: std::sort(marker_cmp) + std::unique over its (few) words, in place.
__global__ void __launch_bounds__(kBlock) seed_sort_kernel(DevSeedOut o, uint64_t n_seeds) {
    for (uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; s < n_seeds; s += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t raw = o.seeds[s].mk_raw;
        if (raw < 2) continue;
        uint64_t* w = o.words + o.seeds[s].mk_off;
        if (raw <= 24) {
            for (uint32_t i = 1; i < raw; ++i) {
                const uint64_t x = w[i];
                uint32_t j = i;
                while (j > 0 && marker_before(x, w[j - 1])) { w[j] = w[j - 1]; --j; }
                w[j] = x;
            }
        } else {
            for (uint64_t i = raw / 2; i-- > 0;) sift_down(w, i, raw);
            for (uint64_t end = raw - 1; end > 0; --end) {
                const uint64_t t = w[0];
                w[0] = w[end];
                w[end] = t;
                sift_down(w, 0, end);
            }
        }
        uint32_t cnt = 1;
        for (uint32_t i = 1; i < raw; ++i)
            if (w[i] != w[cnt - 1]) w[cnt++] = w[i];
        o.seeds[s].mk_cnt = cnt;
    }
}

}  // namespace

int launch_greedy(const DevLeafDir& D, const DevFtab& ft, const DevMarkers& M, const DevBatch& b, const GreedyParams& P,
                  const DevSeedOut& o, bool emit, DevCounters* ctr, cudaStream_t st) {
    if (b.r1 <= b.r0) return 0;
    const int grid = grid_for(2 * (b.r1 - b.r0), kBlock, 8);
    if (emit) greedy_kernel<true><<<grid, kBlock, 0, st>>>(D, ft, M, b, P, o, ctr);
    else greedy_kernel<false><<<grid, kBlock, 0, st>>>(D, ft, M, b, P, o, ctr);
    return 1;
}

int launch_seed_sort(const DevSeedOut& o, uint64_t n_seeds, cudaStream_t st) {
    if (!n_seeds) return 0;
    seed_sort_kernel<<<grid_for(n_seeds, kBlock, 8), kBlock, 0, st>>>(o, n_seeds);
    return 1;
}

}  // namespace rbg
