// The 32-byte phi slot and its decode, shared by locate_kernel and the host-side self-check
// (rbg_selftest_phi) so both read slots identically.
//
// ToeholdSA::phi (include/toehold_sa.hpp:56-72) is a strict-predecessor query over the r sampled
// text positions `pred_` followed by two packed-array reads.  Here text positions are cut into
// buckets of 2^s positions; every NON-EMPTY bucket has a slot -- one 32-byte sector -- holding everything phi
// needs for any i in the bucket, and a position in an empty bucket is answered by the carry of the next slot.
// Which slot: l1[b / 32] = (bitmap of the non-empty ones among 32 buckets) | (non-empty buckets before) << 32,
// so slot = high + popc(bitmap below b).  Samples cluster around variant sites (9 of 10 buckets of 128 positions
// are empty on the BASELINE index), so the slots take 81 MB instead of the n/4 = 812 MB of one slot per bucket:
// inside the reach of the SM TLBs (256 MB) and mostly L2-resident, where the flat table was TLB-miss bound
// (profiles/: 120 B of DRAM traffic per phi step for 35 B of payload).  The l1 words (n/512 bytes) stay in L2.
// A slot:
//   bits [  0, 40)  carry key : the largest sampled position < b * 2^s (circular: the last one, n-1)
//   bits [ 40, 80)  carry prev: samples_last[pred_to_run[.] - 1] of that key
//   INLINE  (bit 250 = 0): bits [80,248) up to 3 in-bucket entries, ascending: 16-bit key - b * 2^s,
//           40-bit prev; bits [248,250) their number
//   BITMAP  (bits 250,251 = 1,0; s <= 7): bits [120,248) one bit per position of the bucket (sampled or
//           not), bits [80,120) index of the bucket's first sample in a dense array of prev values
//   SEARCH  (bits 250,251 = 1,1; s > 7): bits [80,120) first index, [120,152) count of the bucket's
//           entries in side arrays of (key, prev) that are binary-searched
// Sampled positions come in dense stretches around variant sites (74 % of the gaps are 1 on the
// BASELINE index) with long empty spans between them: about 9 of 10 buckets of 128 positions hold no
// sample at all (the carry answers: ONE sector), nearly all others are BITMAP (popcount, then one more
// sector for prev).  s = 7 unless the slots would not fit the memory budget (layout.cpp).
#pragma once
#include <cstdint>

#include "leaf.cuh"

namespace rbg {

constexpr uint32_t kPhiSlotEntries = 3;
constexpr uint32_t kPhiMaxShift = 16;
constexpr uint32_t kPhiBitmapShift = 7;       // buckets of up to 128 positions can be BITMAP slots

template <uint32_t OFF, uint32_t LEN>
RBG_HD uint64_t slot_get(const uint64_t (&q)[4]) {
    constexpr uint32_t wi = OFF >> 6, sh = OFF & 63;
    uint64_t v = q[wi] >> sh;
    if (sh + LEN > 64) v |= q[wi + 1 < 4 ? wi + 1 : 3] << ((64 - sh) & 63);
    return LEN == 64 ? v : v & ((1ull << LEN) - 1);
}
inline void slot_put(uint64_t (&q)[4], uint32_t off, uint32_t len, uint64_t v) {      // ORs the low `len` bits of v in at bit `off`
    if (len < 64) v &= (1ull << len) - 1;
    const uint32_t wi = off >> 6, sh = off & 63;
    q[wi] |= v << sh;
    if (sh + len > 64 && wi + 1 < 4) q[wi + 1] |= v >> (64 - sh);
}

// Slot index of bucket (group word g, bucket b & 31 = bit) and whether that bucket holds a sample itself.
RBG_HD uint64_t phi_slot_index(uint64_t g, uint32_t bit, bool& here) {
    const uint32_t bm = (uint32_t) g;
    here = (bm >> bit) & 1u;
    return (g >> 32) + rbg_popc(bm & ((1u << bit) - 1u));
}

RBG_HD bool slot_overflow(const uint64_t (&q)[4]) { return (q[3] >> 58) & 1; }          // bit 250: BITMAP or SEARCH
RBG_HD bool slot_search(const uint64_t (&q)[4]) { return (q[3] >> 59) & 1; }            // bit 251
RBG_HD uint32_t slot_count(const uint64_t (&q)[4]) { return (uint32_t) (q[3] >> 56) & 3u; }   // bits 248..249
RBG_HD uint64_t slot_ovf_start(const uint64_t (&q)[4]) { return slot_get<80, 40>(q); }
RBG_HD uint32_t slot_ovf_count(const uint64_t (&q)[4]) { return (uint32_t) slot_get<120, 32>(q); }

// Strict predecessor of text position i = base + rel within a regular slot: (key, prev).
RBG_HD void slot_pred(const uint64_t (&q)[4], uint64_t base, uint32_t rel, uint64_t& key, uint64_t& prev) {
    key = slot_get<0, 40>(q);
    prev = slot_get<40, 40>(q);
    const uint32_t cnt = slot_count(q);
    const uint32_t r0 = (uint32_t) slot_get<80, 16>(q), r1 = (uint32_t) slot_get<136, 16>(q), r2 = (uint32_t) slot_get<192, 16>(q);
    if (cnt > 0 && r0 < rel) { key = base + r0; prev = slot_get<96, 40>(q); }
    if (cnt > 1 && r1 < rel) { key = base + r1; prev = slot_get<152, 40>(q); }
    if (cnt > 2 && r2 < rel) { key = base + r2; prev = slot_get<208, 40>(q); }
}

// BITMAP slot: is there a sample below `rel`?  If so its key and its index in the dense prev array.
RBG_HD bool slot_bitmap_pred(const uint64_t (&q)[4], uint64_t base, uint32_t rel, uint64_t& key, uint64_t& idx) {
    const uint64_t lo = slot_get<120, 64>(q), hi = slot_get<184, 64>(q);
    const uint64_t mlo = rel >= 64 ? lo : lo & ((1ull << rel) - 1);
    const uint64_t mhi = rel > 64 ? hi & ((1ull << ((rel - 64) & 63)) - 1) : 0;       // rel <= 128 - 1
    if ((mlo | mhi) == 0) return false;
#if defined(__CUDA_ARCH__)
    const uint32_t below = (uint32_t) (__popcll(mlo) + __popcll(mhi));
    const uint32_t top = mhi ? 127u - (uint32_t) __clzll((long long) mhi) : 63u - (uint32_t) __clzll((long long) mlo);
#else
    const uint32_t below = (uint32_t) (__builtin_popcountll(mlo) + __builtin_popcountll(mhi));
    const uint32_t top = mhi ? 127u - (uint32_t) __builtin_clzll(mhi) : 63u - (uint32_t) __builtin_clzll(mlo);
#endif
    key = base + top;
    idx = slot_ovf_start(q) + below - 1;
    return true;
}

// (prev_sample + delta) % n with delta as in include/toehold_sa.hpp:64
RBG_HD uint64_t phi_value(uint64_t key, uint64_t prev, uint64_t i, uint64_t n) {
    const uint64_t delta = key < i ? i - key : i + 1;
    const uint64_t v = prev + delta;                               // < 2n
    return v >= n ? v - n : v;
}

}  // namespace rbg
