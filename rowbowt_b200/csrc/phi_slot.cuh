// The 32-byte phi slot and its decode, shared by locate_kernel and the host-side self-check
// (rbg_selftest_phi) so both read slots identically.
//
// ToeholdSA::phi (include/toehold_sa.hpp:56-72) is a strict-predecessor query over the r sampled
// text positions `pred_` followed by two packed-array reads.  Here text positions are cut into
// buckets of 2^s positions and bucket b IS slot b -- one 32-byte DRAM sector, address computed
// from i alone -- holding everything phi needs for any i in the bucket:
//   bits [  0, 40)  carry key : the largest sampled position < b * 2^s (circular: the last one, n-1)
//   bits [ 40, 80)  carry prev: samples_last[pred_to_run[.] - 1] of that key
//   bits [ 80,248)  up to 3 in-bucket entries, ascending: 16-bit key - b * 2^s, 40-bit prev
//   bits [248,250)  number of in-bucket entries; bit 250: OVERFLOW
// A bucket with more than 3 sampled positions is an OVERFLOW slot: bits [80,120) = first index,
// [120,152) = count of its entries in a side array of (key, prev) pairs that is binary-searched.
// s is chosen at load so that this is rare (layout.cpp).
#pragma once
#include <cstdint>

#include "leaf.cuh"

namespace rbg {

constexpr uint32_t kPhiSlotEntries = 3;
constexpr uint32_t kPhiMaxShift = 16;

template <uint32_t OFF, uint32_t LEN>
RBG_HD uint64_t slot_get(const uint64_t (&q)[4]) {
    constexpr uint32_t wi = OFF >> 6, sh = OFF & 63;
    uint64_t v = q[wi] >> sh;
    if (sh + LEN > 64) v |= q[wi + 1 < 4 ? wi + 1 : 3] << ((64 - sh) & 63);
    return LEN == 64 ? v : v & ((1ull << LEN) - 1);
}
inline void slot_put(uint64_t (&q)[4], uint32_t off, uint32_t len, uint64_t v) {
    for (uint32_t b = 0; b < len; ++b)
        if ((v >> b) & 1) q[(off + b) >> 6] |= 1ull << ((off + b) & 63);
}

RBG_HD bool slot_overflow(const uint64_t (&q)[4]) { return (q[3] >> 58) & 1; }          // bit 250
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

// (prev_sample + delta) % n with delta as in include/toehold_sa.hpp:64
RBG_HD uint64_t phi_value(uint64_t key, uint64_t prev, uint64_t i, uint64_t n) {
    const uint64_t delta = key < i ? i - key : i + 1;
    const uint64_t v = prev + delta;                               // < 2n
    return v >= n ? v - n : v;
}

}  // namespace rbg
