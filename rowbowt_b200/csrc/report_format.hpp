// rb_report's text (src/rb_align.cpp:118-145) written at memory speed.
//
// With -s the report is ~1.4 KB per read of the BASELINE workload (13.9 GB for 10 M reads): one "<loc>/<doc>:<off> "
// item per occurrence.  The first version built it with std::string appends and a digit-at-a-time put_u64 and spent
// 228 ns per item (profiles/r2_e2e_binaries_c2_10m.json, RBG_HOST_STATS: 129 s of formatter time for 565 M items).
// Here a slice is written through a raw pointer into a buffer that is sized once from the result's offsets (no
// capacity checks in the loops), numbers go out two digits at a time behind a length computed from the leading-zero
// count, the document of a location comes from a bucket table over the document starts instead of a binary search,
// and its "/name:" text is precomposed.  Output bytes are identical (tests/test_host_cpu.py compares both writers).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rowbowt_gpu.h"

namespace rbhost {

struct OutBuf {                     // grows, never shrinks: recycled with its batch, so the pages stay mapped
    char* p = nullptr;
    size_t len = 0, cap = 0;
    OutBuf() = default;
    OutBuf(const OutBuf&) = delete;
    OutBuf& operator=(const OutBuf&) = delete;
    OutBuf(OutBuf&& o) noexcept : p(o.p), len(o.len), cap(o.cap) { o.p = nullptr; o.len = o.cap = 0; }
    OutBuf& operator=(OutBuf&& o) noexcept {
        if (this != &o) { free(p); p = o.p; len = o.len; cap = o.cap; o.p = nullptr; o.len = o.cap = 0; }
        return *this;
    }
    ~OutBuf() { free(p); }
    void reserve(size_t want) {
        if (want <= cap) return;
        free(p);                     // contents are never kept across a reserve
        cap = want + want / 8 + 4096;
        p = (char*) malloc(cap);
        if (!p) { fprintf(stderr, "out of memory (%zu bytes of report text)\n", cap); exit(1); }
        len = 0;
    }
};

namespace fmt_detail {
struct Digits2 {
    char d[200];
    constexpr Digits2() : d() {
        for (int i = 0; i < 100; ++i) { d[2 * i] = (char) ('0' + i / 10); d[2 * i + 1] = (char) ('0' + i % 10); }
    }
};
static constexpr Digits2 kDigits2{};
static constexpr uint64_t kPow10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                        1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                        100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                                        1000000000000000000ull, 10000000000000000000ull};
}  // namespace fmt_detail

// decimal digits of v (1..20)
inline unsigned dec_len(uint64_t v) {
    const unsigned bits = 64u - (unsigned) __builtin_clzll(v | 1);
    const unsigned t = (bits * 1233u) >> 12;                      // floor(log10(2^bits)) or one less
    return t + ((v | 1) >= fmt_detail::kPow10[t] ? 1u : 0u);       // (v | 1: zero has one digit)
}

// writes v in decimal at o, returns the position behind it
inline char* put_dec(char* o, uint64_t v) {
    const unsigned n = dec_len(v);
    char* p = o + n;
    while (v >= 100) {
        const uint64_t q = v / 100;
        const unsigned r = (unsigned) (v - q * 100);
        v = q;
        p -= 2;
        memcpy(p, fmt_detail::kDigits2.d + 2 * r, 2);
    }
    if (v >= 10) memcpy(p - 2, fmt_detail::kDigits2.d + 2 * v, 2);
    else p[-1] = (char) ('0' + v);
    return o + n;
}
constexpr size_t kMaxDec = 20;

// DocList::doc_and_offset_at (include/doclist.hpp:46-50) for many locations: rank = #starts <= i from a bucket table
// (at most a few comparisons), the document's text precomposed as "/<name>:".
class DocResolver {
  public:
    // names/starts exactly as rbhost::DocList holds them (starts sorted + uniqued, names in file order: the reference's
    // own pairing, kept as is)
    void init(const std::vector<std::string>& names, const std::vector<uint64_t>& starts) {
        starts_ = starts;
        text_.clear();
        max_text_ = 0;
        for (const std::string& n : names) {
            text_.push_back("/" + n + ":");
            max_text_ = std::max(max_text_, text_.back().size());
        }
        n_names_ = names.size();
        const uint64_t last = starts_.empty() ? 0 : starts_.back();
        shift_ = 0;
        while ((last >> shift_) >= (1u << 16)) ++shift_;
        table_.assign((size_t) (last >> shift_) + 2, 0);
        size_t k = 0;
        for (size_t b = 0; b < table_.size(); ++b) {               // table[b] = #starts < (b << shift)
            const uint64_t lim = (uint64_t) b << shift_;
            while (k < starts_.size() && starts_[k] < lim) ++k;
            table_[b] = (uint32_t) k;
        }
    }
    size_t max_text() const { return max_text_; }
    // the text "/<name>:" of the document of location i and the offset inside it
    inline const std::string& resolve(uint64_t i, uint64_t& off) const {
        const uint64_t b = i >> shift_;
        size_t rank = b + 1 < table_.size() ? table_[b] : starts_.size();
        while (rank < starts_.size() && starts_[rank] <= i) ++rank;          // rank = #starts <= i
        if (rank == 0) rank = 1;
        off = i - starts_[rank - 1];
        return text_[std::min(rank, n_names_) - 1];
    }

  private:
    std::vector<uint64_t> starts_;
    std::vector<std::string> text_;
    std::vector<uint32_t> table_;
    size_t n_names_ = 0, max_text_ = 0;
    uint32_t shift_ = 0;
};

inline uint64_t report_marker_pos(uint64_t m) { return m & 0x00000FFFFFFFFFFFull; }            // MarkerT, pfbwt-f/include/marker.hpp:19-21,35-37
inline uint64_t report_marker_allele(uint64_t m) { return (m & 0xF000000000000000ull) >> 60; }

// Reads [i0, i1) of one result.  `name_of(i, len)` returns the read's name (cut at a NUL: printed as a C string).
template <class NameOf>
void format_report(bool sam, bool markers, const DocResolver& docs, const rbg_result& r, NameOf name_of, uint64_t i0, uint64_t i1,
                   OutBuf& out) {
    static const char kNoMarkers[] = "no markers (consider building the marker array with a larger window size)";
    size_t names = 0;
    for (uint64_t i = i0; i < i1; ++i) { size_t nl; name_of(i, nl); names += nl; }
    size_t want = names + (i1 - i0) * (2 + 1 + 9 + 1 + 3 * kMaxDec);
    if (sam) want += (r.loc_off[i1] - r.loc_off[i0]) * (2 * kMaxDec + docs.max_text() + 1) + (i1 - i0) * 8;
    if (markers) want += (r.mk_off[i1] - r.mk_off[i0]) * (2 * kMaxDec + 2) + (i1 - i0) * (11 + sizeof kNoMarkers);
    out.reserve(want);
    char* o = out.p;
    for (uint64_t i = i0; i < i1; ++i) {
        size_t nl;
        const char* nm = name_of(i, nl);
        memcpy(o, nm, nl);
        o += nl;
        const uint64_t lo = r.lo ? r.lo[i] : r.lo32[i], hi = r.hi ? r.hi[i] : r.hi32[i];      // RBG_NARROW_RANGES: u32 planes
        memcpy(o, " (", 2);
        o = put_dec(o + 2, lo);
        *o++ = ',';
        o = put_dec(o, hi);
        memcpy(o, "), count=", 9);
        o = put_dec(o + 9, hi - lo + 1);                           // 64-bit wraparound, as printed by the reference
        *o++ = '\n';
        if (sam) {
            memcpy(o, "\tlocs: ", 7);
            o += 7;
            for (uint64_t j = r.loc_off[i]; j < r.loc_off[i + 1]; ++j) {
                // RBG_NARROW_LOCS: a u32 plane plus, for an index with n > 2^32, a u8 plane; u64 otherwise
                const uint64_t loc = r.locs_lo32 ? (uint64_t) r.locs_lo32[j] | (r.locs_hi8 ? (uint64_t) r.locs_hi8[j] << 32 : 0ull) : r.locs[j];
                uint64_t off;
                const std::string& dn = docs.resolve(loc, off);
                o = put_dec(o, loc);
                memcpy(o, dn.data(), dn.size());
                o = put_dec(o + dn.size(), off);
                *o++ = ' ';
            }
            *o++ = '\n';
        }
        if (markers) {
            memcpy(o, "\tmarkers: ", 10);
            o += 10;
            if (r.mk_off[i] == r.mk_off[i + 1]) { memcpy(o, kNoMarkers, sizeof kNoMarkers - 1); o += sizeof kNoMarkers - 1; }
            for (uint64_t j = r.mk_off[i]; j < r.mk_off[i + 1]; ++j) {
                o = put_dec(o, report_marker_pos(r.markers[j]));
                *o++ = '/';
                o = put_dec(o, report_marker_allele(r.markers[j]));
                *o++ = ' ';
            }
            *o++ = '\n';
        }
    }
    out.len = (size_t) (o - out.p);
}

}  // namespace rbhost
