// See layout.hpp.
#include "layout.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "leaf.cuh"

namespace rbg {
namespace {

struct Piece { uint32_t start; uint8_t code; };      // one run clipped to a window, window-relative start

// One NORMAL/TERM line from pieces [i0, i1) (at most 18) with the symbol counts at the first one.
void emit_line(uint32_t* w, const std::vector<Piece>& pc, size_t i0, size_t i1, const uint64_t cnt[4]) {
    if (i1 - i0 > (size_t) kLeafEntries) throw std::logic_error("leaf overflow");
    memset(w, 0, 64);
    for (int c = 0; c < 4; ++c) {
        if (cnt[c] >> 40) throw std::runtime_error("BWT position exceeds 40 bits");
        w[c] = (uint32_t) cnt[c];
        w[4] |= (uint32_t) ((cnt[c] >> 32) & 0xFF) << (8 * c);
    }
    uint16_t st[kLeafEntries];
    for (int e = 0; e < kLeafEntries; ++e) st[e] = (uint16_t) kLeafPad;
    bool term = false;
    for (size_t i = i0; i < i1; ++i) {
        const uint32_t e = (uint32_t) (i - i0);
        uint32_t code = pc[i].code;
        if (code == 4) { code = 0; term = true; }        // terminator rides as an 'A' entry, corrected in the kernel
        st[e] = (uint16_t) pc[i].start;
        if (e < 16) w[5] |= code << leaf_head_bit(e);
        else w[15] |= code << (8 * (e - 16));
    }
    for (int j = 0; j < kLeafEntries / 2; ++j) w[6 + j] = (uint32_t) st[2 * j] | ((uint32_t) st[2 * j + 1] << 16);
    if (term) w[15] |= kModeTerm;
}

// Walks the runs window by window.  out == nullptr: only counts lines.
struct LeafWalker {
    const RunsBwt& bwt;
    const int8_t* code;
    uint32_t g;
    // returns false when some window holds more pieces than one index line can address (18 children)
    bool run(LeafDir* out, uint64_t& n_children, uint64_t& n_split, const uint64_t Fcode[4]) const {
        const uint64_t W = 1ull << g;
        const uint64_t n_direct = (bwt.n + W - 1) >> g;
        uint64_t j = 0, jstart = 0;                 // run covering the current window start
        uint64_t cum[4] = {0, 0, 0, 0};             // symbol counts in BWT[0, jstart)
        n_children = n_split = 0;
        std::vector<Piece> pc;
        for (uint64_t t = 0; t < n_direct; ++t) {
            const uint64_t P = t << g, Pend = std::min(P + W, bwt.n);
            while (jstart + bwt.lens[j] <= P) {
                const int8_t c = code[bwt.heads[j]];
                if (c < 4) cum[c] += bwt.lens[j];
                jstart += bwt.lens[j];
                ++j;
            }
            pc.clear();
            uint64_t st = jstart;
            for (uint64_t i = j; i < bwt.R && st < Pend; st += bwt.lens[i], ++i)
                pc.push_back({(uint32_t) (st > P ? st - P : 0), (uint8_t) code[bwt.heads[i]]});
            const size_t np = pc.size();
            const size_t nchild = (np + kLeafEntries - 1) / kLeafEntries;
            if (nchild > (size_t) kLeafEntries) return false;
            if (out) {
                uint64_t cnt[4];
                for (int c = 0; c < 4; ++c) cnt[c] = Fcode[c] + cum[c];
                const int8_t c0 = code[bwt.heads[j]];
                if (c0 < 4) cnt[c0] += P - jstart;
                uint32_t* w = out->lines.data() + t * kLineWords;
                if (nchild <= 1) {
                    emit_line(w, pc, 0, np, cnt);
                } else {
                    const uint64_t child0 = n_direct + n_children;
                    if ((child0 + nchild) >> 32) throw std::runtime_error("rank directory exceeds 2^32 lines");
                    memset(w, 0, 64);
                    w[0] = (uint32_t) child0;
                    w[15] = kModeSplit;
                    uint16_t cs[kLeafEntries];
                    for (int e = 0; e < kLeafEntries; ++e) cs[e] = (uint16_t) kLeafPad;
                    for (size_t ch = 0; ch < nchild; ++ch) {
                        const size_t i0 = ch * kLeafEntries, i1 = std::min(np, i0 + kLeafEntries);
                        cs[ch] = (uint16_t) pc[i0].start;
                        emit_line(out->lines.data() + (child0 + ch) * kLineWords, pc, i0, i1, cnt);
                        for (size_t i = i0; i < i1; ++i) {           // advance the counts to the next child's start
                            const uint64_t end = i + 1 < np ? pc[i + 1].start : (Pend - P);
                            if (pc[i].code < 4) cnt[pc[i].code] += end - pc[i].start;
                        }
                    }
                    for (int e = 0; e < kLeafEntries / 2; ++e) w[6 + e] = (uint32_t) cs[2 * e] | ((uint32_t) cs[2 * e + 1] << 16);
                }
            }
            if (nchild > 1) { n_children += nchild; ++n_split; }
        }
        return true;
    }
};

}  // namespace

PredTable build_pred_table(std::vector<uint64_t>&& keys, uint64_t universe, double keys_per_bucket) {
    PredTable t;
    t.keys = std::move(keys);
    const double want = std::max(1.0, (double) t.keys.size() / keys_per_bucket);
    uint32_t shift = 0;
    while (shift < 63 && (double) (universe >> shift) > want) ++shift;
    t.shift = shift;
    const uint64_t nb = (universe >> shift) + 2;
    t.table.assign(nb + 1, 0);
    // table[b] = #keys < (b << shift)
    size_t i = 0;
    for (uint64_t b = 0; b <= nb; ++b) {
        const uint64_t lim = b << shift;
        while (i < t.keys.size() && t.keys[i] < lim) ++i;
        t.table[b] = (uint32_t) i;
    }
    if (t.keys.size() >> 32) throw std::runtime_error("more than 2^32 keys in a PredTable");
    return t;
}

LeafDir build_leaf_dir(const RunsBwt& bwt, uint32_t leaf_bits) {
    LeafDir d;
    d.n = bwt.n;
    if (bwt.n == 0 || bwt.R == 0) throw format_error("empty BWT");
    static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
    int8_t code[256];
    memset(code, -1, sizeof code);
    for (int c = 0; c < 4; ++c) code[sym[c]] = (int8_t) c;
    code[1] = 4;
    uint64_t counts256[256] = {0};
    uint64_t pos = 0;
    for (uint64_t j = 0; j < bwt.R; ++j) {
        const uint8_t h = bwt.heads[j];
        if (code[h] < 0)
            throw alphabet_error("BWT contains byte " + std::to_string((int) h) +
                                 ": only {terminator,A,C,G,T} indexes are supported (build with pfbwt-f --non-acgt-to-a)");
        if (bwt.lens[j] == 0) throw format_error("zero-length run");
        if (code[h] == 4)
            for (uint64_t t = 0; t < bwt.lens[j]; ++t) {
                if (d.n_term >= (uint32_t) kMaxTerm) throw alphabet_error("more than 8 terminator symbols in the BWT");
                d.term_pos[d.n_term++] = pos + t;
            }
        counts256[h] += bwt.lens[j];
        pos += bwt.lens[j];
    }
    if (pos != bwt.n) throw format_error("run lengths do not sum to n");
    d.F[0] = 0;                                                     // RowBowt::build_f, include/rowbowt.hpp:770-778
    for (int i = 0; i < 255; ++i) d.F[i + 1] = d.F[i] + counts256[i];
    for (int c = 0; c < 4; ++c) { d.Fcode[c] = d.F[sym[c]]; d.count[c] = counts256[sym[c]]; }
    memset(d.code_of, -1, sizeof d.code_of);
    for (int c = 0; c < 4; ++c) if (d.count[c]) d.code_of[sym[c]] = (int8_t) c;
    if (d.n_term) d.code_of[1] = 4;

    auto total_lines = [&](uint32_t g, uint64_t& children, uint64_t& split) -> uint64_t {
        if (!LeafWalker{bwt, code, g}.run(nullptr, children, split, d.Fcode)) return ~0ull;
        return ((bwt.n + (1ull << g) - 1) >> g) + children;
    };
    uint32_t g = leaf_bits;
    if (g == 0) if (const char* e = getenv("RBG_LEAF_BITS")) g = (uint32_t) atoi(e);
    uint64_t children = 0, split = 0;
    if (g == 0) {
        // aim at ~14 of the 18 entries used on average, then keep the smallest of g-1, g, g+1
        const double avg = (double) bwt.n / (double) bwt.R;
        int guess = (int) std::floor(std::log2(14.0 * avg));
        guess = std::min(kMaxLeafBits, std::max(kMinLeafBits, guess));
        uint64_t best = ~0ull;
        for (int cand = std::max(kMinLeafBits, guess - 1); cand <= std::min(kMaxLeafBits, guess + 1); ++cand) {
            const uint64_t total = total_lines((uint32_t) cand, children, split);
            if (total < best) { best = total; g = (uint32_t) cand; }
        }
        if (best == ~0ull) g = (uint32_t) std::max(kMinLeafBits, guess - 1);
        while (best == ~0ull && g > (uint32_t) kMinLeafBits) best = total_lines(--g, children, split);   // pathological density
    }
    if (g < (uint32_t) kMinLeafBits || g > (uint32_t) kMaxLeafBits) throw std::runtime_error("leaf_bits out of range [4,15]");
    if (total_lines(g, children, split) == ~0ull)
        throw std::runtime_error("leaf_bits too large for this BWT: a window holds more than 324 runs");
    d.g = g;
    d.n_direct = (bwt.n + (1ull << g) - 1) >> g;
    d.lines.assign((d.n_direct + children) * kLineWords, 0);
    LeafWalker{bwt, code, g}.run(&d, children, d.n_split, d.Fcode);
    return d;
}

ToeholdDir build_toehold_dir(const RunsBwt& bwt, const uint64_t (&F)[256], const ToeholdArrays& tsa) {
    if (tsa.r != bwt.R || tsa.n != bwt.n) throw format_error("toehold SA does not match the BWT (r/n differ)");
    ToeholdDir t;
    // LF(end of run j) = F[c] + (#c in BWT[0, end_j]) - 1.  Visiting symbols in byte order and runs in
    // BWT order enumerates these rows in increasing order.
    std::vector<uint64_t> rows(bwt.R), sample(bwt.R);
    uint64_t cnt[256] = {0}, fill[256];
    for (uint64_t j = 0; j < bwt.R; ++j) cnt[bwt.heads[j]]++;
    uint64_t acc = 0;
    for (int c = 0; c < 256; ++c) { fill[c] = acc; acc += cnt[c]; }
    uint64_t seen[256] = {0};
    for (uint64_t j = 0; j < bwt.R; ++j) {
        const uint8_t c = bwt.heads[j];
        seen[c] += bwt.lens[j];
        const uint64_t slot = fill[c]++;
        rows[slot] = F[c] + seen[c] - 1;
        sample[slot] = tsa.samples_last[j];
    }
    t.rows = build_pred_table(std::move(rows), bwt.n, 2.0);
    t.sample = std::move(sample);
    t.toehold0 = (tsa.samples_last[tsa.r - 1] + 1) % tsa.n;     // include/toehold_sa.hpp:97-99
    return t;
}

PhiDir build_phi_dir(const ToeholdArrays& tsa) {
    PhiDir p;
    p.prev.resize(tsa.r);
    for (uint64_t i = 0; i < tsa.r; ++i) {
        const uint64_t run = tsa.pred_to_run[i];
        // pred_to_run == 0 only for phi(SA[0]), which locate_range never evaluates (toehold_sa.hpp:65-66)
        p.prev[i] = run ? tsa.samples_last[run - 1] : 0;
    }
    std::vector<uint64_t> keys = tsa.pred;
    p.pred = build_pred_table(std::move(keys), tsa.n, 2.0);
    return p;
}

}  // namespace rbg
