// See layout.hpp.
#include "layout.hpp"
#include "leaf.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace rbg {
namespace {

struct Piece { uint64_t start, len; };   // one c-run, absolute BWT coordinates

// Smallest k (0..s-kMinLeafBits) such that no child of 2^(s-k) positions intersects more
// than kRunEntries of the bucket's pieces; children of 256 positions may overflow (BITS).
// pieces: clipped to the bucket, offsets relative to the bucket start.
uint32_t choose_split(const std::vector<std::pair<uint32_t, uint32_t>>& pieces, uint32_t s,
                      std::vector<uint32_t>& scratch) {
    const uint32_t kmax = s - kMinLeafBits;
    uint32_t k = 0;
    while ((pieces.size() + kRunEntries - 1) / kRunEntries > (1ull << k) && k < kmax) ++k;
    for (; k < kmax; ++k) {
        const uint32_t g = s - k;
        scratch.assign(1u << k, 0);
        bool ok = true;
        for (const auto& p : pieces) {
            uint32_t a = p.first >> g, b = (p.first + p.second - 1) >> g;
            for (uint32_t t = a; t <= b; ++t)
                if (++scratch[t] > (uint32_t) kRunEntries) { ok = false; break; }
            if (!ok) break;
        }
        if (ok) return k;
    }
    return kmax;
}

// Streams the c-runs of one symbol through the bucket grid.  With emit == false only counts lines.
struct SymbolDirBuilder {
    const std::vector<Piece>& runs;
    uint64_t n;
    uint32_t s;

    // line_off: index of this symbol's first line in the global pool; f_c: F[c], folded into
    // every leaf header so that an LF step needs no separate F lookup.
    uint64_t build(bool emit, uint32_t* table, std::vector<uint32_t>* lines, uint64_t line_off, uint64_t f_c) const {
        const uint64_t nb = (n + (1ull << s) - 1) >> s;
        const uint64_t bsz = 1ull << s;
        size_t cur = 0;                 // first run not entirely before the current bucket
        uint64_t cum = 0;               // #c in BWT[0, start of runs[cur])
        uint64_t n_lines = 0;
        std::vector<std::pair<uint32_t, uint32_t>> pieces;
        std::vector<uint32_t> scratch;
        for (uint64_t b = 0; b < nb; ++b) {
            const uint64_t P = b << s, Pend = P + bsz;
            while (cur < runs.size() && runs[cur].start + runs[cur].len <= P) { cum += runs[cur].len; ++cur; }
            // count of c before P: whole runs before cur, plus the part of runs[cur] left of P
            uint64_t cum_at_P = cum;
            if (cur < runs.size() && runs[cur].start < P) cum_at_P += P - runs[cur].start;
            pieces.clear();
            for (size_t i = cur; i < runs.size() && runs[i].start < Pend; ++i) {
                uint64_t a = std::max(runs[i].start, P), e = std::min(runs[i].start + runs[i].len, Pend);
                pieces.emplace_back((uint32_t) (a - P), (uint32_t) (e - a));
            }
            const uint32_t k = choose_split(pieces, s, scratch);
            const uint32_t g = s - k;
            const uint64_t nleaf = 1ull << k;
            if (emit) {
                if ((line_off + n_lines + nleaf) >> 28) throw std::runtime_error("rank directory exceeds 2^28 lines");
                table[b] = (uint32_t) ((line_off + n_lines) << 4) | k;
                uint32_t* out = lines->data() + (line_off + n_lines) * kLineWords;
                memset(out, 0, nleaf * kLineWords * sizeof(uint32_t));
                size_t pi = 0;
                uint64_t running = f_c + cum_at_P;
                for (uint64_t t = 0; t < nleaf; ++t) {
                    uint32_t* w = out + t * kLineWords;
                    const uint32_t L0 = (uint32_t) (t << g), L1 = L0 + (1u << g);
                    // pieces intersecting this leaf: [pi, pj)
                    while (pi < pieces.size() && pieces[pi].first + pieces[pi].second <= L0) ++pi;
                    size_t pj = pi;
                    while (pj < pieces.size() && pieces[pj].first < L1) ++pj;
                    w[0] = (uint32_t) running;
                    uint32_t mode = (pj - pi) > (size_t) kRunEntries ? kBits : kRuns;
                    w[1] = (uint32_t) ((running >> 32) & 0xFF) | (mode << 8);
                    if (running >> 40) throw std::runtime_error("BWT position exceeds 40 bits");
                    uint64_t in_leaf = 0;
                    for (size_t i = pi; i < pj; ++i) {
                        uint32_t a = std::max(pieces[i].first, L0), e = std::min(pieces[i].first + pieces[i].second, L1);
                        if (mode == kRuns) {
                            w[2 + (i - pi)] = ((e - a) << 16) | (a - L0);
                        } else {
                            for (uint32_t p = a - L0; p < e - L0; ++p) w[2 + (p >> 5)] |= 1u << (p & 31);
                        }
                        in_leaf += e - a;
                    }
                    running += in_leaf;
                }
            }
            n_lines += nleaf;
        }
        return n_lines;
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

RankDir build_rank_dir(const RunsBwt& bwt, uint32_t bucket_bits) {
    RankDir d;
    d.n = bwt.n;
    if (bwt.n == 0 || bwt.R == 0) throw format_error("empty BWT");
    static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
    int8_t code[256];
    memset(code, -1, sizeof code);
    for (int c = 0; c < 4; ++c) code[sym[c]] = (int8_t) c;
    code[1] = 4;

    std::vector<Piece> runs[4];
    uint64_t counts256[256] = {0};
    uint64_t pos = 0;
    for (uint64_t j = 0; j < bwt.R; ++j) {
        const uint8_t h = bwt.heads[j];
        const int8_t c = code[h];
        if (c < 0)
            throw alphabet_error("BWT contains byte " + std::to_string((int) h) +
                                 ": only {terminator,A,C,G,T} indexes are supported (build with pfbwt-f --non-acgt-to-a)");
        if (c == 4) {
            for (uint64_t t = 0; t < bwt.lens[j]; ++t) {
                if (d.n_term >= (uint32_t) kMaxTerm) throw alphabet_error("more than 8 terminator symbols in the BWT");
                d.term_pos[d.n_term++] = pos + t;
            }
        } else {
            // adjacent runs of one symbol cannot occur in a run-length BWT, but merge defensively
            if (!runs[c].empty() && runs[c].back().start + runs[c].back().len == pos) runs[c].back().len += bwt.lens[j];
            else runs[c].push_back({pos, bwt.lens[j]});
        }
        counts256[h] += bwt.lens[j];
        pos += bwt.lens[j];
    }
    if (pos != bwt.n) throw format_error("run lengths do not sum to n");
    // RowBowt::build_f, include/rowbowt.hpp:770-778
    d.F[0] = 0;
    for (int i = 0; i < 255; ++i) d.F[i + 1] = d.F[i] + counts256[i];
    for (int c = 0; c < 4; ++c) { d.Fcode[c] = d.F[sym[c]]; d.count[c] = counts256[sym[c]]; }
    memset(d.code_of, -1, sizeof d.code_of);
    for (int c = 0; c < 4; ++c) if (d.count[c]) d.code_of[sym[c]] = (int8_t) c;
    if (d.n_term) d.code_of[1] = 4;

    // bucket size: aim at ~8 runs of each symbol per bucket, then keep the cheapest of s-1, s, s+1
    auto total_lines = [&](uint32_t s) {
        uint64_t t = 0;
        for (int c = 0; c < 4; ++c) t += SymbolDirBuilder{runs[c], d.n, s}.build(false, nullptr, nullptr, 0, 0);
        return t;
    };
    uint32_t s = bucket_bits;
    if (s == 0) {
        if (const char* e = getenv("RBG_BUCKET_BITS")) s = (uint32_t) atoi(e);
    }
    if (s == 0) {
        const double avg = (double) bwt.n / (double) bwt.R;
        int guess = (int) std::lround(std::log2(32.0 * avg));
        guess = std::min(kMaxLeafBits, std::max(kMinLeafBits, guess));
        uint64_t best = ~0ull;
        for (int cand = std::max(kMinLeafBits, guess - 1); cand <= std::min(kMaxLeafBits, guess + 1); ++cand) {
            const uint64_t nb = (d.n + (1ull << cand) - 1) >> cand;
            const uint64_t bytes = total_lines((uint32_t) cand) * 64 + nb * 16;
            if (bytes < best) { best = bytes; s = (uint32_t) cand; }
        }
    }
    if (s < (uint32_t) kMinLeafBits || s > (uint32_t) kMaxLeafBits) throw std::runtime_error("bucket_bits out of range [8,15]");
    d.s = s;
    d.n_buckets = (d.n + (1ull << s) - 1) >> s;
    d.table.assign(4 * d.n_buckets, 0);
    uint64_t per[4], tot = 0;
    for (int c = 0; c < 4; ++c) {
        per[c] = SymbolDirBuilder{runs[c], d.n, s}.build(false, nullptr, nullptr, 0, 0);
        d.line_base[c] = tot;
        tot += per[c];
    }
    d.lines.assign(tot * kLineWords, 0);
    for (int c = 0; c < 4; ++c)
        SymbolDirBuilder{runs[c], d.n, s}.build(true, d.table.data() + c * d.n_buckets, &d.lines, d.line_base[c], d.Fcode[c]);
    return d;
}

// ---------------------------------------------------------------------------------------------
// Layout v2 (mixed leaves, leaf.cuh)
namespace {

struct MixPiece { uint32_t start; uint8_t code; };      // start relative to the direct leaf

// #pieces intersecting [a,b) of the leaf: the one covering a plus those starting inside (a,b).
inline uint32_t pieces_in(const std::vector<MixPiece>& pc, uint32_t a, uint32_t b) {
    auto lt = [](const MixPiece& p, uint32_t x) { return p.start < x; };
    const auto i0 = std::lower_bound(pc.begin(), pc.end(), a + 1, lt);
    const auto i1 = std::lower_bound(pc.begin(), pc.end(), b, lt);
    return (uint32_t) (i1 - i0) + 1u;
}

// Smallest k such that every child of 2^(g-k) positions holds at most kMixEntries pieces.
inline uint32_t mix_split_k(const std::vector<MixPiece>& pc, uint32_t g) {
    for (uint32_t k = 1; k + kMixMinBits <= g; ++k) {
        const uint32_t cs = 1u << (g - k);
        bool ok = true;
        for (uint32_t t = 0; t < (1u << k) && ok; ++t) ok = pieces_in(pc, t * cs, (t + 1) * cs) <= (uint32_t) kMixEntries;
        if (ok) return k;
    }
    return g - kMixMinBits;      // 16-position children always fit
}

// One line for leaf-relative window [a, a+size) given the leaf's pieces and the symbol counts at a.
void mix_emit(uint32_t* w, const std::vector<MixPiece>& pc, uint32_t a, uint32_t size, const uint64_t cnt[4]) {
    for (int c = 0; c < 4; ++c) {
        if (cnt[c] >> 40) throw std::runtime_error("BWT position exceeds 40 bits");
        w[c] = (uint32_t) cnt[c];
    }
    w[4] = 0;
    for (int c = 0; c < 4; ++c) w[4] |= (uint32_t) ((cnt[c] >> 32) & 0xFF) << (8 * c);
    uint16_t e[kMixEntries];
    for (int i = 0; i < kMixEntries; ++i) e[i] = (uint16_t) size;           // padding: covers nothing
    auto lt = [](const MixPiece& p, uint32_t x) { return p.start < x; };
    size_t i = (size_t) (std::lower_bound(pc.begin(), pc.end(), a + 1, lt) - pc.begin()) - 1;   // piece covering a
    int k = 0;
    for (; i < pc.size() && pc[i].start < a + size; ++i, ++k) {
        if (k >= kMixEntries) throw std::logic_error("mixed leaf overflow");
        const uint32_t st = pc[i].start > a ? pc[i].start - a : 0;
        e[k] = (uint16_t) ((uint32_t) pc[i].code << kMixHeadShift | st);
    }
    for (int j = 0; j < kMixEntries / 2; ++j) w[5 + j] = (uint32_t) e[2 * j] | ((uint32_t) e[2 * j + 1] << 16);
}

// Walks the runs leaf by leaf.  emit == nullptr: only counts (direct, overflow, split) lines.
struct MixWalker {
    const RunsBwt& bwt;
    const int8_t* code;
    uint32_t g;
    void run(MixDir* out, uint64_t& n_overflow, uint64_t& n_split, const uint64_t Fcode[4]) const {
        const uint64_t LS = 1ull << g;
        const uint64_t n_direct = (bwt.n + LS - 1) >> g;
        uint64_t j = 0, jstart = 0;                 // run covering the current leaf start
        uint64_t cum[4] = {0, 0, 0, 0};             // symbol counts in BWT[0, jstart)
        n_overflow = n_split = 0;
        std::vector<MixPiece> pc;
        for (uint64_t t = 0; t < n_direct; ++t) {
            const uint64_t P = t << g, Pend = std::min(P + LS, bwt.n);
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
            const bool split = pc.size() > (size_t) kMixEntries;
            uint32_t k = 0;
            if (split) {
                k = mix_split_k(pc, g);
                ++n_split;
            }
            if (out) {
                uint64_t cnt[4];
                for (int c = 0; c < 4; ++c) cnt[c] = Fcode[c] + cum[c];
                const int8_t c0 = code[bwt.heads[j]];
                if (c0 < 4) cnt[c0] += P - jstart;
                uint32_t* w = out->lines.data() + t * kLineWords;
                if (!split) {
                    mix_emit(w, pc, 0, (uint32_t) LS, cnt);
                } else {
                    const uint64_t child0 = n_direct + n_overflow;
                    if ((child0 + (1ull << k)) >> 32) throw std::runtime_error("rank directory exceeds 2^32 lines");
                    memset(w, 0, 64);
                    w[5] = kMixSplit;
                    w[6] = (uint32_t) child0;
                    w[7] = k;
                    const uint32_t cs = 1u << (g - k);
                    size_t pi = 0;                  // piece covering `cur`
                    uint32_t cur = 0;               // cnt[] = symbol counts at leaf position cur
                    for (uint32_t ch = 0; ch < (1u << k); ++ch) {
                        const uint32_t a = ch * cs;
                        while (cur < a) {
                            const uint32_t end = pi + 1 < pc.size() ? pc[pi + 1].start : (uint32_t) LS;
                            const uint32_t step = std::min(end, a) - cur;
                            if (pc[pi].code < 4) cnt[pc[pi].code] += step;
                            cur += step;
                            if (cur == end && pi + 1 < pc.size()) ++pi;
                        }
                        mix_emit(out->lines.data() + (child0 + ch) * kLineWords, pc, a, cs, cnt);
                    }
                }
            }
            if (split) n_overflow += 1ull << k;
        }
    }
};

}  // namespace

MixDir build_mix_dir(const RunsBwt& bwt, uint32_t leaf_bits) {
    MixDir d;
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

    uint32_t g = leaf_bits;
    if (g == 0) if (const char* e = getenv("RBG_LEAF_BITS")) g = (uint32_t) atoi(e);
    if (g == 0) {
        // aim at ~14 of the 22 entries used on average, then keep the smallest of g-1, g, g+1
        const double avg = (double) bwt.n / (double) bwt.R;
        int guess = (int) std::floor(std::log2(14.0 * avg));
        guess = std::min(kMixMaxBits, std::max(kMixMinBits, guess));
        uint64_t best = ~0ull;
        for (int cand = std::max(kMixMinBits, guess - 1); cand <= std::min(kMixMaxBits, guess + 1); ++cand) {
            uint64_t ovf, ns;
            MixWalker{bwt, code, (uint32_t) cand}.run(nullptr, ovf, ns, d.Fcode);
            const uint64_t total = ((bwt.n + (1ull << cand) - 1) >> cand) + ovf;
            if (total < best) { best = total; g = (uint32_t) cand; }
        }
    }
    if (g < (uint32_t) kMixMinBits || g > (uint32_t) kMixMaxBits) throw std::runtime_error("leaf_bits out of range [4,12]");
    d.g = g;
    d.n_direct = (bwt.n + (1ull << g) - 1) >> g;
    uint64_t ovf, ns;
    MixWalker{bwt, code, g}.run(nullptr, ovf, ns, d.Fcode);
    d.lines.assign((d.n_direct + ovf) * kLineWords, 0);
    MixWalker{bwt, code, g}.run(&d, ovf, d.n_split, d.Fcode);
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
