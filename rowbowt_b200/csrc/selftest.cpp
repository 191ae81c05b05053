// Host-side consistency check of the load-time re-layout (diagnostic; not on the query path):
// rebuilds the rank directory from <prefix>.rbwt and compares rank_c at p and p+1 (hence
// BWT[p]==c) decoded from the 64-byte lines -- the same leaf.cuh code the kernels run, cluster
// windows with their raw children and the terminator correction included -- with a direct
// count over the runs.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rowbowt_gpu.h"
#include "formats.hpp"
#include "host_pack.hpp"
#include "layout.hpp"
#include "leaf.cuh"
#include "sdsl_writer.hpp"

using namespace rbg;

namespace {
// F[c] + rank_c(pos) through the directory, pos in [0, n]; mirrors lf_step / leaf_rank_slow
// (device_index.cuh): rank at pos is taken in the window of position `at` with offset q.
uint64_t dir_rank(const LeafDir& d, uint32_t c, uint64_t pos) {
    const bool use_prev = pos == d.n || (pos % d.window == 0 && pos > 0 && (pos & 1));   // also exercise the "hi" form (q = W)
    const uint64_t at = use_prev ? pos - 1 : pos;
    const uint64_t widx = (uint64_t) (((__uint128_t) at * d.magic) >> 64);
    if (widx != at / d.window) return ~0ull;
    const uint32_t q = (uint32_t) (at - widx * d.window) + (use_prev ? 1u : 0u);
    uint32_t w[16];
    memcpy(w, d.lines.data() + widx * 16, 64);
    const bool v5 = d.version == 5;
    uint32_t rel = v5 ? leaf_rel_count<5>(w, c) : leaf_rel_count<4>(w, c);
    uint32_t r = v5 ? leaf_rank<5>(w, leaf_cpat(c), q) : leaf_rank<4>(w, leaf_cpat(c), q);
    uint64_t from = pos - q;
    if (leaf_inside_cluster(w, q)) {
        const uint32_t s = leaf_cluster_begin(w);
        const uint32_t ch = (q - s) / kRawSymbols, p = (q - s) - ch * kRawSymbols;
        const uint32_t* cw = d.lines.data() + ((uint64_t) leaf_child_ptr(w) + ch) * 16;
        rel = (v5 ? rel : 0u) + raw_rel_count(cw, c);             // layout 5: child counts are relative to the window start
        r = raw_rank(cw, leaf_cpat(c), p);
        from += s + ch * kRawSymbols;
    }
    if ((w[15] & kFlagTerm) && c == 0)
        for (uint32_t t = 0; t < d.n_term; ++t) r -= (d.term_pos[t] >= from && d.term_pos[t] < pos) ? 1 : 0;
    return d.super[(uint64_t) c * d.n_super + (widx >> d.sb_shift)] + (uint64_t) rel + r;
}

// The walks below visit every run / sample of a full-size index (10^8 checks): they are cut into contiguous parts, one per
// host thread.  fn(part, begin, end) -> 0 or its failure code; the smallest failure code of any part is returned.
template <class Fn>
int parallel_parts(uint64_t n_items, unsigned n_parts, Fn&& fn) {
    n_parts = (unsigned) std::max<uint64_t>(1, std::min<uint64_t>(n_parts, n_items ? n_items : 1));
    std::vector<int> rc(n_parts, 0);
    std::vector<std::thread> th;
    for (unsigned p = 0; p < n_parts; ++p)
        th.emplace_back([&, p] {
            try { rc[p] = fn(p, n_items * p / n_parts, n_items * (p + 1) / n_parts); } catch (...) { rc[p] = -1; }
        });
    for (auto& t : th) t.join();
    int out = 0;
    for (int r : rc) if (r != 0 && (out == 0 || r < out)) out = r;
    return out;
}
unsigned selftest_threads() { return std::max(1u, std::min(32u, std::thread::hardware_concurrency())); }
}  // namespace

extern "C" int rbg_selftest_layout(const char* prefix, uint32_t window, uint64_t stride, uint64_t* checked,
                                   uint64_t* n_lines, uint64_t* n_cluster) {
    try {
        RunsBwt bwt = read_rbwt(std::string(prefix) + ".rbwt");
        const uint32_t layout = window >> 16;                     // bits 16..: force layout 4 or 5 (0 = the loader's own choice)
        window &= 0xFFFFu;
        if (layout) setenv("RBG_LAYOUT", layout == 5 ? "5" : "4", 1);
        LeafDir d = build_leaf_dir(bwt, window);
        if (layout) unsetenv("RBG_LAYOUT");
        if (layout && d.version != (int) layout) return 3;
        if (n_lines) *n_lines = d.n_lines();
        if (n_cluster) *n_cluster = d.n_cluster;
        static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
        if (stride == 0) stride = 1;
        int8_t code_of_head[256];
        memset(code_of_head, -1, sizeof code_of_head);
        for (int c = 0; c < 4; ++c) code_of_head[sym[c]] = (int8_t) c;
        // where every part starts: position and per-symbol counts in front of its first run
        const unsigned n_parts = selftest_threads();
        struct Start { uint64_t pos, cum[4]; };
        std::vector<Start> start(n_parts + 1, Start{0, {0, 0, 0, 0}});
        {
            Start cur{0, {0, 0, 0, 0}};
            unsigned p = 0;
            for (uint64_t j = 0; j <= bwt.R; ++j) {
                while (p <= n_parts && j == bwt.R * p / n_parts) start[p++] = cur;
                if (j == bwt.R) break;
                const int hc = code_of_head[bwt.heads[j]];
                if (hc >= 0) cur.cum[hc] += bwt.lens[j];
                cur.pos += bwt.lens[j];
            }
        }
        std::atomic<uint64_t> total{0};
        const int rc = parallel_parts(bwt.R, n_parts, [&](unsigned part, uint64_t j0, uint64_t j1) {
            uint64_t cum[4], pos = start[part].pos, n_checked = 0;
            memcpy(cum, start[part].cum, sizeof cum);
            for (uint64_t j = j0; j < j1; ++j) {
                const int hc = code_of_head[bwt.heads[j]];
                // the first and last position of every run and every stride-th position inside
                for (uint64_t t = 0; t < bwt.lens[j]; t = (t + stride < bwt.lens[j] || t == bwt.lens[j] - 1) ? t + stride : bwt.lens[j] - 1) {
                    const uint64_t p = pos + t;
                    for (uint32_t c = 0; c < 4; ++c) {
                        if (!d.count[c]) continue;
                        const uint64_t want = d.Fcode[c] + cum[c] + ((int) c == hc ? t : 0);
                        if (dir_rank(d, c, p) != want) return 1;
                        if (dir_rank(d, c, p + 1) != want + ((int) c == hc ? 1 : 0)) return 2;
                        ++n_checked;
                    }
                }
                if (hc >= 0) cum[hc] += bwt.lens[j];
                pos += bwt.lens[j];
            }
            total += n_checked;
            return 0;
        });
        if (rc) return rc;
        const uint64_t n_checked = total;
        if (checked) *checked = n_checked;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Same for the phi slots: phi(i) decoded from the 32-byte slots (phi_slot.cuh, the code locate_kernel
// runs) against ToeholdSA::phi restated directly over the .tsa arrays (strict circular predecessor,
// include/toehold_sa.hpp:56-72), for every stride-th text position and both neighbours of every sample.
extern "C" int rbg_selftest_phi(const char* prefix, uint32_t shift, uint64_t stride, uint64_t* checked,
                                uint64_t* n_slots, uint64_t* n_overflow) {
    try {
        ToeholdArrays t = read_tsa(std::string(prefix) + ".tsa");
        PhiDir p = build_phi_dir(t, shift & 0xFFu, (shift >> 8) ? (uint64_t) (shift >> 8) : ~0ull);     // bits 8.. = slot budget in bytes (0 = none)
        if (n_slots) *n_slots = p.n_slots;
        if (n_overflow) *n_overflow = p.n_overflow;
        auto direct = [&](uint64_t i) {
            uint64_t rk = std::lower_bound(t.pred.begin(), t.pred.end(), i) - t.pred.begin();      // #samples < i
            const uint64_t jr = rk == 0 ? t.r - 1 : rk - 1;
            const uint64_t j = t.pred[jr];
            const uint64_t delta = j < i ? i - j : i + 1;
            const uint64_t run = t.pred_to_run[jr];
            return ((run ? t.samples_last[run - 1] : 0) + delta) % t.n;
        };
        if (stride == 0) stride = 1;
        std::atomic<uint64_t> total{0};
        const uint64_t n_strided = (t.n + stride - 1) / stride;
        int rc = parallel_parts(n_strided, selftest_threads(), [&](unsigned, uint64_t a, uint64_t b) {
            for (uint64_t x = a; x < b; ++x)
                if (phi_dir_eval(p, t.n, x * stride) != direct(x * stride)) return 1;
            total += b - a;
            return 0;
        });
        if (rc) return rc;
        rc = parallel_parts(t.r, selftest_threads(), [&](unsigned, uint64_t a, uint64_t b) {
            uint64_t n_checked = 0;
            for (uint64_t k = a; k < b; ++k)
                for (uint64_t i : {t.pred[k], t.pred[k] + 1, t.pred[k] ? t.pred[k] - 1 : 0}) {
                    if (i >= t.n) continue;
                    if (phi_dir_eval(p, t.n, i) != direct(i)) return 2;
                    ++n_checked;
                }
            total += n_checked;
            return 0;
        });
        if (rc) return rc;
        if (checked) *checked = total;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// The toehold directory (ToeholdDir, layout.hpp): for EVERY run j of the BWT, the row LF(end of run j) looked up
// through the bucket table and the reduced keys -- the host twin of toehold_at_row (device_index.cuh) -- must select
// samples_last[j] (include/rowbowt.hpp:562-566: after a non-trivial step the toehold is that run's sample), and rows
// that are no run-end image must not alias one.  shift = 0: the layout's own choice; else forces the bucket width
// (every key width: 1, 2, 4 bytes).
extern "C" int rbg_selftest_toehold(const char* prefix, uint32_t shift, uint64_t* checked, uint64_t* dir_bytes) {
    try {
        RunsBwt bwt = read_rbwt(std::string(prefix) + ".rbwt");
        ToeholdArrays t = read_tsa(std::string(prefix) + ".tsa");
        LeafDir d = build_leaf_dir(bwt, 0);
        char buf[16];
        if (shift) { snprintf(buf, sizeof buf, "%u", shift); setenv("RBG_TOEHOLD_SHIFT", buf, 1); }
        ToeholdDir td = build_toehold_dir(bwt, d.F, t);
        if (shift) unsetenv("RBG_TOEHOLD_SHIFT");
        if (shift && td.shift != shift) return 3;
        if (dir_bytes) *dir_bytes = td.bytes();
        // per-symbol counts in front of every part's first run
        const unsigned n_parts = selftest_threads();
        std::vector<std::vector<uint64_t>> seen0(n_parts + 1, std::vector<uint64_t>(256, 0));
        {
            std::vector<uint64_t> cur(256, 0);
            unsigned p = 0;
            for (uint64_t j = 0; j <= bwt.R; ++j) {
                while (p <= n_parts && j == bwt.R * p / n_parts) seen0[p++] = cur;
                if (j == bwt.R) break;
                cur[bwt.heads[j]] += bwt.lens[j];
            }
        }
        std::atomic<uint64_t> total{0};
        const int rc = parallel_parts(bwt.R, n_parts, [&](unsigned part, uint64_t j0, uint64_t j1) {
            std::vector<uint64_t> seen = seen0[part];
            for (uint64_t j = j0; j < j1; ++j) {
                const uint8_t c = bwt.heads[j];
                seen[c] += bwt.lens[j];
                const uint64_t row = d.F[c] + seen[c] - 1;
                const uint64_t k = toehold_dir_rank(td, row);
                if (k >= td.n_keys || td.sample.get(k) != t.samples_last[j]) return 1;
                if (toehold_dir_rank(td, row + 1) != k + 1) return 2;          // #keys < row + 1: this key and no other
            }
            total += j1 - j0;
            return 0;
        });
        if (rc) return rc;
        if (td.toehold0 != (t.samples_last[t.r - 1] + 1) % t.n) return 4;
        if (checked) *checked = total;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// rbg_pack_bytes without an index handle (CPU-only tests of the host packer): same call with an explicit
// byte -> code table (0..3, 4 = terminator, -1 = no symbol).
extern "C" int rbg_selftest_pack(const int8_t* code_of, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                                 uint64_t x0, uint64_t x1, uint64_t* packed, uint8_t* flags, uint64_t* n_exotic) {
    const uint64_t ex = pack_bytes_host(code_of, (const uint8_t*) bases, offsets, n_reads, x0, x1, packed, flags);
    if (n_exotic) *n_exotic += ex;
    return 0;
}

// Writer check without a GPU: decode <prefix>.rbwt/.tsa/.mab with the readers (formats.cpp) and serialize the
// flat arrays again with sdsl_writer.hpp into <out_prefix>.*; the caller compares the files byte for byte.
// parts: bit 0 .rbwt, bit 1 .tsa, bit 2 .mab; bit 3: <prefix>.rbwt is a wt_fbb (`rb_build --fbb`) -- decoded and
// written as an rle_string .rbwt, which must equal what the reference's plain rb_build writes for the same BWT.
extern "C" int rbg_selftest_rewrite(const char* prefix, const char* out_prefix, uint32_t parts) {
    try {
        const std::string in(prefix), out(out_prefix);
        if (parts & 8) write_rbwt(read_rbwt_fbb(in + ".rbwt"), out + ".rbwt");
        else if (parts & 1) write_rbwt(read_rbwt(in + ".rbwt"), out + ".rbwt");
        if (parts & 2) write_tsa(read_tsa(in + ".tsa"), out + ".tsa");
        if (parts & 4) write_mab(read_mab(in + ".mab"), out + ".mab");
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "rbg_selftest_rewrite: %s\n", e.what());
        return -1;
    }
}
