// Host-side consistency check of the load-time re-layout (diagnostic; not on the query path):
// rebuilds the rank directory from <prefix>.rbwt and compares rank_c at p and p+1 (hence
// BWT[p]==c) decoded from the 64-byte lines -- the same leaf.cuh code the kernels run, split
// windows and the terminator correction included -- with a direct count over the runs.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rowbowt_gpu.h"
#include "formats.hpp"
#include "layout.hpp"
#include "leaf.cuh"

using namespace rbg;

namespace {
// F[c] + rank_c(pos) through the directory, pos in [0, n]; mirrors lf_step (device_index.cuh).
uint64_t dir_rank(const LeafDir& d, uint32_t c, uint64_t pos) {
    // rank at pos is taken from the line of position pos-1 with q = offset+1 (as for hi), or of pos with q = offset
    const bool use_prev = pos == d.n;
    const uint64_t at = use_prev ? pos - 1 : pos;
    const uint32_t wmask = (1u << d.g) - 1u;
    uint32_t w[16];
    memcpy(w, d.lines.data() + (at >> d.g) * 16, 64);
    if ((w[15] & kModeMask) == kModeSplit) {
        const uint64_t child = (uint64_t) w[0] + leaf_child_of(w, (uint32_t) at & wmask);
        memcpy(w, d.lines.data() + child * 16, 64);
    }
    const uint32_t q = ((uint32_t) at & wmask) + (use_prev ? 1u : 0u);
    uint64_t r = leaf_base_count(w, c) + leaf_rank(w, leaf_cpat(c), q);
    if ((w[15] & kModeMask) == kModeTerm && c == 0) {
        const uint64_t ws = at - ((uint32_t) at & wmask), from = ws + leaf_first_start(w), to = ws + q;
        for (uint32_t t = 0; t < d.n_term; ++t) r -= (d.term_pos[t] >= from && d.term_pos[t] < to) ? 1 : 0;
    }
    return r;
}
}  // namespace

extern "C" int rbg_selftest_layout(const char* prefix, uint32_t leaf_bits, uint64_t stride, uint64_t* checked,
                                   uint64_t* n_lines, uint64_t* n_split) {
    try {
        RunsBwt bwt = read_rbwt(std::string(prefix) + ".rbwt");
        LeafDir d = build_leaf_dir(bwt, leaf_bits);
        if (n_lines) *n_lines = d.n_lines();
        if (n_split) *n_split = d.n_split;
        static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
        uint64_t cum[4] = {0, 0, 0, 0}, pos = 0, n_checked = 0;
        if (stride == 0) stride = 1;
        for (uint64_t j = 0; j < bwt.R; ++j) {
            int hc = -1;
            for (int c = 0; c < 4; ++c) if (bwt.heads[j] == sym[c]) hc = c;
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
        if (checked) *checked = n_checked;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
