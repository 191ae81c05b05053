// Host-side consistency check of the load-time re-layout (diagnostic; not on the query path):
// rebuilds the rank directory from <prefix>.rbwt and compares rank_c(i) decoded from the
// 64-byte leaves (the same leaf.cuh code the kernels run) with a direct count over the runs.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rowbowt_gpu.h"
#include "formats.hpp"
#include "layout.hpp"
#include "leaf.cuh"

using namespace rbg;

namespace {
uint64_t dir_rank(const RankDir& d, int c, uint64_t pos, bool* is_c) {
    const uint32_t e = d.table[(uint64_t) c * d.n_buckets + (pos >> d.s)];
    const uint32_t k = e & 15, g = d.s - k;
    const uint64_t leaf = (e >> 4) + ((pos >> g) & ((1u << k) - 1));
    uint32_t w[16];
    memcpy(w, d.lines.data() + leaf * 16, 64);
    bool in;
    uint64_t r = leaf_base_count(w) + leaf_count(w, (uint32_t) (pos & ((1u << g) - 1)), in) - d.Fcode[c];
    *is_c = in;
    return r;
}
}  // namespace

extern "C" int rbg_selftest_layout(const char* prefix, uint32_t bucket_bits, uint64_t stride, uint64_t* checked,
                                   uint64_t* n_lines) {
    try {
        RunsBwt bwt = read_rbwt(std::string(prefix) + ".rbwt");
        RankDir d = build_rank_dir(bwt, bucket_bits);
        if (n_lines) *n_lines = d.n_lines();
        static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
        uint64_t cum[4] = {0, 0, 0, 0}, pos = 0, n_checked = 0;
        if (stride == 0) stride = 1;
        for (uint64_t j = 0; j < bwt.R; ++j) {
            int hc = -1;
            for (int c = 0; c < 4; ++c) if (bwt.heads[j] == sym[c]) hc = c;
            // check the first and last position of every run and every stride-th position inside
            for (uint64_t t = 0; t < bwt.lens[j]; t = (t + stride < bwt.lens[j] || t == bwt.lens[j] - 1) ? t + stride : bwt.lens[j] - 1) {
                const uint64_t p = pos + t;
                for (int c = 0; c < 4; ++c) {
                    if (!d.count[c]) continue;
                    bool is_c;
                    const uint64_t got = dir_rank(d, c, p, &is_c);
                    const uint64_t want = cum[c] + (c == hc ? t : 0);
                    if (got != want || is_c != (c == hc)) return 1;
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

// Same check for layout v2 (mixed leaves): rank_c at p and p+1 decoded from the 64-byte lines
// (the leaf.cuh code the kernels run, split leaves included) against a direct count.
namespace {
void mix_locate(const MixDir& d, uint64_t pos, uint32_t (&w)[16], uint32_t& q, uint32_t& size) {
    memcpy(w, d.lines.data() + (pos >> d.g) * 16, 64);
    size = 1u << d.g;
    if (mix_is_split(w)) {
        const uint32_t k = w[7], cg = d.g - k;
        const uint64_t child = (uint64_t) w[6] + ((pos >> cg) & ((1u << k) - 1u));
        memcpy(w, d.lines.data() + child * 16, 64);
        size = 1u << cg;
    }
    q = (uint32_t) (pos & (size - 1));
}
}  // namespace

extern "C" int rbg_selftest_mix(const char* prefix, uint32_t leaf_bits, uint64_t stride, uint64_t* checked,
                                uint64_t* n_lines, uint64_t* n_split) {
    try {
        RunsBwt bwt = read_rbwt(std::string(prefix) + ".rbwt");
        MixDir d = build_mix_dir(bwt, leaf_bits);
        if (n_lines) *n_lines = d.n_lines();
        if (n_split) *n_split = d.n_split;
        static const uint8_t sym[4] = {'A', 'C', 'G', 'T'};
        uint64_t cum[4] = {0, 0, 0, 0}, pos = 0, n_checked = 0;
        if (stride == 0) stride = 1;
        for (uint64_t j = 0; j < bwt.R; ++j) {
            int hc = -1;
            for (int c = 0; c < 4; ++c) if (bwt.heads[j] == sym[c]) hc = c;
            for (uint64_t t = 0; t < bwt.lens[j]; t = (t + stride < bwt.lens[j] || t == bwt.lens[j] - 1) ? t + stride : bwt.lens[j] - 1) {
                const uint64_t p = pos + t;
                uint32_t w[16], q, size;
                mix_locate(d, p, w, q, size);
                for (uint32_t c = 0; c < 4; ++c) {
                    if (!d.count[c]) continue;
                    uint32_t ra, rb, rc;
                    mix_count<true, true>(w, c, size, q, q + 1, q ? q - 1 : 0, ra, rb, rc);
                    const uint64_t want = d.Fcode[c] + cum[c] + ((int) c == hc ? t : 0);
                    if (mix_base_count(w, c) + ra != want) return 1;
                    if ((rb - ra == 1) != ((int) c == hc)) return 2;
                    if (q && t && mix_base_count(w, c) + rc != want - ((int) c == hc ? 1 : 0)) return 3;   // rank at p-1, same run
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
