// Writers for the reference's serialized index formats (.rbwt / .tsa / .mab), byte for byte.
//
// rb_build stays drop-in: the files this produces must be IDENTICAL to
// what the reference's rb_build writes, so that either tool's output loads in either rb_align.
// The reference serializes sdsl (xxsds v3) objects whose auxiliary tables (select_support_mcl,
// rank_support_v, the Huffman-shaped wavelet tree) are part of the file; this header restates
// their construction from the published sources, operating on flat arrays -- no sdsl dependency.
//
//   int_vector / bit_vector  sdsl/int_vector.hpp:832-842,1815-1823   u64 (width<<56 | bits) + words
//   select_support_mcl       sdsl/select_support_mcl.hpp:105-113 (slow below 100000 bits, else fast),
//                            :190-244 init_slow, :247-345 init_fast, :404-418 initData, :427-466 serialize
//   rank_support_v           sdsl/rank_support_v.hpp:56-95
//   sd_vector                sdsl/sd_vector.hpp:194-232 (construction), :374-387 (serialize)
//   sparse_sd_vector         include/sparse_sd_vector.hpp:23-32,182-189
//   wt_huff                  sdsl/wt_pc.hpp:68-104,165-208 (construction), :610-623 (serialize);
//                            sdsl/wt_huff.hpp:73-100 (tree shape); sdsl/wt_helper.hpp:192-262 (BFS layout,
//                            paths), :264-271 (node ranks), :313-327 + :117-129 (tree, 22-byte nodes)
//   rle_string               include/rle_string.hpp:44-97 (what the bit vectors mean), :248-260
//   ToeholdSA                include/toehold_sa.hpp:74-83,105-131
//   rle_window_arr           pfbwt-f/include/rle_window_array.hpp:15-50,174-187
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <queue>
#include <string>
#include <vector>

#include "formats.hpp"

namespace rbg {
namespace sdslw {

inline uint32_t hi(uint64_t x) { return x ? 63 - (uint32_t) __builtin_clzll(x) : 0; }   // bits::hi (hi(0) = 0)

// i-th (1-based) set bit of w
inline uint32_t sel64(uint64_t w, uint32_t i) {
    for (uint32_t k = 1; k < i; ++k) w &= w - 1;
    return (uint32_t) __builtin_ctzll(w);
}

class Out {
  public:
    explicit Out(const std::string& path) : fp_(fopen(path.c_str(), "wb")) {
        if (!fp_) throw io_error("cannot write " + path);
        setvbuf(fp_, nullptr, _IOFBF, 1 << 22);
    }
    ~Out() { if (fp_) fclose(fp_); }
    void raw(const void* p, size_t n) {
        if (n && fwrite(p, 1, n, fp_) != n) throw io_error("short write");
    }
    void u64(uint64_t v) { raw(&v, 8); }
    void u8(uint8_t v) { raw(&v, 1); }
    void close() {
        if (fp_ && fclose(fp_) != 0) { fp_ = nullptr; throw io_error("close failed"); }
        fp_ = nullptr;
    }

  private:
    FILE* fp_;
};

// sdsl::int_vector<0> / bit_vector (width 1): elements packed LSB-first at i*width.
struct IntVec {
    uint8_t width = 64;
    uint64_t size = 0;                       // elements
    std::vector<uint64_t> w;                 // ceil(size*width/64) words, padding bits 0
    IntVec() {}
    IntVec(uint64_t n, uint8_t wd) : width(wd ? wd : 64), size(n), w((n * (uint64_t) (wd ? wd : 64) + 63) / 64, 0) {}
    uint64_t bits() const { return size * width; }
    void set(uint64_t i, uint64_t v) {       // the vector is zero-initialised and every element written once
        if (width < 64) v &= (1ull << width) - 1;
        const uint64_t b = i * width;
        const uint32_t s = b & 63;
        w[b >> 6] |= v << s;
        if (s + width > 64) w[(b >> 6) + 1] |= v >> (64 - s);
    }
    bool bit(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    void write(Out& o) const {
        o.u64(((uint64_t) width << 56) | bits());
        o.raw(w.data(), ((bits() + 63) / 64) * 8);
    }
};

// select_support_mcl<t_b,1> over bit vector `v`, serialized.
inline void write_select_mcl(Out& o, const IntVec& v, int t_b) {
    const uint64_t nbits = v.size;
    const uint64_t nwords = (nbits + 63) >> 6;
    const uint64_t* data = v.w.data();
    uint64_t ones = 0;
    for (uint64_t i = 0; i < nwords; ++i) ones += (uint64_t) __builtin_popcountll(data[i]);
    const uint64_t arg_cnt = t_b ? ones : nbits - ones;
    o.u64(arg_cnt);
    if (!arg_cnt) return;
    auto found = [&](uint64_t i) { return (((data[i >> 6] >> (i & 63)) & 1) != 0) == (t_b != 0); };   // may look at padding bits, as sdsl does
    const uint64_t SB = 4096;
    const uint32_t logn = hi(((nbits + 63) >> 6) << 6) + 1;
    const uint64_t logn4 = (uint64_t) logn * logn * logn * logn;
    const uint64_t sb = (arg_cnt + SB - 1) / SB;
    IntVec superblock(sb, (uint8_t) logn);
    std::vector<IntVec> mini(sb), lng(sb + 1);
    std::vector<char> has_mini(sb + 1, 0), has_long(sb + 1, 0);
    bool any_long = false;
    std::vector<uint64_t> pos(SB, 0);
    if (nbits < 100000) {                                    // init_slow
        uint64_t cnt = 0, sb_cnt = 0;
        for (uint64_t i = 0; i < nbits; ++i) {
            if (!found(i)) continue;
            pos[cnt % SB] = i;
            ++cnt;
            if (cnt % SB == 0 || cnt == arg_cnt) {
                superblock.set(sb_cnt, pos[0]);
                const uint64_t last = (cnt - 1) % SB;
                const uint64_t diff = pos[last] - pos[0];
                if (diff > logn4) {
                    any_long = true;
                    lng[sb_cnt] = IntVec(SB, (uint8_t) (hi(pos[last]) + 1));
                    has_long[sb_cnt] = 1;
                    for (uint64_t j = 0; j <= last; ++j) lng[sb_cnt].set(j, pos[j]);
                } else {
                    mini[sb_cnt] = IntVec(64, (uint8_t) (hi(diff) + 1));
                    has_mini[sb_cnt] = 1;
                    for (uint64_t j = 0; j <= last; j += 64) mini[sb_cnt].set(j / 64, pos[j] - pos[0]);
                }
                ++sb_cnt;
            }
        }
    } else {                                                 // init_fast
        uint64_t last_k64 = 1, sb_cnt = 0, cnt_old = 0, cnt_new = 0, last_k64_sum = 1;
        for (uint64_t wi = 0; wi < nwords; ++wi) {
            const uint64_t word = t_b ? data[wi] : ~data[wi];
            cnt_new += (uint64_t) __builtin_popcountll(word);
            if (cnt_new >= last_k64_sum) {
                pos[last_k64 - 1] = wi * 64 + sel64(word, (uint32_t) (last_k64_sum - cnt_old));
                last_k64 += 64;
                last_k64_sum += 64;
                if (last_k64 == SB + 1) {
                    if (sb_cnt < sb) superblock.set(sb_cnt, pos[0]);
                    uint64_t last_pos = pos[last_k64 - 65];
                    for (uint64_t ii = pos[last_k64 - 65] + 1, j = last_k64 - 65; ii < nbits && j < SB; ++ii)
                        if (found(ii)) { last_pos = ii; ++j; }
                    const uint64_t diff = last_pos - pos[0];
                    if (diff > logn4) {
                        any_long = true;
                        IntVec L(SB, (uint8_t) (hi(last_pos) + 1));
                        for (uint64_t j = pos[0], k = 0; k < SB && j <= last_pos; ++j)
                            if (found(j)) L.set(k++, j);
                        lng[std::min(sb_cnt, sb)] = std::move(L);
                        has_long[std::min(sb_cnt, sb)] = 1;
                    } else if (sb_cnt < sb) {
                        mini[sb_cnt] = IntVec(64, (uint8_t) (hi(diff) + 1));
                        has_mini[sb_cnt] = 1;
                        for (uint64_t j = 0; j < SB; j += 64) mini[sb_cnt].set(j / 64, pos[j] - pos[0]);
                    }
                    ++sb_cnt;
                    last_k64 = 1;
                }
            }
            cnt_old = cnt_new;
        }
        if (last_k64 > 1) {                                  // the last, partial block is always stored long
            any_long = true;
            IntVec L(SB, (uint8_t) (hi(nbits - 1) + 1));
            for (uint64_t i = pos[0], k = 0; i < nbits; ++i)
                if (found(i)) L.set(k++, i);
            lng[std::min(sb_cnt, sb)] = std::move(L);
            has_long[std::min(sb_cnt, sb)] = 1;
        }
    }
    superblock.write(o);
    IntVec mini_or_long(any_long ? sb : 0, 1);
    if (any_long)
        for (uint64_t i = 0; i < sb; ++i)
            if (has_mini[i]) mini_or_long.set(i, 1);
    mini_or_long.write(o);
    for (uint64_t i = 0; i < sb; ++i) {
        if (any_long && !has_mini[i]) lng[i].write(o);
        else mini[i].write(o);
    }
}

// rank_support_v<1,1> over `v`, serialized (one int_vector<64>).
inline void write_rank_v(Out& o, const IntVec& v) {
    const uint64_t nbits = v.size;
    std::vector<uint64_t> bb;
    if (nbits == 0) {
        bb.assign(2, 0);
    } else {
        bb.assign((((nbits + 63) >> 9) + 1) << 1, 0);
        const uint64_t* data = v.w.data();
        uint64_t i, j = 0;
        uint64_t sum = (uint64_t) __builtin_popcountll(data[0]), second = 0;
        const uint64_t nwords = (nbits + 63) >> 6;
        for (i = 1; i < nwords; ++i) {
            if (!(i & 7)) {
                j += 2;
                bb[j - 1] = second;
                bb[j] = bb[j - 2] + sum;
                second = sum = 0;
            } else {
                second |= sum << (63 - 9 * (i & 7));
            }
            sum += (uint64_t) __builtin_popcountll(data[i]);
        }
        if (i & 7) {
            second |= sum << (63 - 9 * (i & 7));
            bb[j + 1] = second;
        } else {
            j += 2;
            bb[j - 1] = second;
            bb[j] = bb[j - 2] + sum;
            bb[j + 1] = 0;
        }
    }
    o.u64(((uint64_t) 64 << 56) | (bb.size() * 64));
    o.raw(bb.data(), bb.size() * 8);
}

// sd_vector over a bit vector of `size` bits whose ones are `ones` (sorted, distinct), serialized.
inline void write_sd_vector(Out& o, uint64_t size, const uint64_t* ones, uint64_t m) {
    uint8_t logm = (uint8_t) (hi(m) + 1), logn = (uint8_t) (hi(size) + 1);
    if (logm == logn) --logm;
    const uint8_t wl = (uint8_t) (logn - logm);
    IntVec low(m, wl), high(m + (1ull << logm), 1);
    for (uint64_t i = 0; i < m; ++i) {
        low.set(i, ones[i]);
        const uint64_t hp = (ones[i] >> wl) + i;
        high.w[hp >> 6] |= 1ull << (hp & 63);
    }
    o.u64(size);
    o.u8(wl);
    low.write(o);
    high.write(o);
    write_select_mcl(o, high, 1);
    write_select_mcl(o, high, 0);
}

// ri::sparse_sd_vector built from a vector<bool> of `u` bits
inline void write_sparse_sd(Out& o, uint64_t u, const uint64_t* ones, uint64_t m) {
    o.u64(u);
    if (u == 0) return;
    write_sd_vector(o, u, ones, m);
}

// sdsl::wt_huff<> over the byte sequence `s` (no zero bytes), serialized.
inline void write_wt_huff(Out& o, const uint8_t* s, uint64_t n) {
    if (n == 0) {                                            // default-constructed members
        o.u64(0); o.u64(0);
        IntVec bv(0, 1);
        bv.write(o);
        o.u64((uint64_t) 64 << 56);                         // rank_support_v of a null vector: empty int_vector<64>
        o.u64(0); o.u64(0);                                 // two empty select supports
        o.u64(0);                                           // tree: no nodes
        std::vector<uint8_t> z(256 * 2 + 256 * 8, 0);
        o.raw(z.data(), z.size());
        return;
    }
    uint64_t C[256] = {0};
    for (uint64_t i = 0; i < n; ++i) ++C[s[i]];
    uint64_t sigma = 0;
    for (int c = 0; c < 256; ++c) sigma += C[c] > 0;
    // Huffman shape: min-heap of (frequency, node index), ties by node index (std::greater on pairs)
    struct PcNode { uint64_t freq, sym, parent, child[2]; };
    const uint64_t UNDEF = ~0ull;
    std::vector<PcNode> tmp;
    typedef std::pair<uint64_t, uint64_t> P;
    std::priority_queue<P, std::vector<P>, std::greater<P>> pq;
    for (uint64_t c = 0; c < 256; ++c)
        if (C[c]) { pq.push(P(C[c], tmp.size())); tmp.push_back(PcNode{C[c], c, UNDEF, {UNDEF, UNDEF}}); }
    while (pq.size() > 1) {
        P v1 = pq.top(); pq.pop();
        P v2 = pq.top(); pq.pop();
        tmp[v1.second].parent = tmp.size();
        tmp[v2.second].parent = tmp.size();
        pq.push(P(v1.first + v2.first, tmp.size()));
        tmp.push_back(PcNode{v1.first + v2.first, 0, UNDEF, {v1.second, v2.second}});
    }
    // _byte_tree: BFS order, root first; bv_pos = start of the node's bits
    struct Node { uint64_t bv_pos, bv_pos_rank; uint16_t parent, child[2]; };
    const uint16_t U16 = 0xFFFF;
    auto conv = [&](const PcNode& p) { return Node{p.freq, p.sym, (uint16_t) p.parent, {(uint16_t) p.child[0], (uint16_t) p.child[1]}}; };
    std::vector<Node> nodes(tmp.size());
    nodes[0] = conv(tmp.back());
    uint64_t bv_size = 0;
    size_t node_cnt = 1;
    uint16_t last_parent = U16;
    std::deque<uint16_t> q;
    q.push_back(0);
    while (!q.empty()) {
        const uint16_t idx = q.front();
        q.pop_front();
        const uint64_t frq = nodes[idx].bv_pos;
        nodes[idx].bv_pos = bv_size;
        if (nodes[idx].child[0] != U16) bv_size += frq;
        if (idx > 0) {
            if (last_parent != nodes[idx].parent) nodes[nodes[idx].parent].child[0] = idx;
            else nodes[nodes[idx].parent].child[1] = idx;
            last_parent = nodes[idx].parent;
        }
        if (nodes[idx].child[0] != U16) {
            for (int k = 0; k < 2; ++k) {
                nodes[node_cnt] = conv(tmp[nodes[idx].child[k]]);
                nodes[node_cnt].parent = idx;
                q.push_back((uint16_t) node_cnt);
                nodes[idx].child[k] = (uint16_t) node_cnt++;
            }
        }
    }
    uint16_t c_to_leaf[256];
    uint64_t path[256];
    for (int c = 0; c < 256; ++c) c_to_leaf[c] = U16;
    for (size_t v = 0; v < nodes.size(); ++v)
        if (nodes[v].child[0] == U16) c_to_leaf[(uint8_t) nodes[v].bv_pos_rank] = (uint16_t) v;
    for (uint32_t c = 0, prev_c = 0; c < 256; ++c) {
        if (c_to_leaf[c] != U16) {
            uint16_t v = c_to_leaf[c];
            uint64_t pw = 0, pl = 0;
            while (v != 0) {
                pw <<= 1;
                if (nodes[nodes[v].parent].child[1] == v) pw |= 1;
                ++pl;
                v = nodes[v].parent;
            }
            path[c] = pw | (pl << 56);
            prev_c = c;
        } else {
            path[c] = prev_c;
        }
    }
    // the bit sequence: symbol by symbol along its path, each node's bits consecutive from its bv_pos
    IntVec bv(bv_size, 1);
    std::vector<uint64_t> cur(nodes.size());
    for (size_t v = 0; v < nodes.size(); ++v) cur[v] = nodes[v].bv_pos;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t p = path[s[i]];
        const uint32_t len = (uint32_t) (p >> 56);
        uint16_t v = 0;
        for (uint32_t l = 0; l < len; ++l, p >>= 1) {
            if (p & 1) bv.w[cur[v] >> 6] |= 1ull << (cur[v] & 63);
            ++cur[v];
            v = nodes[v].child[p & 1];
        }
    }
    // inner nodes: bv_pos_rank = ones before bv_pos
    {
        std::vector<uint64_t> pre(bv.w.size() + 1, 0);
        for (size_t i = 0; i < bv.w.size(); ++i) pre[i + 1] = pre[i] + (uint64_t) __builtin_popcountll(bv.w[i]);
        for (auto& nd : nodes)
            if (nd.child[0] != U16) {
                const uint64_t p = nd.bv_pos;
                uint64_t r = pre[p >> 6];
                if (p & 63) r += (uint64_t) __builtin_popcountll(bv.w[p >> 6] & ((1ull << (p & 63)) - 1));
                nd.bv_pos_rank = r;
            }
    }
    o.u64(n);
    o.u64(sigma);
    bv.write(o);
    write_rank_v(o, bv);
    write_select_mcl(o, bv, 1);
    write_select_mcl(o, bv, 0);
    o.u64(nodes.size());
    for (const Node& nd : nodes) {
        o.u64(nd.bv_pos);
        o.u64(nd.bv_pos_rank);
        o.raw(&nd.parent, 2);
        o.raw(nd.child, 4);
    }
    o.raw(c_to_leaf, sizeof c_to_leaf);
    o.raw(path, sizeof path);
}

}  // namespace sdslw

// rle_string::serialize of the run-length BWT (B = 2).
inline void write_rbwt(const RunsBwt& b, const std::string& path) {
    sdslw::Out o(path);
    const uint64_t B = 2;
    o.u64(b.n);
    o.u64(b.R);
    o.u64(B);
    if (b.n == 0) { o.close(); return; }
    // runs: 1 at the last position of every B-th run, never for the final run
    std::vector<uint64_t> ones;
    ones.reserve(b.R / B + 1);
    std::vector<std::vector<uint64_t>> per(256);
    uint64_t cnt[256] = {0};
    uint64_t pos = 0;
    for (uint64_t j = 0; j < b.R; ++j) {
        pos += b.lens[j];
        if (j % B == B - 1 && j + 1 < b.R) ones.push_back(pos - 1);
        const uint8_t c = b.heads[j];
        cnt[c] += b.lens[j];
        per[c].push_back(cnt[c] - 1);                   // 1 at the last position of every c-run
    }
    sdslw::write_sparse_sd(o, b.n, ones.data(), ones.size());
    for (int c = 0; c < 256; ++c) sdslw::write_sparse_sd(o, cnt[c], per[c].data(), per[c].size());
    sdslw::write_wt_huff(o, b.heads.data(), b.R);
    o.close();
}

inline uint8_t bitsize(uint64_t x) { return x ? (uint8_t) (64 - __builtin_clzll(x)) : 1; }   // include/utils.hpp:21-24

// ToeholdSA::serialize
inline void write_tsa(const ToeholdArrays& t, const std::string& path) {
    sdslw::Out o(path);
    o.u64(t.r);
    o.u64(t.n);
    sdslw::write_sparse_sd(o, t.n, t.pred.data(), t.pred.size());
    sdslw::IntVec sl(t.r, bitsize(t.n)), p2r(t.r, bitsize(t.r));
    for (uint64_t i = 0; i < t.samples_last.size() && i < t.r; ++i) sl.set(i, t.samples_last[i]);
    for (uint64_t i = 0; i < t.pred_to_run.size() && i < t.r; ++i) p2r.set(i, t.pred_to_run[i]);
    sl.write(o);
    p2r.write(o);
    o.close();
}

// rle_window_arr::serialize
inline void write_mab(const MarkerArrays& m, const std::string& path) {
    sdslw::Out o(path);
    sdslw::write_sd_vector(o, m.size_starts, m.starts.data(), m.starts.size());
    sdslw::write_sd_vector(o, m.size_ends, m.ends.data(), m.ends.size());
    sdslw::write_sd_vector(o, m.size_idxs, m.idxs.data(), m.idxs.size());
    o.u64(m.arr.size());
    o.raw(m.arr.data(), m.arr.size() * 8);
    o.raw(&m.wsize, 4);
    o.close();
}

}  // namespace rbg
