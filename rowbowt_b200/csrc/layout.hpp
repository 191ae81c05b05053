// Load-time re-layout of the run-length BWT and its locate structures for the GPU.
//
// The reference answers rank_c(i) (include/rle_string.hpp:131-161) with ~10 dependent
// probes into sd_vectors and a wavelet tree.  Here the same function is answered from
// ONE 64-byte line whose address is computed from (c, i) and a small, L2-resident table.
//
// RankDir — one directory per symbol c in {A,C,G,T} (the analogue of runs_per_letter[c],
// but indexed by BWT position instead of by c-rank, so no second structure is needed):
//   BWT positions are cut into buckets of 2^s positions.  table[c][b] = (base<<4 | k):
//   bucket b is stored as 2^k consecutive "leaf" lines from line `base`, leaf t covering the 2^(s-k)
//   positions [b*2^s + t*2^(s-k), ...).  k is the smallest split for which every leaf
//   holds its c-runs in one line, so dense (short-run) regions are cut finer and long-run
//   regions stay coarse; the finest leaf (256 positions) always fits as a bitmap.
//
//   A leaf line (16 x u32, 64-byte aligned), for symbol c:
//     w[0]      low 32 bits of F[c] + #c in BWT[0, leaf_start)   (i.e. LF of the leaf start, were it a c)
//     w[1]      bits 0-7: bits 32-39 of that value; bits 8-11: mode
//     w[2..15]  payload
//        mode RUNS : 14 entries (len << 16 | start): the c-runs intersecting the leaf, clipped to
//                    it, start relative to the leaf (< 2^15), len in 1..2^15, 0 = padding
//        mode BITS : (256-position leaves only) w[2..9] bit p = (BWT[leaf_start+p] == c)
//   F[c] + rank_c(i) = header + sum_e clamp(q - start_e, 0, len_e),  q = i - leaf_start
//   BWT[i] == c <=> some entry has 0 <= q - start_e < len_e
//   The terminator (byte 1) has no directory: its (few, normally one) positions are kept
//   sorted in `term_pos`.
//
// Toehold / phi structures are sorted arrays with a radix bucket table in front (PredTable).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "formats.hpp"

namespace rbg {

struct alphabet_error : std::runtime_error { using std::runtime_error::runtime_error; };

constexpr int kLineWords = 16;          // 64-byte lines
constexpr int kMinLeafBits = 8;         // finest leaf = 256 positions (BITS always fits)
constexpr int kMaxLeafBits = 15;        // start/len fit 16 bits
constexpr int kRunEntries = 14;
constexpr int kMaxTerm = 8;             // terminator positions carried in kernel params
enum LeafMode : uint32_t { kRuns = 0, kBits = 1 };

struct RankDir {
    uint64_t n = 0;
    uint32_t s = 0;                      // log2 positions per bucket
    uint64_t n_buckets = 0;
    std::vector<uint32_t> table;         // [4][n_buckets]  base<<4 | k   (base = global line index)
    std::vector<uint32_t> lines;         // [n_lines * 16], all four symbols back to back
    uint64_t line_base[4] = {0, 0, 0, 0};
    uint64_t n_lines() const { return lines.size() / kLineWords; }
    uint64_t F[256] = {0};               // RowBowt::build_f (include/rowbowt.hpp:770-778), by byte value
    uint64_t Fcode[4] = {0, 0, 0, 0};    // F of A,C,G,T
    uint64_t count[4] = {0, 0, 0, 0};    // occurrences of A,C,G,T
    int8_t code_of[256];                 // byte -> 0..3, 4 = terminator (byte 1) present, -1 = not in the BWT
    uint32_t n_term = 0;
    uint64_t term_pos[kMaxTerm] = {0};
};

// Sorted u64 keys with a radix table: table[b] = #keys < (b << shift), b in [0, (universe>>shift)+1].
struct PredTable {
    uint32_t shift = 0;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> table;
};

// Toehold resolution: rows of the F column that are LF images of BWT run ends, with the
// SA sample of that run end.  After a non-trivial LF_w_loc step (include/rowbowt.hpp:562-566)
// the new hi IS such a row, and its toehold is samples_last[run] -- one lookup instead of
// rank + select + run_of_position.
struct ToeholdDir {
    PredTable rows;                      // keys = LF(end of run j), ascending
    std::vector<uint64_t> sample;        // sample[i] = samples_last of the run whose end maps to rows.keys[i]
    uint64_t toehold0 = 0;               // ToeholdSA::get_last_run_sample (include/toehold_sa.hpp:97-99)
};

// phi (include/toehold_sa.hpp:56-72): predecessor over `pred`, value fused at load:
// prev[jr] = samples_last[pred_to_run[jr] - 1].
struct PhiDir {
    PredTable pred;
    std::vector<uint64_t> prev;
};

// Layout v2: mixed leaves (leaf.cuh).  Direct-mapped: BWT position p lives in line p >> g; no
// table in front.  Leaves with more than 22 runs point to 2^k children in the overflow area
// [n_direct, n_lines).  One 64-byte line per rank, all symbols, 2 bytes per run.
struct MixDir {
    uint64_t n = 0;
    uint32_t g = 0;                      // log2 positions per direct leaf, 4..12
    uint64_t n_direct = 0;               // ceil(n / 2^g)
    uint64_t n_split = 0;                // direct leaves that are split
    std::vector<uint32_t> lines;         // [(n_direct + overflow) * 16]
    uint64_t n_lines() const { return lines.size() / kLineWords; }
    uint64_t F[256] = {0};               // RowBowt::build_f (include/rowbowt.hpp:770-778), by byte value
    uint64_t Fcode[4] = {0, 0, 0, 0};
    uint64_t count[4] = {0, 0, 0, 0};
    int8_t code_of[256];                 // byte -> 0..3, 4 = terminator (byte 1) present, -1 = not in the BWT
    uint32_t n_term = 0;
    uint64_t term_pos[kMaxTerm] = {0};
};
constexpr int kMixMinBits = 4;           // a 16-position child always fits (<= 16 runs)
constexpr int kMixMaxBits = 12;          // start field is 13 bits, the padding value 2^g must fit
MixDir build_mix_dir(const RunsBwt& bwt, uint32_t leaf_bits = 0);

// Validates the alphabet and builds the rank directory.  bucket_bits = 0 -> choose automatically.
RankDir build_rank_dir(const RunsBwt& bwt, uint32_t bucket_bits = 0);
ToeholdDir build_toehold_dir(const RunsBwt& bwt, const uint64_t (&F)[256], const ToeholdArrays& tsa);
PhiDir build_phi_dir(const ToeholdArrays& tsa);
PredTable build_pred_table(std::vector<uint64_t>&& keys, uint64_t universe, double keys_per_bucket);

}  // namespace rbg
