"""ORACLE (test infrastructure, not product code) — independent numpy readers for
the reference's serialized index files.

Each reader restates the byte layout written by the reference's `serialize`
methods (all integers little-endian; sdsl is xxsds v3, header-only, vendored
under /root/reference/sdsl-lite) and returns flat numpy arrays that mean what
SURVEY.md Appendix B.8 says they mean.  Nothing here is imported by the product
(`rowbowt_b200/`), which has its own C++ reader; tests compare the two and both
against the compiled reference (`oracle/_ref/ref_probe`).

Layouts followed (reference file:line):
  int_vector<w>      sdsl/int_vector.hpp:813-842,1815-1838   u64 (width<<56 | size_in_bits), ceil(bits/64) u64 words
  select_support_mcl sdsl/select_support_mcl.hpp:427-498     skipped (u64 arg_cnt; superblock; mini_or_long; per-superblock vectors)
  rank_support_v     sdsl/rank_support_v.hpp:121-135         one int_vector<64>, skipped
  sd_vector          sdsl/sd_vector.hpp:194-232,374-397      u64 size, u8 wl, low, high, select_1, select_0
  sparse_sd_vector   include/sparse_sd_vector.hpp:182-200    u64 u, then sd_vector iff u>0
  wt_huff (wt_pc)    sdsl/wt_pc.hpp:610-636; tree sdsl/wt_helper.hpp:117-134,313-340
  rle_string         include/rle_string.hpp:248-275
  ToeholdSA          include/toehold_sa.hpp:74-91
  rle_window_arr     pfbwt-f/include/rle_window_array.hpp:174-198
  DocList            include/doclist.hpp:57-73
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np


class _Reader:
    def __init__(self, path: str):
        self.buf = np.fromfile(path, dtype=np.uint8)
        self.pos = 0
        self.path = path

    def u64(self) -> int:
        v = struct.unpack_from("<Q", self.buf, self.pos)[0]
        self.pos += 8
        return v

    def u8(self) -> int:
        v = int(self.buf[self.pos])
        self.pos += 1
        return v

    def i32(self) -> int:
        v = struct.unpack_from("<i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def raw(self, nbytes: int) -> np.ndarray:
        a = self.buf[self.pos:self.pos + nbytes]
        if len(a) != nbytes:
            raise ValueError("truncated file %s" % self.path)
        self.pos += nbytes
        return a

    def words(self, n: int) -> np.ndarray:
        return self.raw(8 * n).view("<u8")

    def done(self) -> bool:
        return self.pos == len(self.buf)


def _int_vector(r: _Reader):
    """-> (width, size_in_bits, u64 words)"""
    h = r.u64()
    width, bits = h >> 56, h & ((1 << 56) - 1)
    return width, bits, r.words((bits + 63) // 64)


def _unpack(words: np.ndarray, width: int, count: int) -> np.ndarray:
    """count elements of `width` bits packed LSB-first at bit i*width."""
    if count == 0:
        return np.zeros(0, dtype=np.uint64)
    if width == 64:
        return words[:count].astype(np.uint64)
    w = np.concatenate([words, np.zeros(1, dtype=np.uint64)]).astype(np.uint64)
    bitpos = np.arange(count, dtype=np.uint64) * np.uint64(width)
    wi = (bitpos >> np.uint64(6)).astype(np.int64)
    sh = bitpos & np.uint64(63)
    lo = w[wi] >> sh
    # bits spilling into the next word (shift by 64 is undefined -> guard)
    spill = (sh + np.uint64(width)) > np.uint64(64)
    hi = np.where(spill, w[wi + 1] << ((np.uint64(64) - sh) & np.uint64(63)), np.uint64(0))
    return (lo | hi) & np.uint64((1 << width) - 1)


def _bits_set(words: np.ndarray, nbits: int) -> np.ndarray:
    """positions of the 1 bits of a bit_vector (LSB-first within each u64)."""
    b = np.unpackbits(words.view(np.uint8), bitorder="little")[:nbits]
    return np.nonzero(b)[0].astype(np.uint64)


def _skip_select_mcl(r: _Reader) -> None:
    arg_cnt = r.u64()
    if arg_cnt == 0:
        return
    _int_vector(r)                       # superblock
    _, mbits, mwords = _int_vector(r)    # mini_or_long
    sb = (arg_cnt + 4095) >> 12
    for _ in range(sb):
        _int_vector(r)                   # miniblock[i] or longsuperblock[i]


def _sd_vector(r: _Reader):
    """-> (size, sorted positions of the ones as uint64)"""
    size = r.u64()
    wl = r.u8()
    lw, lbits, lwords = _int_vector(r)
    _, hbits, hwords = _int_vector(r)
    _skip_select_mcl(r)
    _skip_select_mcl(r)
    m = lbits // lw if lw else 0
    if m == 0:
        return size, np.zeros(0, dtype=np.uint64)
    assert lw == wl, (lw, wl)
    low = _unpack(lwords, wl, m)
    hp = _bits_set(hwords, hbits)
    assert len(hp) == m, (len(hp), m)
    high = hp - np.arange(m, dtype=np.uint64)   # zeros before the i-th one
    return size, (high << np.uint64(wl)) | low


def _sparse_sd_vector(r: _Reader):
    u = r.u64()
    if u == 0:
        return 0, np.zeros(0, dtype=np.uint64)
    size, ones = _sd_vector(r)
    assert size == u
    return u, ones


def _wt_huff(r: _Reader) -> np.ndarray:
    """Decode every symbol of a Huffman-shaped wavelet tree (uint8[size])."""
    size = r.u64()
    sigma = r.u64()
    _, bvbits, bvwords = _int_vector(r)
    _int_vector(r)                       # rank_support_v basic blocks
    _skip_select_mcl(r)
    _skip_select_mcl(r)
    n_nodes = r.u64()
    nodes = []
    for _ in range(n_nodes):
        bv_pos, bv_pos_rank = r.u64(), r.u64()
        parent, c0, c1 = struct.unpack_from("<HHH", r.buf, r.pos)
        r.pos += 6
        nodes.append((bv_pos, bv_pos_rank, parent, c0, c1))
    r.raw(256 * 2)                       # c_to_leaf
    r.raw(256 * 8)                       # path
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    bv = np.unpackbits(bvwords.view(np.uint8), bitorder="little")[:bvbits]
    UNDEF = 0xFFFF

    def decode(v: int, length: int) -> np.ndarray:
        bv_pos, sym, _, c0, c1 = nodes[v]
        if c0 == UNDEF:                  # leaf: bv_pos_rank holds the symbol
            return np.full(length, sym & 0xFF, dtype=np.uint8)
        b = bv[bv_pos:bv_pos + length].astype(bool)
        out = np.empty(length, dtype=np.uint8)
        n1 = int(b.sum())
        out[~b] = decode(c0, length - n1)
        out[b] = decode(c1, n1)
        return out

    if n_nodes == 1:                     # sigma == 1: a lone leaf
        return np.full(size, nodes[0][1] & 0xFF, dtype=np.uint8)
    return decode(0, size)


@dataclass
class Rlbwt:
    n: int
    R: int
    B: int
    heads: np.ndarray    # uint8[R]  (terminator stored as 1, rle_string.hpp:57-62)
    lens: np.ndarray     # uint64[R]


def read_rbwt(path: str) -> Rlbwt:
    r = _Reader(path)
    n, R, B = r.u64(), r.u64(), r.u64()
    if n == 0:
        return Rlbwt(0, 0, B, np.zeros(0, np.uint8), np.zeros(0, np.uint64))
    runs_u, runs_ones = _sparse_sd_vector(r)
    per_letter = [_sparse_sd_vector(r) for _ in range(256)]
    heads = _wt_huff(r)
    assert r.done(), "trailing bytes in %s" % path
    assert len(heads) == R
    lens = np.zeros(R, dtype=np.uint64)
    tot = 0
    for c in range(256):
        u, ones = per_letter[c]
        idx = np.nonzero(heads == c)[0]
        assert len(idx) == len(ones), (c, len(idx), len(ones))
        if len(ones) == 0:
            continue
        ends = ones.astype(np.int64)
        lc = np.diff(np.concatenate([[-1], ends]))
        assert u == ends[-1] + 1 and (lc > 0).all()
        lens[idx] = lc.astype(np.uint64)
        tot += u
    assert tot == n and int(lens.sum()) == n
    # `runs` marks the last position of every B-th run, except the final run
    ends = np.cumsum(lens.astype(np.int64)) - 1
    expect = ends[B - 1::B]
    expect = expect[expect != n - 1] if (R % B == 0) else expect
    assert runs_u == n and np.array_equal(expect.astype(np.uint64), runs_ones), "runs bitvector mismatch"
    return Rlbwt(n, R, B, heads, lens)


@dataclass
class Toehold:
    r: int
    n: int
    pred: np.ndarray          # uint64[r] sorted text positions SA[run start]-1 (mod n)
    samples_last: np.ndarray  # uint64[r] SA[run end]-1 (mod n), BWT order
    pred_to_run: np.ndarray   # uint64[r]


def read_tsa(path: str) -> Toehold:
    rd = _Reader(path)
    r, n = rd.u64(), rd.u64()
    u, pred = _sparse_sd_vector(rd)
    w1, b1, d1 = _int_vector(rd)
    w2, b2, d2 = _int_vector(rd)
    assert rd.done(), "trailing bytes in %s" % path
    assert u == n and len(pred) == r
    samples_last = _unpack(d1, w1, b1 // w1)
    pred_to_run = _unpack(d2, w2, b2 // w2)
    assert len(samples_last) == r and len(pred_to_run) == r
    return Toehold(r, n, pred, samples_last, pred_to_run)


@dataclass
class MarkerWindows:
    size_starts: int
    size_ends: int
    size_idxs: int
    starts: np.ndarray   # uint64[W] sorted rows where a window starts
    ends: np.ndarray     # uint64[W] sorted rows where a window ends
    idxs: np.ndarray     # uint64[W] offsets into arr where each window's values start
    arr: np.ndarray      # uint64[] marker words
    wsize: int


def read_mab(path: str) -> MarkerWindows:
    r = _Reader(path)
    s1, starts = _sd_vector(r)
    s2, ends = _sd_vector(r)
    s3, idxs = _sd_vector(r)
    arr_size = r.u64()
    arr = r.words(arr_size).astype(np.uint64)
    wsize = r.i32()
    assert r.done(), "trailing bytes in %s" % path
    return MarkerWindows(s1, s2, s3, starts, ends, idxs, arr, wsize)


def read_docs(path: str):
    """-> (names, starts) exactly as DocList::load tokenises: `ifs >> name >> pos`."""
    names, starts = [], []
    toks = open(path, "rb").read().split()
    for i in range(0, len(toks) - 1, 2):
        try:
            p = int(toks[i + 1])
        except ValueError:
            break
        names.append(toks[i].decode())
        starts.append(p)
    return names, np.array(starts, dtype=np.uint64)


# ------------------------------------------------------------------------------------------------
# wt_fbb (`rb_build --fbb`): include/fbb_string.hpp -> faster-minuter/include/wt_fbb.hpp, compiled with
# ADD_NAVIGATIONAL_BLOCK_HEADER, ALLOW_VARIABLE_BLOCK_SIZE, no SPARSE_SUPERBLOCK_MAPPING (:48-51), over
# sdsl::hyb_vector<16> (sdsl/hyb_vector.hpp).  Only what a sequential decode of the whole string needs.
#   wt_fbb::serialize :1849-1861  u64 size; vector<u64> count; vector<u64> hyperblock_rank;
#                                 vector<u32> superblock_rank; vector<u8> global_mapping; vector<superblock_header>
#   std::vector<X>                sdsl/io.hpp:145-152,358-377: u64 n, then the n elements (PODs raw)
#   superblock_header :124-146    u8 sigma-1, u8 block_size_log, hyb_vector, rank_support_hyb (0 bytes,
#                                 hyb_vector.hpp:771-776), vector<u8> var_block_headers,
#                                 vector<block_header_item> (14 packed bytes, :93-101), vector<u8> mapping
#   hyb_vector :31-41             u64 size, int_vector<8> trunk, int_vector<8> sblock_header, int_vector<64> hblock_header
#   block body :343-431, header :434-492; canonical codes :245-267
def _vec(r: _Reader, elem_bytes: int) -> np.ndarray:
    n = r.u64()
    return r.raw(n * elem_bytes)


def _hyb_bits(r: _Reader) -> np.ndarray:
    """sdsl::hyb_vector<16> -> uint8[size] of bits.  Blocks of 256 bits, one u16 header each inside a 40-byte
    superblock header (8 + 2*16): ones = h & 0x1ff, special = bit 9, encoded bytes = h >> 10
    (hyb_vector.hpp:256-265): 0 = at most two runs, 32 = plain, min(ones, zeros) = minority positions, else the
    end positions of all runs but the last two (:267-357; decode logic restated from access0 :372-510)."""
    size = r.u64()
    _, tb, tw = _int_vector(r)
    trunk = tw.view(np.uint8)[:tb // 8]
    _, sb, sw = _int_vector(r)
    sbh = sw.view(np.uint8)[:sb // 8]
    _int_vector(r)                                            # hblock headers (trunk / rank bases): not needed sequentially
    n_blocks = (size + 255) // 256
    out = np.zeros(n_blocks * 256, dtype=np.uint8)
    tp = 0
    for b in range(n_blocks):
        hp = (b // 16) * 40 + 8 + (b % 16) * 2
        h = int(sbh[hp]) | (int(sbh[hp + 1]) << 8)
        ones, special, enc = h & 0x1FF, (h >> 9) & 1, h >> 10
        zeros = 256 - ones
        blk = out[b * 256:(b + 1) * 256]
        if enc == 0:
            first = ones if special else zeros
            blk[:first] = special
            blk[first:] = 1 - special
        elif enc >= 32:
            blk[:] = np.unpackbits(trunk[tp:tp + 32], bitorder="little")
        elif enc == min(ones, zeros):
            blk[:] = 1 - special
            blk[trunk[tp:tp + enc].astype(np.int64)] = special
        else:
            bit, pos, cnt = special, 0, [0, 0]
            for e in trunk[tp:tp + enc].astype(np.int64):
                blk[pos:e + 1] = bit
                cnt[bit] += e + 1 - pos
                pos, bit = e + 1, 1 - bit
            first = (ones - cnt[1]) if bit else (zeros - cnt[0])   # the last two runs follow from the popcount
            blk[pos:pos + first] = bit
            blk[pos + first:] = 1 - bit
        tp += enc
    assert tp == len(trunk), "hyb_vector: trunk not consumed"
    return out[:size]


def read_fbb_text(path: str) -> np.ndarray:
    """uint8[n]: the string a wt_fbb .rbwt holds (the raw .bwt bytes, terminator = byte 0)."""
    r = _Reader(path)
    n = r.u64()
    _vec(r, 8); _vec(r, 8); _vec(r, 4); _vec(r, 1)           # count, hyperblock_rank, superblock_rank, global_mapping
    n_sb = r.u64()
    text = np.zeros(n, dtype=np.uint8)
    SB = 1 << 20
    for sb in range(n_sb):
        r.u8()                                               # superblock sigma - 1
        bs_log = r.u8()
        bv = _hyb_bits(r)
        var = _vec(r, 1)
        bh = _vec(r, 14)
        _vec(r, 1)                                           # mapping
        n_blk = len(bh) // 14
        for b in range(n_blk):
            _bv_rank, bv_off, var_off, sig1, height = struct.unpack_from("<IIIBB", bh, 14 * b)
            beg = sb * SB + (b << bs_log)
            bsz = min(1 << bs_log, n - beg)
            if height == 0:
                text[beg:beg + bsz] = var[var_off]
                continue
            sigma = sig1 + 1
            leaves_at = [0] + [int(var[var_off + 3 * (d - 1)]) for d in range(1, height)]
            leaves_at.append(sigma - sum(leaves_at))
            lp = var_off + 3 * (height - 1)
            syms = [int(var[lp + 4 * k]) for k in range(sigma)]
            # level structure: existing nodes of depth d are the children of depth d-1's internal nodes,
            # leaves leftmost (canonical codes, shortest first)
            leaf_base = np.cumsum([0] + leaves_at)
            sizes = [[bsz]]                                   # per level: sizes of the internal nodes
            offs = []
            p = bv_off
            for d in range(height):
                offs.append([])
                nxt = []
                for sz in sizes[d]:
                    offs[d].append(p)
                    o = int(bv[p:p + sz].sum())
                    nxt += [sz - o, o]
                    p += sz
                sizes.append(nxt[leaves_at[d + 1]:] if d + 1 <= height else [])
            cur = [[0] * len(o) for o in offs]
            for i in range(bsz):
                d, t = 0, 0
                while True:
                    bit = int(bv[offs[d][t] + cur[d][t]])
                    cur[d][t] += 1
                    idx = 2 * t + bit
                    if idx < leaves_at[d + 1]:
                        text[beg + i] = syms[leaf_base[d + 1] + idx]
                        break
                    t = idx - leaves_at[d + 1]
                    d += 1
    if not r.done():
        raise ValueError("wt_fbb: trailing bytes in %s" % path)
    return text
