"""ORACLE (test infrastructure) — ctypes front end of oracle/liboracle.so, the
plain-C restatement of the reference rb_align query path, plus a text renderer
that restates rb_report's stdout grammar (src/rb_align.cpp:118-145).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import rbformats as F

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REFBIN = os.path.join(HERE, "_ref")
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)

POS_MASK = 0x00000FFFFFFFFFFF   # pfbwt-f/include/marker.hpp:11
ALE_SHIFT = 60                  # pfbwt-f/include/marker.hpp:13
NO_MARKERS = "no markers (consider building the marker array with a larger window size)"


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "rlbwt_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", LIB, src])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint64, C.c_uint64, u8p, u64p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_tsa.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p]
        L.orc_set_markers.argtypes = [C.c_void_p] + [C.c_uint64] * 3 + [C.c_uint64, u64p] * 4
        for name, args in {
            "orc_rank": [C.c_void_p, C.c_uint64, C.c_uint8],
            "orc_select": [C.c_void_p, C.c_uint64, C.c_uint8],
            "orc_run_of_position": [C.c_void_p, C.c_uint64],
            "orc_phi": [C.c_void_p, C.c_uint64],
            "orc_last_run_sample": [C.c_void_p],
            "orc_F": [C.c_void_p, C.c_int],
            "orc_lf_steps": [C.c_void_p],
            "orc_locate_range": [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u64p, C.c_uint64],
            "orc_markers_at": [C.c_void_p, C.c_uint64, u64p, C.c_uint64],
            "orc_markers_at_range": [C.c_void_p, C.c_uint64, C.c_uint64, u64p, C.c_uint64],
        }.items():
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = args
        L.orc_access.restype = C.c_uint8
        L.orc_access.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_find_ranges.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_int, u64p, u64p, u64p]
        L.orc_rb_markers.restype = C.c_int
        L.orc_rb_markers.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     u64p, u64p, u64p, C.c_void_p, C.c_uint64, u64p, C.c_uint64, u64p, u64p]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(u64p)


def pack_reads(reads):
    """list of bytes / 2-D uint8 array -> (bases uint8[], offsets uint64[n+1])"""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, m = reads.shape
        return np.ascontiguousarray(reads).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(m))
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return bases, offs


class OracleIndex:
    """Flat-array model of RowBowt<rle_string_sd> (+ToeholdSA, +MarkerArray, +DocList)."""

    def __init__(self, bwt: F.Rlbwt, tsa: F.Toehold | None = None, ma: F.MarkerWindows | None = None,
                 docs=None):
        L = lib()
        self.bwt, self.tsa, self.ma, self.docs = bwt, tsa, ma, docs
        heads = np.ascontiguousarray(bwt.heads, dtype=np.uint8)
        lens = np.ascontiguousarray(bwt.lens, dtype=np.uint64)
        self.h = L.orc_create(bwt.n, bwt.R, heads.ctypes.data_as(u8p), _p64(lens))
        if not self.h:
            raise ValueError("inconsistent run arrays")
        if tsa is not None:
            a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (tsa.pred, tsa.samples_last, tsa.pred_to_run)]
            L.orc_set_tsa(self.h, tsa.r, *[_p64(x) for x in a])
        if ma is not None:
            a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (ma.starts, ma.ends, ma.idxs, ma.arr)]
            L.orc_set_markers(self.h, ma.size_starts, ma.size_ends, ma.size_idxs,
                              len(a[0]), _p64(a[0]), len(a[1]), _p64(a[1]), len(a[2]), _p64(a[2]),
                              len(a[3]), _p64(a[3]))
        self.n = bwt.n

    @classmethod
    def open(cls, prefix: str, sa: bool = False, markers: bool = False) -> "OracleIndex":
        """load_rowbowt (include/rowbowt_io.hpp:176-189) with rb_align's flags (src/rb_align.cpp:147-160)."""
        bwt = F.read_rbwt(prefix + ".rbwt")
        tsa = F.read_tsa(prefix + ".tsa") if sa else None
        docs = F.read_docs(prefix + ".docs") if sa else None
        ma = F.read_mab(prefix + ".mab") if markers else None
        return cls(bwt, tsa, ma, docs)

    def __del__(self):
        try:
            if self.h:
                lib().orc_free(self.h)
                self.h = None
        except Exception:
            pass

    # scalar probes
    def rank(self, i, c): return lib().orc_rank(self.h, i, c)
    def select(self, i, c): return lib().orc_select(self.h, i, c)
    def access(self, i): return lib().orc_access(self.h, i)
    def phi(self, i): return lib().orc_phi(self.h, i)
    def F(self, c): return lib().orc_F(self.h, c)
    def lf_steps(self): return lib().orc_lf_steps(self.h)

    def find_ranges(self, reads, toehold: bool = False):
        bases, offs = pack_reads(reads)
        n = len(offs) - 1
        lo = np.zeros(n, np.uint64); hi = np.zeros(n, np.uint64); k = np.zeros(n, np.uint64)
        lib().orc_find_ranges(self.h, bases.ctypes.data_as(u8p), _p64(offs), n, int(toehold),
                              _p64(lo), _p64(hi), _p64(k))
        return lo, hi, k

    def build_ftab(self, k: int):
        """RowBowt::build_ftab(k), include/rowbowt.hpp:726-743: find_range of every k-mer; k-mer x spells
        base i as "ACGT"[(x >> 2i) & 3].  Returns (kmers uint8[4^k, k], lo, hi); absent k-mers are (1,0)
        (the reference leaves them out of the map, :735-737)."""
        x = np.arange(4 ** k, dtype=np.uint64)
        codes = np.stack([(x >> np.uint64(2 * i)) & np.uint64(3) for i in range(k)], axis=1)
        kmers = np.frombuffer(b"ACGT", np.uint8)[codes.astype(np.intp)]
        lo, hi, _ = self.find_ranges(kmers)
        return kmers, lo, hi

    def ftab_text(self, k: int) -> bytes:
        """FTab::serialize, include/ftab.hpp:30-34: std::map order = k-mers ascending as strings."""
        kmers, lo, hi = self.build_ftab(k)
        keep = np.nonzero(lo <= hi)[0]
        strs = [kmers[i].tobytes() for i in keep]
        order = sorted(range(len(keep)), key=lambda j: strs[j])
        return b"".join(b"%s %d %d\n" % (strs[j], int(lo[keep[j]]), int(hi[keep[j]])) for j in order)

    def locate(self, lo, hi, k, max_hits=0xFFFFFFFFFFFFFFFF):
        cnt = int(hi) - int(lo) + 1 if hi >= lo else 0
        cnt = min(cnt, max_hits)
        out = np.zeros(max(cnt, 1), np.uint64)
        w = lib().orc_locate_range(self.h, int(lo), int(hi), int(k), max_hits, _p64(out), len(out))
        return out[:w]

    def markers_at_range(self, s, e):
        out = np.zeros(16, np.uint64)
        w = lib().orc_markers_at_range(self.h, int(s), int(e), _p64(out), len(out))
        if w > len(out):
            out = np.zeros(w, np.uint64)
            w = lib().orc_markers_at_range(self.h, int(s), int(e), _p64(out), len(out))
        return out[:w]

    def markers_at(self, i):
        out = np.zeros(16, np.uint64)
        w = lib().orc_markers_at(self.h, int(i), _p64(out), len(out))
        if w > len(out):
            out = np.zeros(w, np.uint64)
            w = lib().orc_markers_at(self.h, int(i), _p64(out), len(out))
        return out[:w]

    def rb_markers(self, reads, wsize=19, max_range=1000, min_range=0, ftab_k=0):
        """The default rb_markers worker (src/rb_markers.cpp:347-415) over a batch: both strands of every read through
        get_markers_greedy_seeding (include/rowbowt.hpp:406-482).  Returns (seed_off uint64[2n+1], seeds SEED_DTYPE[],
        words uint64[]): seeds of read i strand s (0 = +, 1 = -) are seeds[seed_off[2i+s]:seed_off[2i+s+1]], the sorted
        unique markers of a seed are words[mk_off : mk_off + mk_cnt].  ftab_k > 0 seeds through build_ftab(ftab_k)."""
        bases, offs = pack_reads(reads)
        n = len(offs) - 1
        ft_lo = ft_hi = None
        if ftab_k:
            _, ft_lo, ft_hi = self.build_ftab(ftab_k)
            ft_lo = np.ascontiguousarray(ft_lo); ft_hi = np.ascontiguousarray(ft_hi)
        seed_off = np.zeros(2 * n + 1, np.uint64)
        seed_cap, word_cap = max(64, 16 * n), max(64, 64 * n)
        while True:
            seeds = np.zeros(seed_cap, SEED_DTYPE)
            words = np.zeros(word_cap, np.uint64)
            ns, nw = C.c_uint64(0), C.c_uint64(0)
            rc = lib().orc_rb_markers(self.h, bases.ctypes.data_as(u8p), _p64(offs), n, wsize, max_range, min_range, ftab_k,
                                      _p64(ft_lo) if ftab_k else None, _p64(ft_hi) if ftab_k else None, _p64(seed_off),
                                      seeds.ctypes.data, seed_cap, _p64(words), word_cap, C.byref(ns), C.byref(nw))
            if rc != 0:
                raise ValueError("rb_markers: the reference exits here (ftab k - 1 > wsize, or a read shorter than k)")
            if ns.value <= seed_cap and nw.value <= word_cap:
                return seed_off, seeds[:ns.value], words[:nw.value]
            seed_cap, word_cap = max(seed_cap, ns.value), max(word_cap, nw.value)

    def rb_markers_text(self, names, reads, **kw) -> str:
        """stdout of rb_markers at -t 1 (MarkerSeed::print_buf, src/rb_markers.cpp:262-272)."""
        seed_off, seeds, words = self.rb_markers(reads, **kw)
        return render_seeds(names, seed_off, seeds, words)

    def resolve_offset(self, i):
        """DocList::doc_and_offset_at, include/doclist.hpp:46-50,77-79."""
        names, starts = self.docs
        rank = int(np.searchsorted(starts, np.uint64(i), side="right"))
        return names[rank - 1], int(i) - int(starts[rank - 1])

    def report(self, names, reads, sa=False, markers=False) -> str:
        """rb_report's stdout for a batch, src/rb_align.cpp:118-145."""
        lo, hi, k = self.find_ranges(reads, toehold=sa)
        out = []
        for i, nm in enumerate(names):
            l, h = int(lo[i]), int(hi[i])
            out.append("%s (%d,%d), count=%d\n" % (nm, l, h, (h - l + 1) & 0xFFFFFFFFFFFFFFFF))
            if sa:
                s = "\tlocs: "
                for x in self.locate(l, h, int(k[i])):
                    d, o = self.resolve_offset(int(x))
                    s += "%d/%s:%d " % (int(x), d, o)
                out.append(s + "\n")
            if markers:
                s = "\tmarkers: "
                ms = self.markers_at_range(l, h)
                if len(ms) == 0:
                    s += NO_MARKERS
                for m in ms:
                    s += "%d/%d " % (int(m) & POS_MASK, int(m) >> ALE_SHIFT)
                out.append(s + "\n")
        return "".join(out)


SEED_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("mk_off", "<u8"), ("qstart", "<u4"), ("qlen", "<u4"),
                       ("mk_raw", "<u4"), ("mk_cnt", "<u4")])
SEQ_MASK, SEQ_SHIFT = 0x0FFFF00000000000, 46     # pfbwt-f/include/marker.hpp:10,12


def render_seeds(names, seed_off, seeds, words) -> str:
    """MarkerSeed::print_buf (src/rb_markers.cpp:262-272) for every seed, reads in input order, + strand first."""
    out = []
    for i, nm in enumerate(names):
        for s in (0, 1):
            for j in range(int(seed_off[2 * i + s]), int(seed_off[2 * i + s + 1])):
                sd = seeds[j]
                qs = int(sd["qstart"])
                if qs == 0xFFFFFFFF:
                    qs = 0xFFFFFFFFFFFFFFFF          # size_t(-1), printed as such by the reference
                line = "%s %d %s %d %d" % (nm, (int(sd["hi"]) - int(sd["lo"]) + 1) & 0xFFFFFFFFFFFFFFFF, "+-"[s], qs, int(sd["qlen"]))
                if sd["mk_cnt"]:
                    o = int(sd["mk_off"])
                    for m in words[o:o + int(sd["mk_cnt"])]:
                        m = int(m)
                        line += " %d/%d/%d" % ((m & SEQ_MASK) >> SEQ_SHIFT, m & POS_MASK, m >> ALE_SHIFT)
                else:
                    line += " ."
                out.append(line + "\n")
    return "".join(out)


def ref_rb_markers(prefix: str, fastq: str, wsize=19, max_range=1000, min_range=0, ftab=False) -> str:
    """stdout of the compiled, unmodified reference rb_markers (oracle/_ref), single worker thread."""
    cmd = [os.path.join(REFBIN, "rb_markers"), "-w", str(wsize), "-r", str(max_range), "-m", str(min_range), "-t", "1"]
    cmd += (["--ftab"] if ftab else []) + [prefix, fastq]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()


def ref_rb_align(prefix: str, fastq: str, sa=False, markers=False) -> str:
    """stdout of the compiled, unmodified reference rb_align (oracle/_ref)."""
    cmd = [os.path.join(REFBIN, "rb_align")] + (["-s"] if sa else []) + (["-m"] if markers else []) + [prefix, fastq]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()


def have_ref() -> bool:
    return os.path.exists(os.path.join(REFBIN, "rb_align"))
