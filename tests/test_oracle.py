"""Pins the oracle (oracle/rlbwt_oracle.c + oracle/rbformats.py) against
(1) the golden vectors hard-coded in the reference's own tests
    (/root/reference/tests/rb_tests.cpp), (2) stdout of the compiled reference
    rb_align committed under tests/golden/expected, (3) reference-internal
    probes (rank / phi / at / at_range) committed as probes.json, and, when the
    compiled reference is present (oracle/_ref), (4) the live binary.
CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import FIXTURES, GOLDEN, fixture_cases, read_fastx
from oracle import oracle as O
from oracle import rbformats as F

POS_MASK, ALE_SHIFT = O.POS_MASK, O.ALE_SHIFT


@pytest.fixture(scope="module")
def toy():
    return O.OracleIndex.open(os.path.join(GOLDEN, "toy", "small.fa"), sa=True, markers=True)


def test_toy_header_matches_survey(toy):
    assert (toy.bwt.n, toy.bwt.R, toy.bwt.B) == (30031, 7573, 2)
    assert sorted(set(toy.bwt.heads.tolist())) == [1, 65, 67, 71, 84]
    assert len(toy.ma.starts) == 190 and toy.ma.wsize == 10
    assert toy.F(ord("A")) == 1          # terminator sorts first


SIMPLE = ["r1.ref", "r1.sample0.0", "r2.ref", "r2.sample0.0", "r3.ref", "r3.sample0.0"]


def test_count_golden_rb_tests_115_120(toy):
    """tests/rb_tests.cpp:115-120 (one live + five commented, all confirmed in SURVEY §8c)."""
    names, seqs = read_fastx(os.path.join(GOLDEN, "toy", "simple_query.fq"))
    assert names == SIMPLE
    lo, hi, _ = toy.find_ranges(seqs)
    exp = [(24279, 24280), (24175, 24175), (27430, 27432), (27430, 27432), (17409, 17409), (17416, 17417)]
    assert list(zip(lo.tolist(), hi.tolist())) == exp


def test_locate_golden_rb_tests_47_58(toy):
    """tests/rb_tests.cpp:47-58: concatenated locations in emission order."""
    _, seqs = read_fastx(os.path.join(GOLDEN, "toy", "simple_query.fq"))
    lo, hi, k = toy.find_ranges(seqs, toehold=True)
    got = []
    for i in range(len(seqs)):
        got += toy.locate(lo[i], hi[i], k[i]).tolist()
    assert got == [20306, 286, 10296, 11897, 21907, 1887, 11897, 21907, 1887, 4644, 14654, 24664]


def test_marker_golden_rb_tests_131_140(toy):
    _, seqs = read_fastx(os.path.join(GOLDEN, "toy", "simple_query.fq"))
    lo, hi, _ = toy.find_ranges(seqs)
    got = [[(int(m) & POS_MASK, int(m) >> ALE_SHIFT) for m in toy.markers_at_range(lo[i], hi[i])]
           for i in range(len(seqs))]
    assert got == [[(289, 0)], [(289, 1)], [], [], [(4650, 0)], [(4650, 1)]]


def test_kmer_ranges_rb_tests_147_173(toy):
    """FTab goldens: each ftab entry is literally find_range(kmer) (rowbowt.hpp:726-743)."""
    kmers = {b"TTCGTCGTAA": (28942, 28944), b"CCGCGGACAT": (10673, 10675), b"GGCAGGCGGA": (19418, 19423),
             b"TATCGTGGAA": (24272, 24274), b"GTATCGTGGAA": (21142, 21144), b"GGAGATATTG": (19097, 19099),
             b"TGGAGATATTG": (27180, 27182)}
    lo, hi, _ = toy.find_ranges(list(kmers))
    assert list(zip(lo.tolist(), hi.tolist())) == list(kmers.values())


def test_edge_reads(toy):
    lo, hi, _ = toy.find_ranges([b"GGCAGNCGGA", b"ggcaggcgga", b"A", b"\x02", b"\xff"])
    assert list(zip(lo.tolist(), hi.tolist())) == [(1, 0), (1, 0), (1, 7649), (1, 0), (1, 0)]


def test_at_range_quirks_survey_B6(toy):
    """window 0 is skipped in the multi-window branch when E(s)=0."""
    d = lambda s, e: [(int(m) & POS_MASK, int(m) >> ALE_SHIFT) for m in toy.markers_at_range(s, e)]
    assert d(300, 320) == []
    assert d(300, 970) == [(4121, 1)]
    assert d(312, 312) == [(9035, 1)]
    assert d(312, 970) == [(4121, 1)]
    assert d(960, 1130) == [(4121, 1), (9035, 1)]
    assert len(d(0, 29599)) == 189
    assert d(29590, 40000) == [(9035, 0)]
    assert d(1, 0) == []


def test_greedy_fixture():
    ix = O.OracleIndex.open(os.path.join(GOLDEN, "greedy", "ref.fa"), sa=True)
    names, seqs = read_fastx(os.path.join(GOLDEN, "greedy", "query.fq"))
    lo, hi, k = ix.find_ranges(seqs, toehold=True)
    res = dict(zip(names, zip(lo.tolist(), hi.tolist(), k.tolist())))
    assert res["1019_good"][:2] == (12567, 12567) and res["1019_good"][2] == 10000
    assert res["1019_10"] == (1, 0, 0)
    assert ix.resolve_offset(10000) == ("greedy_seeding", 10000)


@pytest.mark.parametrize("d,pre,fq,tag,sa,ma", list(fixture_cases()))
def test_report_matches_reference_stdout(d, pre, fq, tag, sa, ma):
    """Byte-compare the oracle's rendering with the committed stdout of reference rb_align."""
    ix = O.OracleIndex.open(os.path.join(GOLDEN, d, pre), sa=sa, markers=ma)
    names, seqs = read_fastx(os.path.join(GOLDEN, d, fq))
    exp = open(os.path.join(GOLDEN, "expected", "%s.%s.%s.txt" % (d, fq, tag))).read()
    assert ix.report(names, seqs, sa=sa, markers=ma) == exp


@pytest.mark.parametrize("name", ["toy", "tiny", "greedy"])
def test_internal_probes(name):
    """rank / phi / at / at_range of the reference classes themselves (ref_probe)."""
    d, pre, _, has_ma = FIXTURES[name]
    k = json.load(open(os.path.join(GOLDEN, "expected", "probes.json")))[name]
    ix = O.OracleIndex.open(os.path.join(GOLDEN, d, pre), sa=True, markers=has_ma)
    assert ix.n == k["n"] and ix.tsa.r == k["r"]
    assert O.lib().orc_last_run_sample(ix.h) == k["last_run_sample"]
    for c, vals in k["rank"].items():
        assert [ix.rank(i, int(c)) for i in k["rank_pos"]] == vals
    assert [ix.phi(i) for i in k["phi_in"]] == k["phi_out"]
    if has_ma:
        assert [ix.markers_at_range(s, e).tolist() for s, e in k["at_range_in"]] == k["at_range_out"]
        assert [ix.markers_at(i).tolist() for i in k["at_in"]] == k["at_out"]


def test_flat_model_self_consistency():
    """select/rank/access agree with a brute-force expansion of the runs (Appendix B.8)."""
    b = F.read_rbwt(os.path.join(GOLDEN, "greedy", "ref.fa.rbwt"))
    ix = O.OracleIndex(b)
    text = np.repeat(b.heads, b.lens.astype(np.int64))
    rng = np.random.default_rng(5)
    for c in (1, 65, 67, 71, 84):
        cum = np.concatenate([[0], np.cumsum(text == c)])
        for i in rng.integers(0, b.n + 1, 300).tolist() + [0, b.n]:
            assert ix.rank(i, c) == cum[i]
        where = np.nonzero(text == c)[0]
        for t in rng.integers(0, len(where), 100).tolist():
            assert ix.select(t, c) == where[t]
    for i in rng.integers(0, b.n, 300).tolist():
        assert ix.access(i) == text[i]


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("tag,sa,ma", [("count", False, False), ("sm", True, True)])
def test_live_reference_binary(tag, sa, ma, tmp_path):
    """The committed expected files are what the live reference prints today."""
    pre, fq = os.path.join(GOLDEN, "tiny", "tiny"), os.path.join(GOLDEN, "tiny", "noisy.fq")
    exp = open(os.path.join(GOLDEN, "expected", "tiny.noisy.fq.%s.txt" % tag)).read()
    assert O.ref_rb_align(pre, fq, sa=sa, markers=ma) == exp


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_oracle_report_equals_live_reference_on_random_reads(name, tmp_path):
    """Random reads (short random strings with many occurrences, committed reads cut / mutated / N-sprinkled) through the live
    reference with every flag set: the oracle's report is its stdout.  tools/fuzz_oracle_reference.py is the long-running form
    (130 000 read reports over the three fixtures, no difference)."""
    import random
    from tools.fuzz_oracle_reference import make_reads
    d, pre, fqs, has_ma = FIXTURES[name]
    prefix = os.path.join(GOLDEN, d, pre)
    has_sa = os.path.exists(prefix + ".tsa")
    ix = O.OracleIndex.open(prefix, sa=has_sa, markers=has_ma)
    pool = []
    for fq in fqs:
        pool += read_fastx(os.path.join(GOLDEN, d, fq))[1]
    for b in range(2):
        seqs = make_reads(random.Random(900 + b), pool, 150)
        names = ["q%d" % i for i in range(len(seqs))]
        fq = tmp_path / "o.fq"
        fq.write_bytes(b"".join(b"@%s\n%s\n+\n%s\n" % (nm.encode(), s, b"I" * len(s)) for nm, s in zip(names, seqs)))
        for sa, ma in ((False, False), (has_sa, has_ma)):
            assert O.ref_rb_align(prefix, str(fq), sa=sa, markers=ma) == ix.report(names, seqs, sa=sa, markers=ma), (b, sa, ma)


# ---- FTab (include/ftab.hpp, RowBowt::build_ftab) ------------------------------------------------
FTAB_CASES = [("toy", "small.fa", 4), ("toy", "small.fa", 6), ("tiny", "tiny", 5), ("toy", "small.fa", 10),
              ("greedy", "ref.fa", 7)]


@pytest.mark.parametrize("d,pre,k", FTAB_CASES)
def test_oracle_ftab_equals_reference_rb_build_ftab(d, pre, k):
    """oracle build_ftab + serialize == the file the unmodified `rb_build --ftab-only -k K` wrote
    (tests/golden/make_ftab_golden.py): the text for small k, its sha256 for all."""
    import hashlib
    orc = O.OracleIndex.open(os.path.join(GOLDEN, d, pre))
    txt = orc.ftab_text(k)
    name = "%s.k%d.ftab" % (d, k)
    sums = json.load(open(os.path.join(GOLDEN, "expected", "ftab_sha256.json")))
    assert hashlib.sha256(txt).hexdigest() == sums[name]["sha256"]
    path = os.path.join(GOLDEN, "expected", name)
    if os.path.exists(path):
        assert txt == open(path, "rb").read()


def test_ftab_kmer_goldens_rb_tests_147_173(toy):
    """tests/rb_tests.cpp:147-173: ftab lookups equal plain searches (k = 10 entries of the table)."""
    kmers, lo, hi = toy.build_ftab(10)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    exp = {b"TTCGTCGTAA": (28942, 28944), b"CCGCGGACAT": (10673, 10675), b"GGCAGGCGGA": (19418, 19423),
           b"TATCGTGGAA": (24272, 24274), b"GGAGATATTG": (19097, 19099)}
    for km, rng in exp.items():
        x = sum(code[b] << (2 * i) for i, b in enumerate(km))
        assert kmers[x].tobytes() == km
        assert (int(lo[x]), int(hi[x])) == rng
