#!/usr/bin/env python3
"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

  toy/      the reference's own serialized fixtures tests/data/small.fa.{rbwt,tsa,mab} and
            query files (data, not source), plus the .docs file rb_align -s needs
            (`ref 0 / h1 10010 / h2 20020`)
  greedy/   tests/greedy_seeding/{ref.fa.rbwt,.tsa,.docs,query.fq}
  tiny/     a synthetic 20 kbp x 4 haplotype index built by the UNMODIFIED reference
            builder (tools/synth.py -> oracle/_ref/{pfbwt-f64,rb_build,...}) + reads
  expected/ stdout of the compiled reference rb_align (oracle/_ref/rb_align) for every
            flag combination, and reference-internal probes (ref_probe) as JSON KATs
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")


def run(cmd, inp=None):
    return subprocess.run(cmd, input=inp, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout


def rb_align(prefix, fq, flags):
    return run([os.path.join(BIN, "rb_align")] + flags + [prefix, fq])


def probe(cmd, prefix, lines):
    out = run([os.path.join(BIN, "ref_probe"), cmd, prefix], ("\n".join(lines) + "\n").encode()).decode()
    return out.split("\n")[:len(lines)]


def main():
    exp = os.path.join(HERE, "expected")
    for d in ("toy", "greedy", "tiny", "expected"):
        shutil.rmtree(os.path.join(HERE, d), ignore_errors=True)
        os.makedirs(os.path.join(HERE, d))
    toy = os.path.join(HERE, "toy")
    for f in ("small.fa.rbwt", "small.fa.tsa", "small.fa.mab", "simple_query.fq", "error_query.fq"):
        shutil.copy(os.path.join(REF, "tests/data", f), toy)
    open(os.path.join(toy, "small.fa.docs"), "w").write("ref 0\nh1 10010\nh2 20020\n")
    # edge reads (SURVEY.md §8c "Edge reads"): N, lowercase, single base, empty-ish, long
    open(os.path.join(toy, "edge_query.fq"), "w").write(
        ">n_inside\nGGCAGNCGGA\n>lower\nggcaggcgga\n>single\nA\n>kmer1\nTTCGTCGTAA\n>kmer2\nCCGCGGACAT\n"
        ">kmer3\nGGCAGGCGGA\n>kmer4\nTATCGTGGAA\n>kmer5\nGTATCGTGGAA\n>kmer6\nGGAGATATTG\n>kmer7\nTGGAGATATTG\n"
        ">two\nAC\n>polyA\nAAAAAAAAAA\n")
    g = os.path.join(HERE, "greedy")
    for f in ("ref.fa.rbwt", "ref.fa.tsa", "ref.fa.docs", "query.fq"):
        shutil.copy(os.path.join(REF, "tests/greedy_seeding", f), g)

    # tiny synthetic index through the reference builder
    tiny = os.path.join(HERE, "tiny")
    panel = synth.make_panel(*synth.CONFIGS["tiny"])
    synth.build_index(panel, os.path.join(tiny, "tiny"), markers=True)
    reads, _, _ = synth.make_reads(panel, 300, read_len=100, seed=3)
    synth.write_fastq(reads, os.path.join(tiny, "exact.fq"))
    reads, _, _ = synth.make_reads(panel, 300, read_len=100, seed=5, err_rate=0.01, n_rate=0.001)
    synth.write_fastq(reads, os.path.join(tiny, "noisy.fq"))
    # short reads landing right at the start of sequences -> near marker windows & many hits
    reads, _, _ = synth.make_reads(panel, 300, read_len=24, seed=7)
    synth.write_fastq(reads, os.path.join(tiny, "short.fq"))
    # reads whose first base lies within wsize=10 of a panel site -> final range rows carry markers
    rng = np.random.default_rng(9)
    st = np.clip(rng.choice(panel.sites, 300) - rng.integers(0, 10, 300), 0, panel.L - 60)
    hs = rng.integers(0, panel.nseq, 300)
    reads = np.stack([panel.sequence(int(h))[s:s + 60] for h, s in zip(hs, st)])
    synth.write_fastq(reads, os.path.join(tiny, "marked.fq"))

    cases = [("toy", "small.fa", ["simple_query.fq", "error_query.fq", "edge_query.fq"], True),
             ("greedy", "ref.fa", ["query.fq"], False),
             ("tiny", "tiny", ["exact.fq", "noisy.fq", "short.fq", "marked.fq"], True)]
    for d, pre, fqs, has_ma in cases:
        for fq in fqs:
            for flags in ([], ["-s"], ["-m"], ["-s", "-m"]):
                if "-m" in flags and not has_ma:
                    continue
                out = rb_align(os.path.join(HERE, d, pre), os.path.join(HERE, d, fq), flags)
                name = "%s.%s.%s.txt" % (d, fq, "".join(f[1] for f in flags) or "count")
                open(os.path.join(exp, name), "wb").write(out)

    # reference-internal KATs
    rng = np.random.default_rng(11)
    kats = {}
    for d, pre, has_ma in (("toy", "small.fa", True), ("tiny", "tiny", True), ("greedy", "ref.fa", False)):
        prefix = os.path.join(HERE, d, pre)
        first = run([os.path.join(BIN, "ref_probe"), "tsa", prefix]).decode().split("\n")[0].split()
        r, n, last = map(int, first)
        k = {"n": n, "r": r, "last_run_sample": last}
        pos = sorted(set([0, 1, n - 1, n] + rng.integers(0, n + 1, 200).tolist()))
        rk = {}
        for c in (1, 65, 67, 71, 84, 78):
            rk[str(c)] = list(map(int, probe("rank", prefix, ["%d %d" % (i, c) for i in pos])))
        k["rank_pos"], k["rank"] = pos, rk
        ph = sorted(set([0, 1, n - 2] + rng.integers(0, n - 1, 200).tolist()))
        ph = [i for i in ph if i != n - 1]
        # phi(SA[0]) is undefined in the reference (asserts pred_to_run>0): keep only defined inputs
        sa0 = None
        vals = []
        keep = []
        for i in ph:
            try:
                v = probe("phi", prefix, [str(i)])[0]
                vals.append(int(v)); keep.append(i)
            except Exception:
                pass
        k["phi_in"], k["phi_out"] = keep, vals
        if has_ma:
            sz = n
            pairs = [(300, 320), (300, 970), (312, 312), (312, 970), (960, 1130), (0, 29599), (29590, 40000), (1, 0),
                     (0, 0), (0, n - 1), (n - 1, n - 1)]
            a = rng.integers(0, sz, 300); w = rng.integers(0, 400, 300)
            pairs += [(int(x), int(min(x + y, sz - 1))) for x, y in zip(a, w)]
            outs = probe("at_range", prefix, ["%d %d" % p for p in pairs])
            k["at_range_in"] = pairs
            k["at_range_out"] = [list(map(int, o.split())) for o in outs]
            pts = sorted(set(rng.integers(0, sz, 300).tolist() + [312, 313, 964, 965, 966]))
            outs = probe("at", prefix, [str(p) for p in pts])
            k["at_in"], k["at_out"] = pts, [list(map(int, o.split())) for o in outs]
        kats[d] = k
    json.dump(kats, open(os.path.join(exp, "probes.json"), "w"))
    print("golden fixtures written under", HERE)


if __name__ == "__main__":
    main()
