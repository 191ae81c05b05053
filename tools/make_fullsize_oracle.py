#!/usr/bin/env python3
"""Caches the ORACLE's answers (oracle/rlbwt_oracle.c through oracle/oracle.py) for the parity sample of a full-size
workload: the first 300 exact reads (seed 3) + 300 noisy reads (seed 5: 1 % substitutions, 0.1 % N) of
tests/test_full_size.py::test_sample_bit_exact_against_oracle.  Loading the c2 index into the numpy readers takes
minutes, so the arrays are computed once in the build container and committed as
tests/golden/expected/<cfg>.oracle.npz; the GPU test compares against them bit for bit.

Also writes the UNMODIFIED reference's stdout for the noisy reads under all four flag sets
(tests/golden/expected/<cfg>.noisy.<tag>.txt).

  python tools/make_fullsize_oracle.py c2 2000000
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

N_EXACT = 300
N_NOISY = 300
N_NOISY_TEXT = 600


def sample_reads(panel, n_reads):
    k = synth.parity_sample_sizes(panel.nseq)
    exact = synth.make_reads(panel, n_reads, 150, seed=3)[0][:k["exact"]]
    noisy = synth.make_reads(panel, k["noisy"], 150, seed=5, err_rate=0.01, n_rate=0.001)[0]
    return exact, noisy


def noisy_text_reads(panel):
    return synth.make_reads(panel, synth.parity_sample_sizes(panel.nseq)["noisy_text"], 150, seed=5, err_rate=0.01, n_rate=0.001)[0]


def main():
    cfg, n_reads = sys.argv[1], int(sys.argv[2])
    prefix = os.path.join(ROOT, "data", cfg, cfg)
    exp = os.path.join(ROOT, "tests", "golden", "expected")
    panel = synth.make_panel(*synth.CONFIGS[cfg])
    has_sa, has_ma = os.path.exists(prefix + ".tsa"), os.path.exists(prefix + ".mab")

    # reference stdout over the noisy reads, every flag set
    noisy600 = noisy_text_reads(panel)
    with tempfile.TemporaryDirectory() as td:
        fq = os.path.join(td, "noisy.fq")
        synth.write_fastq(noisy600, fq)
        for tag, flags, ok in (("count", [], True), ("s", ["-s"], has_sa), ("m", ["-m"], has_ma), ("sm", ["-s", "-m"], has_sa and has_ma)):
            if not ok:
                continue
            out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "rb_align")] + flags + [prefix, fq],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            open(os.path.join(exp, "%s.noisy.%s.txt" % (cfg, tag)), "wb").write(out)
            print(cfg, "noisy", tag, len(out), "bytes", flush=True)

    if "--text-only" in sys.argv:
        return
    from oracle import oracle as O
    exact, noisy = sample_reads(panel, n_reads)
    seqs = [bytes(x) for x in exact] + [bytes(x) for x in noisy]
    orc = O.OracleIndex.open(prefix, sa=has_sa, markers=has_ma)
    lo, hi, k = orc.find_ranges(seqs, toehold=has_sa)
    loc_off, locs, mk_off, mks = [0], [], [0], []
    for i in range(len(seqs)):
        if has_sa:
            locs.append(np.asarray(orc.locate(lo[i], hi[i], k[i]), dtype=np.uint64))
            loc_off.append(loc_off[-1] + len(locs[-1]))
        if has_ma:
            mks.append(np.asarray(orc.markers_at_range(lo[i], hi[i]), dtype=np.uint64))
            mk_off.append(mk_off[-1] + len(mks[-1]))
    np.savez_compressed(os.path.join(exp, "%s.oracle.npz" % cfg), n_reads=np.uint64(n_reads), lo=lo, hi=hi,
                        k=k if has_sa else np.zeros(0, np.uint64),
                        loc_off=np.asarray(loc_off, np.uint64), locs=np.concatenate(locs) if locs else np.zeros(0, np.uint64),
                        mk_off=np.asarray(mk_off, np.uint64), markers=np.concatenate(mks) if mks else np.zeros(0, np.uint64))
    print(cfg, "oracle sample:", len(seqs), "reads,", int(loc_off[-1]), "locs,", int(mk_off[-1]), "marker words")
    json.dump({"n_reads": n_reads, "n_exact": len(exact), "n_noisy": len(noisy), "has_sa": has_sa, "has_ma": has_ma},
              open(os.path.join(exp, "%s.oracle.json" % cfg), "w"))


if __name__ == "__main__":
    main()
