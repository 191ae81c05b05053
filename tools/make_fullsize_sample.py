#!/usr/bin/env python3
"""Runs the UNMODIFIED reference rb_align (oracle/_ref) over the first 600 reads (120 for -s) of a full-size workload
(tools/synth.py config, reads seed 3) and commits its stdout under tests/golden/expected/<cfg>.sample.<tag>.txt.
The index itself is too large to commit; tests/test_full_size.py regenerates the same reads on the GPU box and
compares the host binary's stdout with these files.  Run in the build container after `tools/synth.py <cfg> data/<cfg>`."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402


def main():
    cfg, n_reads = sys.argv[1], int(sys.argv[2])       # n_reads = the batch size tests/test_full_size.py uses for cfg
    prefix = os.path.join(ROOT, "data", cfg, cfg)
    panel = synth.make_panel(*synth.CONFIGS[cfg])
    reads = synth.make_reads(panel, n_reads, 150, seed=3)[0][:600]     # (the generator is not prefix-stable in n_reads)
    SAMPLE = synth.parity_sample_sizes(panel.nseq)                     # -s prints ~28 B per occurrence: keep that file small
    import json
    json.dump({"n_reads": n_reads}, open(os.path.join(ROOT, "tests", "golden", "expected", "%s.sample.json" % cfg), "w"))
    with tempfile.TemporaryDirectory() as td:
        for tag, flags, suf in (("count", [], ".rbwt"), ("s", ["-s"], ".tsa"), ("m", ["-m"], ".mab")):
            if not os.path.exists(prefix + suf):
                continue
            fq = os.path.join(td, "sample_%s.fq" % tag)
            synth.write_fastq(reads[:SAMPLE[tag]], fq)
            out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "rb_align")] + flags + [prefix, fq],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            open(os.path.join(ROOT, "tests", "golden", "expected", "%s.sample.%s.txt" % (cfg, tag)), "wb").write(out)
            print(cfg, tag, len(out), "bytes")


if __name__ == "__main__":
    main()
