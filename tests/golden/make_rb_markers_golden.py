#!/usr/bin/env python3
"""Golden stdout of the UNMODIFIED reference rb_markers (oracle/_ref/rb_markers, built from
/root/reference/src/rb_markers.cpp by oracle/Makefile) for SURVEY.md §8(f) row 1 — run in the build
container after make_golden.py.  Writes

  tiny/mixed.fq                       reads with lower-case bases, N/n and other IUPAC / junk bytes
  expected/rb_markers_cases.json      the manifest: fixture, query file, command-line flags, ftab k, output file
  expected/rbm.<case>.txt             stdout at --threads 1 (the only deterministic order the reference has)

--ftab cases run against a scratch copy of the index with <prefix>.ftab = expected/<fixture>.k<k>.ftab
(written by make_ftab_golden.py through `rb_build --ftab-only`).
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import read_fastx  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "rb_markers")
EXP = os.path.join(HERE, "expected")

CASES = [
    # (fixture dir, prefix, fastq, flags, ftab k)
    ("tiny", "tiny", "marked.fq", ["-w", "10"], 0),
    ("tiny", "tiny", "marked.fq", [], 0),                                   # defaults: wsize 19, max-range 1000
    ("tiny", "tiny", "marked.fq", ["-w", "4", "-r", "5"], 0),
    ("tiny", "tiny", "marked.fq", ["-w", "7", "-m", "3"], 0),
    ("tiny", "tiny", "marked.fq", ["-w", "10", "--ftab"], 5),
    ("tiny", "tiny", "marked.fq", ["-w", "4", "-r", "5", "--ftab"], 5),
    ("tiny", "tiny", "marked.fq", ["-w", "10", "--heuristic"], 0),
    ("tiny", "tiny", "marked.fq", ["-w", "10", "--heuristic", "--best-strand-only", "--min-seed-length", "20", "-l", "60"], 0),
    ("tiny", "tiny", "marked.fq", ["-w", "10", "--heuristic", "--clear-conflicting", "--clear-identical", "-l", "60"], 0),
    ("tiny", "tiny", "noisy.fq", ["-w", "10"], 0),
    ("tiny", "tiny", "noisy.fq", ["-w", "10", "--ftab"], 5),
    ("tiny", "tiny", "noisy.fq", ["-w", "10", "--heuristic", "--best-strand-only", "-y", "25", "-l", "100", "--ftab"], 5),
    ("tiny", "tiny", "mixed.fq", ["-w", "10"], 0),
    ("tiny", "tiny", "mixed.fq", ["-w", "6", "--ftab"], 5),
    ("tiny", "tiny", "short.fq", ["-w", "5"], 0),
    ("toy", "small.fa", "simple_query.fq", ["-w", "10"], 0),
    ("toy", "small.fa", "error_query.fq", ["-w", "10"], 0),
    ("toy", "small.fa", "error_query.fq", ["-w", "10", "--ftab"], 4),
    ("toy", "small.fa", "edge_query.fq", ["-w", "3"], 0),
]


def make_mixed():
    names, seqs = read_fastx(os.path.join(HERE, "tiny", "noisy.fq"))
    _, more = read_fastx(os.path.join(HERE, "tiny", "marked.fq"))
    rng = np.random.default_rng(13)
    out = []
    for i, s in enumerate((seqs + more)[:400]):
        a = np.frombuffer(s, np.uint8).copy()
        kind = i % 5
        if kind == 0:
            a |= 0x20                                           # all lower-case
        elif kind == 1:
            m = rng.random(len(a)) < 0.3
            a[m] |= 0x20                                        # mixed case
        elif kind == 2:
            a[rng.integers(0, len(a), 2)] = np.frombuffer(b"Nn", np.uint8)      # N/n map to A
        elif kind == 3:
            a[rng.integers(0, len(a), 3)] = np.frombuffer(b"R-*", np.uint8)     # never match
        out.append(b"@m%d\n%s\n+\n%s\n" % (i, a.tobytes(), b"I" * len(a)))
    open(os.path.join(HERE, "tiny", "mixed.fq"), "wb").write(b"".join(out))


def main():
    make_mixed()
    manifest = []
    with tempfile.TemporaryDirectory() as td:
        for n, (d, pre, fq, flags, k) in enumerate(CASES):
            prefix = os.path.join(HERE, d, pre)
            if k:
                for suf in (".rbwt", ".mab"):
                    shutil.copy(prefix + suf, os.path.join(td, pre + suf))
                shutil.copy(os.path.join(EXP, "%s.k%d.ftab" % (d, k)), os.path.join(td, pre + ".ftab"))
                prefix = os.path.join(td, pre)
            cmd = [BIN, "-t", "1"] + flags + [prefix, os.path.join(HERE, d, fq)]
            out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            name = "rbm.%02d.%s.%s.txt" % (n, d, fq)
            open(os.path.join(EXP, name), "wb").write(out)
            manifest.append({"fixture": d, "prefix": pre, "fastq": fq, "flags": flags, "ftab_k": k, "out": name})
            print(name, len(out.splitlines()), "lines")
    json.dump(manifest, open(os.path.join(EXP, "rb_markers_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
