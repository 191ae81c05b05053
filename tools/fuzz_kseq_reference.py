#!/usr/bin/env python3
"""Differential fuzzing of the kseq-compatible record reader against the UNMODIFIED reference, no GPU needed (needs oracle/_ref):
random FASTQ-like inputs (tools/fuzz_parsers.py's generator) go through `rb_align --parse-only --threads 1` here and through the
reference rb_align over the toy index; the oracle's report of the (name, sequence) pairs parsed here must be the reference's stdout,
with the same exit status and the same ERROR line (truncated quality string / stream error).

  python tools/fuzz_kseq_reference.py [seed] [iterations]
"""
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

from tools.fuzz_parsers import RB, TMP, gen  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    toy = os.path.join(ROOT, "tests", "golden", "toy", "small.fa")
    ix = O.OracleIndex.open(toy)
    ref = os.path.join(O.REFBIN, "rb_align")
    fq = os.path.join(TMP, "r.fq")
    bad = 0
    for it in range(iters):
        data = gen(random.Random(seed * 100000 + it))
        open(fq, "wb").write(data)
        p = subprocess.run([RB, "--parse-only", "--threads", "1", fq], capture_output=True)
        recs = [ln.split(b"\t") for ln in p.stdout.split(b"\n") if ln]
        names = [r[0].decode(errors="replace") for r in recs]
        seqs = [r[1] if len(r) > 1 else b"" for r in recs]
        ours = [ln for ln in p.stderr.decode(errors="replace").splitlines() if ln.startswith("ERROR")]
        q = subprocess.run([ref, toy, fq], capture_output=True)
        theirs = [ln for ln in q.stderr.decode(errors="replace").splitlines() if ln.startswith("ERROR")]
        if ix.report(names, seqs) != q.stdout.decode(errors="replace") or (p.returncode != 0) != (q.returncode != 0) or ours != theirs:
            bad += 1
            keep = os.path.join(TMP, "kseq_mismatch_%d_%d.fq" % (seed, it))
            open(keep, "wb").write(data)
            print("MISMATCH iteration", it, "exit", p.returncode, q.returncode, ours, theirs, "->", keep, flush=True)
    print("done: seed", seed, "iterations", iters, "mismatches", bad)


if __name__ == "__main__":
    main()
