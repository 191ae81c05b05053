#!/usr/bin/env python3
"""Differential fuzzing of the ORACLE (oracle/: the checker every GPU parity test trusts) against the UNMODIFIED reference binary, no
GPU needed: random reads -- short random ACGT strings (many occurrences: long locate chains, many marker windows), committed reads
cut, mutated and N-sprinkled -- through the reference rb_align with every flag set over the golden indexes; the oracle's report
must be the reference's stdout.

  python tools/fuzz_oracle_reference.py [seed] [batches]
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import FIXTURES, GOLDEN, read_fastx  # noqa: E402
from oracle import oracle as O  # noqa: E402

TMP = os.environ.get("FUZZ_TMP", "/tmp/rbg_fuzz")


def make_reads(rng, pool, n):
    out = []
    for _ in range(n):
        kind = rng.random()
        if kind < 0.35:
            s = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 14)))
        elif kind < 0.5:
            s = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(15, 60)))
        else:
            s = bytearray(rng.choice(pool))
            a = rng.randint(0, max(0, len(s) - 1))
            s = s[a:a + rng.randint(1, len(s))]
            for _ in range(rng.choice([0, 0, 0, 1, 2])):
                if s:
                    s[rng.randrange(len(s))] = rng.choice(b"ACGTNacgt")
            s = bytes(s)
        out.append(s or b"A")
    return out


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    batches = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    os.makedirs(TMP, exist_ok=True)
    bad = 0
    for name, (d, pre, fqs, has_ma) in sorted(FIXTURES.items()):
        prefix = os.path.join(GOLDEN, d, pre)
        has_sa = os.path.exists(prefix + ".tsa")
        ix = O.OracleIndex.open(prefix, sa=has_sa, markers=has_ma)
        pool = []
        for fq in fqs:
            pool += read_fastx(os.path.join(GOLDEN, d, fq))[1]
        for b in range(batches):
            rng = random.Random(seed * 100000 + b)
            seqs = make_reads(rng, pool, 200)
            names = ["q%d" % i for i in range(len(seqs))]
            fq = os.path.join(TMP, "o.fq")
            with open(fq, "wb") as f:
                for nm, s in zip(names, seqs):
                    f.write(b"@%s\n%s\n+\n%s\n" % (nm.encode(), s, b"I" * len(s)))
            for sa, ma in ((False, False), (has_sa, False), (False, has_ma), (has_sa, has_ma)):
                if O.ref_rb_align(prefix, fq, sa=sa, markers=ma) != ix.report(names, seqs, sa=sa, markers=ma):
                    bad += 1
                    print("MISMATCH", name, "batch", b, "sa", sa, "markers", ma, flush=True)
    print("done: seed", seed, "batches", batches, "mismatches", bad)


if __name__ == "__main__":
    main()
