#!/usr/bin/env python3
"""Regenerates tests/golden/fbb/: the `tiny` BWT as a wt_fbb index (`rb_build --fbb`, include/fbb_string.hpp) made by
the UNMODIFIED reference builder from tests/golden/raw/tiny.bwt, plus the reference `rb_align --fbb [-m]` stdout for
an edge-case FASTQ.  For the regular query files the reference prints with --fbb exactly what it prints without
(asserted here), so those reuse tests/golden/expected/tiny.*.  Run in the build container (needs oracle/_ref)."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    out = os.path.join(HERE, "fbb")
    keep = {}
    for fn in ("multi.rbwt", "multi.rle.rbwt"):        # only regenerated when data/medium/medium.bwt is there
        if os.path.exists(os.path.join(out, fn)):
            keep[fn] = open(os.path.join(out, fn), "rb").read()
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for fn, data in keep.items():
        open(os.path.join(out, fn), "wb").write(data)
    tmp = tempfile.mkdtemp()
    for suf in (".bwt", ".ma"):
        shutil.copy(os.path.join(HERE, "raw", "tiny" + suf), os.path.join(tmp, "tiny" + suf))
    subprocess.check_call([os.path.join(REF, "rb_build"), "--fbb", "-m", "-o", os.path.join(out, "tiny"), os.path.join(tmp, "tiny")])
    # edge reads: terminator-coded bytes (byte 1 is NOT a symbol of a wt_fbb index), N, lower case, 1-base reads
    exact = open(os.path.join(HERE, "tiny", "exact.fq"), "rb").read().split(b"\n")
    reads = [exact[1], exact[5], b"A", b"C", b"G", b"T", b"N", b"acgt", b"\x01", b"A\x01", b"\x01A", exact[9][:40] + b"\x01" + exact[9][41:],
             exact[13][:75], exact[17][-30:], b"ACGTNACGT", b"\x02\x03", b"\xff"]
    with open(os.path.join(out, "edge.fq"), "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@e%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)))
    ra = os.path.join(REF, "rb_align")
    for fq_dir, fq in (("tiny", "exact.fq"), ("tiny", "noisy.fq"), ("tiny", "short.fq"), ("tiny", "marked.fq"), ("fbb", "edge.fq")):
        for tag, flags in (("count", []), ("m", ["-m"])):
            got = subprocess.run([ra, "--fbb"] + flags + [os.path.join(out, "tiny"), os.path.join(HERE, fq_dir, fq)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            if fq_dir == "tiny":
                assert got == open(os.path.join(HERE, "expected", "tiny.%s.%s.txt" % (fq, tag)), "rb").read(), (fq, tag)
            else:
                open(os.path.join(HERE, "expected", "fbb.%s.%s.txt" % (fq, tag)), "wb").write(got)
    # a text that spans several wt_fbb superblocks (2^20 symbols each) with a partial last block: the first 2 500 123
    # bytes of the `medium` BWT (python tools/synth.py medium data/medium --keep), as a wt_fbb and as the rle_string
    # .rbwt the reference's plain rb_build writes for the same bytes
    src = os.path.join(ROOT, "data", "medium", "medium.bwt")
    if os.path.exists(src):
        with open(os.path.join(tmp, "multi.bwt"), "wb") as f:
            f.write(open(src, "rb").read(2_500_123))
        subprocess.check_call([os.path.join(REF, "rb_build"), "--fbb", "-o", os.path.join(out, "multi"), os.path.join(tmp, "multi")])
        subprocess.check_call([os.path.join(REF, "rb_build"), "-o", os.path.join(tmp, "multi_rle"), os.path.join(tmp, "multi")])
        shutil.copy(os.path.join(tmp, "multi_rle.rbwt"), os.path.join(out, "multi.rle.rbwt"))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
