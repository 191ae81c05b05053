#!/usr/bin/env python3
"""Regenerates tests/golden/raw/: rb_build's raw INPUTS (.bwt/.ssa/.esa/.ma, written by the unmodified
pfbwt-f64 / mps_to_ma of the reference) for the `tiny` synthetic index, next to the .rbwt/.tsa/.mab the
unmodified reference rb_build makes from them.  The GPU builder (rbg_build_index, rb_build) must reproduce
those three files byte for byte.  Run in the build container (needs oracle/_ref)."""
import filecmp
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402


def main():
    out = os.path.join(HERE, "raw")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    tmp = tempfile.mkdtemp()
    panel = synth.make_panel(*synth.CONFIGS["tiny"])
    pre = os.path.join(tmp, "tiny")
    synth.build_index(panel, pre, markers=True, keep_fasta=True)
    for suf in (".rbwt", ".tsa", ".mab"):      # the raw files belong to the committed tiny index
        assert filecmp.cmp(pre + suf, os.path.join(HERE, "tiny", "tiny" + suf), shallow=False), suf
    for suf in (".bwt", ".ssa", ".esa", ".ma"):
        shutil.copy(pre + suf, os.path.join(out, "tiny" + suf))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
