#!/usr/bin/env python3
"""Runs ON THE GPU BOX: the chunk-boundary race of the pipelined rbg_query (raw bytes in, pack_kernel per chunk on two alternating
streams) made visible.  With RBG_TEST_UNORDERED_PACKS the packs of neighbouring chunks are not ordered (the state before the
fix); RBG_TEST_STALL forces the bad interleaving.  Prints how many reads come back wrong in each configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rowbowt_b200 as rb  # noqa: E402
from conftest import GOLDEN, read_fastx  # noqa: E402

prefix = os.path.join(GOLDEN, "tiny", "tiny")
seqs = []
for fq in ("exact.fq", "noisy.fq", "marked.fq"):
    seqs += read_fastx(os.path.join(GOLDEN, "tiny", fq))[1]
seqs = [s[:75 + (i % 9)] for i, s in enumerate(seqs) if len(s) >= 90][:400]
ix = rb.GpuIndex.open(prefix, sa=True)
os.environ["RBG_CHUNKS"] = "1"
want = ix.query(seqs, rb.RBG_LOCATE)
os.environ["RBG_CHUNKS"] = "16"
for unordered in (False, True):
    if unordered:
        os.environ["RBG_TEST_UNORDERED_PACKS"] = "1"
    for stall in ("0,0", "2000,4000"):
        os.environ["RBG_TEST_STALL"] = stall
        bad = []
        for _ in range(3):
            got = ix.query(seqs, rb.RBG_LOCATE)
            bad.append(int(((got.lo != want.lo) | (got.hi != want.hi)).sum()))
        print(json.dumps({"kind": "pack_race", "packs_ordered": not unordered, "stall_us": stall, "reads": len(seqs), "wrong_reads_per_call": bad}), flush=True)
ix.close()
