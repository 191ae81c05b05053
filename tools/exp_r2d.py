#!/usr/bin/env python3
"""Runs ON THE GPU BOX: search_pair_kernel (two lanes per read) against search_kernel (one thread per read) on the
BASELINE index, device-resident batches: count / toehold x exact / noisy reads x CTAs per SM, with the result digest of
every variant compared against the one-thread-per-read kernel's.  One JSON line per measurement."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from rowbowt_b200 import RBG_LOCATE, RBG_NARROW_LOCS  # noqa: E402
from tools import synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0:4", "1:4", "1:5", "1:6", "1:8"]
prefix = os.path.join(ROOT, "data", cfg, cfg)
panel = synth.make_panel(*synth.CONFIGS[cfg])
sets = {"exact": synth.make_reads(panel, n_reads, 150, seed=3)[0],
        "noisy": synth.make_reads(panel, n_reads, 150, seed=5, err_rate=0.01, n_rate=0.001)[0]}
ix = rb.GpuIndex.open(prefix, sa=True, markers=False)
ix.build_ftab(10)
info = ix.info()
want = {}
for name, reads in sets.items():
    st = ix.upload(reads)
    for v in variants:
        pair, minb = v.split(":")
        os.environ["RBG_SEARCH_PAIR"], os.environ["RBG_SEARCH_MINB"] = pair, minb
        for kind, mode in (("count", 0), ("toehold", RBG_LOCATE | RBG_NARROW_LOCS)):
            cs = ix.query_staged(st, mode, checksum=True)
            ms = []
            for _ in range(4):
                ix.query_staged(st, mode)
                ms.append(ix.stats().ms_search)
            s = ix.stats()
            ref = want.setdefault((name, kind), cs)
            print(json.dumps({"kind": kind, "reads_set": name, "pair": int(pair), "minb": int(minb), "ms_search": float(np.mean(ms)),
                              "ms_min": float(np.min(ms)), "lf_steps": s.lf_steps, "lf_lines": s.lf_lines, "layout": info.layout,
                              "window": info.window, "dir_MB": info.dir_bytes / 1e6, "checksum": cs, "same_digest": cs == ref}), flush=True)
    st.free()
ix.close()
