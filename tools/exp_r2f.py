#!/usr/bin/env python3
"""Runs ON THE GPU BOX: search_kernel A/B on the BASELINE batch (c2, 10 M x 150 bp exact reads, device-resident), one library
per process (RBG_LIB selects an alt build: make -C rowbowt_b200/csrc alt ALTFLAGS=...).  One JSON line per measurement;
the device digest must be the same for every build."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from rowbowt_b200 import RBG_LOCATE, RBG_NARROW_LOCS  # noqa: E402
from tools import synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
tag = os.environ.get("RBG_LIB", "main").split("/")[-1]
prefix = os.path.join(ROOT, "data", "c2", "c2")
panel = synth.make_panel(*synth.CONFIGS["c2"])
exact = synth.make_reads(panel, n_reads, 150, seed=3)[0]
ix = rb.GpuIndex.open(prefix, sa=True, markers=False)
ix.build_ftab(10)
st = ix.upload(exact)
for name, mode in (("count", 0), ("locate", RBG_LOCATE | RBG_NARROW_LOCS)):
    cs = ix.query_staged(st, mode, checksum=True)
    ms = []
    for _ in range(6):
        ix.query_staged(st, mode)
        s = ix.stats()
        ms.append((s.ms_search, s.ms_phi, s.ms_total))
    a = np.array(ms)
    print(json.dumps({"lib": tag, "kind": name, "reads": n_reads, "ms_search": float(a[:, 0].mean()), "ms_search_min": float(a[:, 0].min()),
                      "ms_phi": float(a[:, 1].mean()), "ms_total": float(a[:, 2].mean()), "lf_steps": s.lf_steps, "lf_lines": s.lf_lines,
                      "checksum": int(cs)}), flush=True)
st.free()
ix.close()
