#!/usr/bin/env python3
"""Runs ON THE GPU BOX: where the TLB cliff is for search_kernel.  The BASELINE index is
re-laid out with smaller windows (RBG_WINDOW), which inflates its rank directory from 190 MB to ~0.4 / 0.9 / 1.7 GB --
the footprint a real config-5 directory would have -- and the count search of the BASELINE batch is timed on each,
next to the random 64-byte gather ceiling at that footprint.  One JSON line per layout."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from tools import synth  # noqa: E402

lib = rb.lib()
prefix = os.path.join(ROOT, "data", "c2", "c2")
panel = synth.make_panel(*synth.CONFIGS["c2"])
reads = synth.make_reads(panel, 10_000_000, 150, seed=3)[0]
want = None
# default: the TLB-cliff ladder; argv: layout:window pairs (e.g. 5:1280 5:1536: larger windows than the load-time ladder picks)
configs = [tuple(a.split(":")) for a in sys.argv[1:]] or [("5", "0"), ("4", "0"), ("5", "512"), ("5", "256"), ("5", "128"), ("4", "512"), ("4", "256")]
for layout, window in configs:
    os.environ["RBG_LAYOUT"] = layout
    if window == "0":
        os.environ.pop("RBG_WINDOW", None)
    else:
        os.environ["RBG_WINDOW"] = window
    ix = rb.GpuIndex.open(prefix, sa=False, markers=False)
    ix.build_ftab(10)
    info = ix.info()
    st = ix.upload(reads)
    cs = ix.query_staged(st, 0, checksum=True)
    want = want or cs
    ms = []
    for _ in range(3):
        ix.query_staged(st, 0)
        ms.append(ix.stats().ms_search)
    s = ix.stats()
    g = lib.rbg_gather_roofline(0, max(int(info.dir_bytes), 1 << 20), 64, 256)
    print(json.dumps({"kind": "tlb_cliff", "layout": info.layout, "window": info.window, "dir_MB": info.dir_bytes / 1e6, "lines": info.n_lines,
                      "cluster_windows": info.n_cluster, "ms_search": float(np.mean(ms)), "lf_steps": s.lf_steps,
                      "lines_per_step": s.lf_lines / s.lf_steps, "line_GBps": s.lf_lines * 64 / float(np.mean(ms)) / 1e6,
                      "random_gather_64B_GBps_at_footprint": g, "frac_of_gather": s.lf_lines * 64 / float(np.mean(ms)) / 1e6 / g,
                      "same_digest": cs == want}), flush=True)
    st.free()
    ix.close()
