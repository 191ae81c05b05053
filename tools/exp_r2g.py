#!/usr/bin/env python3
"""Runs ON THE GPU BOX: locate_kernel (narrow locations, 256-bit buffered stores) against resident CTAs per SM (RBG_LOC_CTAS, read at
every launch) on the BASELINE batch, device-resident.  One JSON line per setting."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from rowbowt_b200 import RBG_LOCATE, RBG_NARROW_LOCS  # noqa: E402
from tools import synth  # noqa: E402

def sweep(cfg, n_reads, settings):
    prefix = os.path.join(ROOT, "data", cfg, cfg)
    panel = synth.make_panel(*synth.CONFIGS[cfg])
    exact = synth.make_reads(panel, n_reads, 150, seed=3)[0]
    ix = rb.GpuIndex.open(prefix, sa=True, markers=False)
    ix.build_ftab(10)
    st = ix.upload(exact)
    want = {}
    for ctas in settings:
        os.environ["RBG_LOC_CTAS"] = ctas
        for narrow in (1, 0):
            mode = RBG_LOCATE | (RBG_NARROW_LOCS if narrow else 0)
            cs = ix.query_staged(st, mode, checksum=True)
            want.setdefault(narrow, cs)
            ms = []
            for _ in range(5):
                ix.query_staged(st, mode)
                ms.append(ix.stats().ms_phi)
            s = ix.stats()
            print(json.dumps({"kind": "locate_ctas", "cfg": cfg, "reads": n_reads, "ctas": int(ctas), "narrow": narrow, "ms_phi": float(np.mean(ms)),
                              "ms_phi_min": float(np.min(ms)), "phi_steps": s.phi_steps, "g_phi_per_s": s.phi_steps / float(np.mean(ms)) / 1e6,
                              "same_digest": cs == want[narrow]}), flush=True)
    st.free()
    ix.close()


settings = sys.argv[1].split(",") if len(sys.argv) > 1 else ["4", "5", "6", "7", "8", "4"]
sweep("c2", 10_000_000, settings)
if os.path.exists(os.path.join(ROOT, "data", "c5w", "c5w.tsa")):
    sweep("c5w", 500_000, settings)          # chains of ~2200 steps: locate_draw_kernel, 5-byte locations
