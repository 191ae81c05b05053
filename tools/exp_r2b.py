#!/usr/bin/env python3
"""Runs ON THE GPU BOX: kernel A/B sweeps on the BASELINE index (c2), device-resident batches.
  layouts 4 / 5 (RBG_LAYOUT at open) x count / toehold search, exact and noisy reads, RBG_SEARCH_MINB
  locate_kernel: RBG_LOC_TILE x RBG_LOC_CTAS x narrow / wide
One JSON line per measurement on stdout (-> profiles/r2_*.jsonl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from rowbowt_b200 import RBG_LOCATE, RBG_NARROW_LOCS  # noqa: E402
from tools import synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
what = sys.argv[3] if len(sys.argv) > 3 else "all"
tag = os.environ.get("RBG_LIB", "main").split("/")[-1]
prefix = os.path.join(ROOT, "data", cfg, cfg)
panel = synth.make_panel(*synth.CONFIGS[cfg])
exact = synth.make_reads(panel, n_reads, 150, seed=3)[0]
noisy = synth.make_reads(panel, n_reads, 150, seed=5, err_rate=0.01, n_rate=0.001)[0]


def timed(ix, st, mode, reps=4):
    ix.query_staged(st, mode)
    ms = []
    for _ in range(reps):
        ix.query_staged(st, mode)
        s = ix.stats()
        ms.append((s.ms_search, s.ms_phi, s.ms_total))
    s = ix.stats()
    a = np.array(ms)
    return {"ms_search": float(a[:, 0].mean()), "ms_phi": float(a[:, 1].mean()), "ms_total": float(a[:, 2].mean()),
            "lf_steps": s.lf_steps, "lf_lines": s.lf_lines, "phi_steps": s.phi_steps}


def out(**kw):
    print(json.dumps(dict(lib=tag, cfg=cfg, reads=n_reads, **kw)), flush=True)


for layout in ("5", "4"):
    os.environ["RBG_LAYOUT"] = layout
    ix = rb.GpuIndex.open(prefix, sa=True, markers=False)
    ix.build_ftab(10)
    info = ix.info()
    st_e, st_n = ix.upload(exact), ix.upload(noisy)
    base = dict(layout=info.layout, window=info.window, dir_MB=info.dir_bytes / 1e6, phi_MB=info.phi_bytes / 1e6, toehold_MB=info.toehold_bytes / 1e6)
    if what in ("all", "search"):
        for minb in (("4", "3", "5") if layout == "5" else ("4",)):
            os.environ["RBG_SEARCH_MINB"] = minb
            # the knob is read once per process (static): only the first value takes effect in this process
            out(kind="count", reads_set="exact", minb=minb, **base, **timed(ix, st_e, 0))
            break
        out(kind="count", reads_set="noisy", **base, **timed(ix, st_n, 0))
        out(kind="toehold", reads_set="exact", **base, **timed(ix, st_e, RBG_LOCATE))
        out(kind="toehold", reads_set="noisy", **base, **timed(ix, st_n, RBG_LOCATE))
    if what in ("all", "locate") and layout == "5":
        for tile in ("0", "256", "512", "1024", "2048"):
            for ctas in ("8", "6", "4"):
                for narrow in (1, 0):
                    os.environ["RBG_LOC_TILE"], os.environ["RBG_LOC_CTAS"] = tile, ctas
                    r = timed(ix, st_e, RBG_LOCATE | (RBG_NARROW_LOCS if narrow else 0), reps=3)
                    out(kind="locate", tile=int(tile), ctas=int(ctas), narrow=narrow, ms_phi=r["ms_phi"], phi_steps=r["phi_steps"],
                        g_phi_per_s=r["phi_steps"] / r["ms_phi"] / 1e6)
        os.environ.pop("RBG_LOC_TILE"); os.environ.pop("RBG_LOC_CTAS")
    st_e.free(); st_n.free()
    ix.close()
