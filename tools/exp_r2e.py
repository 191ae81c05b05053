#!/usr/bin/env python3
"""Runs ON THE GPU BOX: (1) latency of one rbg_query_packed call against the batch size (what rb_align's GPU worker pays per
parsed chunk), alone and with 2 / 4 host threads calling on the same handle; (2) search kernel variants on the BASELINE
batch (one thread per read / lane pairs).  One JSON line per measurement."""
import concurrent.futures as cf
import ctypes as C
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

lib = rb.lib()
prefix = os.path.join(ROOT, "data", "c2", "c2")
panel = synth.make_panel(*synth.CONFIGS["c2"])
ix = rb.GpuIndex.open(prefix, sa=True, markers=False)
ix.build_ftab(10)


def pinned(nbytes, dtype):
    p = lib.rbg_host_alloc(nbytes)
    return np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=dtype)


for n in (10_000, 53_000, 212_000, 1_000_000):
    reads = synth.make_reads(panel, n, 150, seed=3)[0]
    bases, offs = pinned(n * 150 + 64, np.uint8), pinned((n + 1) * 8, np.uint64)
    bases[:n * 150] = reads.reshape(-1)
    offs[:] = np.arange(n + 1, dtype=np.uint64) * np.uint64(150)
    packed, flags = pinned(((n * 150 + 31) // 32 + 2) * 8, np.uint64), pinned(n + 8, np.uint8)
    pb, keep = ix.pack((bases[:n * 150], offs), threads=4, out=(packed, flags))
    for mode, name in ((0, "count"), (RBG_LOCATE | RBG_NARROW_LOCS, "locate")):
        for threads in (1, 2, 4):
            reps = max(8, min(200, 4_000_000 // n))

            def loop(_):
                for _ in range(reps):
                    ix.query_raw(pb, mode)
            loop(0)
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(threads) as ex:
                list(ex.map(loop, range(threads)))
            dt = time.perf_counter() - t0
            print(json.dumps({"kind": "call", "mode": name, "reads": n, "host_threads": threads, "calls": reps * threads,
                              "ms_per_call": dt * 1e3 / reps, "reads_per_s": n * reps * threads / dt}), flush=True)

n_reads = 10_000_000
sets = {"exact": synth.make_reads(panel, n_reads, 150, seed=3)[0],
        "noisy": synth.make_reads(panel, n_reads, 150, seed=5, err_rate=0.01, n_rate=0.001)[0]}
want = {}
for name, reads in sets.items():
    st = ix.upload(reads)
    for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0:4", "1:4", "1:5"]):
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
                              "ms_min": float(np.min(ms)), "lf_steps": s.lf_steps, "lf_lines": s.lf_lines, "checksum": cs,
                              "same_digest": cs == ref}), flush=True)
    st.free()
ix.close()
