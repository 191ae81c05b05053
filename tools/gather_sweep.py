#!/usr/bin/env python3
"""Random-gather roofline sweep: achieved GB/s of useful lines for random
32/64/128-byte line reads over footprints from L2-resident to tens of GB, independent and
dependent (LF-like) address chains.  Prints one JSON line per point."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rowbowt_b200 as rb

lib = rb.lib()
foot = [32 << 20, 128 << 20, 256 << 20, 512 << 20, 1 << 30, 2 << 30, 8 << 30, 32 << 30]
for fb in foot:
    for line in (32, 64, 128):
        for dep in (0, 1):
            it = 256
            g = lib.rbg_gather_roofline(0, fb, line, -it if dep else it)
            print(json.dumps({"footprint_MB": fb >> 20, "line_bytes": line, "dependent": dep, "gbs": round(g, 1),
                              "glines_per_s": round(g / line, 2)}), flush=True)
