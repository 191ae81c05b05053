#!/usr/bin/env python3
"""Runs ON THE GPU BOX: what opening the BASELINE index costs -- decoded from the reference's files, with RBG_LOAD_CACHE
writing <prefix>.rbgcache, and from that cache -- inside one process (CUDA context up) and as rb_align runs (whole process,
one read).  One JSON line per measurement."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rowbowt_b200 as rb  # noqa: E402
from tools import synth  # noqa: E402  (inflates data/**/*.zst)

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
prefix = os.path.join(ROOT, "data", cfg, cfg)
if os.path.exists(prefix + ".rbgcache"):
    os.remove(prefix + ".rbgcache")
rb.GpuIndex.open(os.path.join(ROOT, "tests", "golden", "tiny", "tiny")).close()        # CUDA context
for name, kw in (("files", {}), ("files+write_cache", {"cache": True}), ("cache", {"cache": True}), ("cache_again", {"cache": True})):
    t0 = time.perf_counter()
    ix = rb.GpuIndex.open(prefix, sa=True, markers=True, **kw)
    dt = time.perf_counter() - t0
    info = ix.info()
    t0 = time.perf_counter()
    ix.build_ftab(10)
    ft = time.perf_counter() - t0
    r = ix.query([b"ACGTACGTAC", b"GATTACA" * 5], rb.RBG_LOCATE | rb.RBG_MARKERS, max_hits=3)
    print(json.dumps({"kind": "open", "cfg": cfg, "how": name, "open_s": round(dt, 3), "ftab10_s": round(ft, 3), "from_cache": info.from_cache,
                      "device_MB": round((info.dir_bytes + info.phi_bytes + info.toehold_bytes + info.marker_bytes) / 1e6, 1),
                      "cache_MB": round(os.path.getsize(prefix + ".rbgcache") / 1e6, 1) if os.path.exists(prefix + ".rbgcache") else 0,
                      "lo": [int(x) for x in r.lo], "hi": [int(x) for x in r.hi]}), flush=True)
    ix.close()
fq = "/tmp/one.fq"
seq = "ACGTACGTACGTAGCTAGCTAGCATCGATCGATCAGCTAGCTAGCATCGATCGATCGATCGA"
open(fq, "w").write("@r\n%s\n+\n%s\n" % (seq, "I" * len(seq)))
align = os.path.join(ROOT, "rowbowt_b200", "rb_align")
outs = {}
for name, extra in (("rb_align", []), ("rb_align --layout-cache", ["--layout-cache"]), ("rb_align --layout-cache (again)", ["--layout-cache"])):
    t0 = time.perf_counter()
    p = subprocess.run([align, "-s", "-m"] + extra + [prefix, fq], capture_output=True, env=dict(os.environ, RBG_HOST_STATS="1"))
    dt = time.perf_counter() - t0
    outs[name] = p.stdout
    print(json.dumps({"kind": "process", "cfg": cfg, "how": name, "wall_s": round(dt, 3), "rc": p.returncode,
                      "load_query_s": p.stderr.decode().splitlines()[-1:]}), flush=True)
assert len(set(outs.values())) == 1, "reports differ"
os.remove(prefix + ".rbgcache")
