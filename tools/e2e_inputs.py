#!/usr/bin/env python3
"""Binary-level, ON THE GPU BOX: the same reads as a plain FASTQ, a one-stream .gz and a BGZF (bgzip-style) file through
rb_align -- report identical (sha256), reads/s of the query phase -- and the start-up with / without --layout-cache.
One JSON line per run (-> profiles/)."""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402


def write_bgzf(src, dst, level=1, block=65280):
    with open(src, "rb") as f, open(dst, "wb") as o:
        while True:
            chunk = f.read(block)
            if not chunk:
                break
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            cd = c.compress(chunk) + c.flush()
            o.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(cd) + 8 - 1))
            o.write(cd + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        o.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))      # the empty end-of-file block


def write_gz(src, dst, level=1):
    c = zlib.compressobj(level, zlib.DEFLATED, 31)
    with open(src, "rb") as f, open(dst, "wb") as o:
        while True:
            chunk = f.read(8 << 20)
            if not chunk:
                break
            o.write(c.compress(chunk))
        o.write(c.flush())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--tmp", default="/tmp/rbg_inputs")
    a = ap.parse_args()
    os.makedirs(a.tmp, exist_ok=True)
    prefix = os.path.join(ROOT, "data", a.config, a.config)
    panel = synth.make_panel(*synth.CONFIGS[a.config])
    reads, _, _ = synth.make_reads(panel, a.reads, 150, seed=3)
    fq = os.path.join(a.tmp, "reads.fq")
    synth.write_fastq(reads, fq)
    t0 = time.perf_counter()
    write_gz(fq, fq + ".gz")
    write_bgzf(fq, fq + ".bgz.gz")
    prep = time.perf_counter() - t0
    ours = os.path.join(ROOT, "rowbowt_b200", "rb_align")
    cache = prefix + ".rbgcache"
    if os.path.exists(cache):
        os.remove(cache)
    want = {}
    for flags in ([], ["-m"]):
        for kind, path in (("plain", fq), ("bgzf", fq + ".bgz.gz"), ("gz", fq + ".gz"), ("plain", fq), ("bgzf", fq + ".bgz.gz")):
            for extra in ([], ["--layout-cache"], ["--layout-cache"]):
                if extra and kind != "plain":
                    continue
                out = os.path.join(a.tmp, "out.txt")
                t0 = time.perf_counter()
                with open(out, "wb") as f:
                    p = subprocess.run([ours] + flags + extra + [prefix, path], stdout=f, stderr=subprocess.PIPE)
                wall = time.perf_counter() - t0
                err = p.stderr.decode(errors="replace").strip().split("\n")
                assert p.returncode == 0, err[-3:]
                load_s, query_s = (float(x) for x in err[-1].split()[:2])
                h = hashlib.sha256(open(out, "rb").read()).hexdigest()[:16]
                tag = " ".join(flags) or "count"
                want.setdefault(tag, h)
                print(json.dumps({"flags": tag, "input": kind, "options": " ".join(extra), "input_MB": round(os.path.getsize(path) / 1e6, 1),
                                  "reads": a.reads, "wall_s": round(wall, 3), "load_s": round(load_s, 3), "query_s": round(query_s, 3),
                                  "reads_per_s_query": round(a.reads / query_s), "report_sha256": h, "same_report": h == want[tag],
                                  "host_cores": os.cpu_count(), "prep_s": round(prep, 1)}), flush=True)
    if os.path.exists(cache):
        os.remove(cache)


if __name__ == "__main__":
    main()
