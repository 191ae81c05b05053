#!/usr/bin/env python3
"""Fuzzing of the index-file readers, layout builders and writers, no GPU needed: copies of the golden fixtures with flipped bytes,
0xFF stretches or a cut tail go through rbg_selftest_layout / _toehold / _phi / _rewrite in a child process; anything but a clean
return (format error or a self-consistent layout) -- i.e. a crashed child -- is reported and the file kept.

  python tools/fuzz_index_files.py [seed] [iterations]
"""
import os
import random
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TMP = os.environ.get("FUZZ_TMP", "/tmp/rbg_fuzz")
CHILD = """
import sys, ctypes as C
sys.path.insert(0, %r)
import rowbowt_b200 as rb
lib, pre = rb.lib(), sys.argv[1].encode()
a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
print(lib.rbg_selftest_layout(pre, 0, 7, C.byref(a), C.byref(b), C.byref(c)),
      lib.rbg_selftest_toehold(pre, 0, C.byref(a), C.byref(b)),
      lib.rbg_selftest_phi(pre, 0, 7, C.byref(a), C.byref(b), C.byref(c)),
      lib.rbg_selftest_rewrite(pre, (sys.argv[1] + ".out").encode(), 7))
""" % ROOT


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    sources = [os.path.join(GOLDEN, "tiny", "tiny"), os.path.join(GOLDEN, "toy", "small.fa")]
    crashes = 0
    for it in range(iters):
        rng = random.Random(seed * 100000 + it)
        src = rng.choice(sources)
        d = os.path.join(TMP, "idx")
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        base = os.path.join(d, "x")
        for suf in (".rbwt", ".tsa", ".mab"):
            shutil.copy(src + suf, base + suf)
        suf = rng.choice([".rbwt", ".tsa", ".mab"])
        b = bytearray(open(base + suf, "rb").read())
        mode = rng.random()
        if mode < 0.3:
            b = b[:rng.randint(0, len(b))]
        elif mode < 0.8:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        else:
            i = rng.randrange(max(1, len(b) - 8))
            b[i:i + 8] = bytes([0xFF] * 8)
        open(base + suf, "wb").write(b)
        p = subprocess.run([sys.executable, "-c", CHILD, base], capture_output=True, timeout=300)
        if p.returncode != 0:
            crashes += 1
            keep = os.path.join(TMP, "crash_%d_%d%s" % (seed, it, suf))
            shutil.copy(base + suf, keep)
            print("CRASH rc", p.returncode, "iteration", it, os.path.basename(src), suf, "->", keep, flush=True)
    print("done: seed", seed, "iterations", iters, "crashes", crashes)


if __name__ == "__main__":
    main()
