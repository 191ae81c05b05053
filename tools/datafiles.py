#!/usr/bin/env python3
"""Transport packing of the generated benchmark indexes under data/.

The snapshot of the repo that travels to the GPU box is capped at 512 MiB and the BASELINE index alone is 440 MB, so the
compressible files (the Huffman-shaped .rbwt and the .mab; NOT the .tsa, which is bit-packed text positions) travel as
zstd frames (`<file>.zst`, level 19: c2.rbwt 86 -> 48 MB) while the raw files are listed in .gpurunignore.  Whoever needs
the indexes calls inflate_data() first (tools/synth.py does at import, so tests, bench.py and the tools all do): every
`data/**/*.zst` without an up-to-date raw sibling is decompressed next to it (atomically: temp file + rename, so
concurrent ranks of a multi-GPU run cannot see a partial file).  The index FILES are unchanged reference formats.

  python tools/datafiles.py pack data/c2/c2.rbwt data/c2/c2.mab ...     # in the build container, once
"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def inflate_data(root=None):
    root = root or os.path.join(ROOT, "data")
    done = []
    for z in glob.glob(os.path.join(root, "**", "*.zst"), recursive=True):
        raw = z[:-4]
        if os.path.exists(raw) and os.path.getmtime(raw) >= os.path.getmtime(z):
            continue
        import pyarrow as pa
        tmp = "%s.tmp.%d" % (raw, os.getpid())
        with pa.input_stream(z, compression="zstd") as src, open(tmp, "wb") as dst:
            while True:
                buf = src.read(64 << 20)
                if not buf:
                    break
                dst.write(buf)
        os.replace(tmp, raw)
        done.append(raw)
    return done


def pack(paths, level=19):
    import pyarrow as pa
    for p in paths:
        data = open(p, "rb").read()
        out = pa.Codec("zstd", compression_level=level).compress(data, asbytes=True)
        with open(p + ".zst.tmp", "wb") as f:
            f.write(out)
        os.replace(p + ".zst.tmp", p + ".zst")
        os.utime(p + ".zst", (os.path.getatime(p), os.path.getmtime(p)))      # not newer than the raw file it was made from
        print("%s: %d -> %d bytes" % (p, len(data), len(out)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "pack":
        pack(sys.argv[2:])
    else:
        print("\n".join(inflate_data()) or "nothing to inflate")
