"""Differential fuzzing of the FASTQ front end, no GPU needed: random FASTQ-like inputs through `rb_align --parse-only` with the
chunk parser (plain and BGZF input, random chunk / block sizes and view margins) against the sequential reader (--threads 1).

  python tools/fuzz_parsers.py [seed] [iterations]
"""
import os, random, subprocess, sys, struct, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RB = os.environ.get('RBG_FUZZ_BIN', os.path.join(ROOT, 'rowbowt_b200', 'rb_align'))      # e.g. an ASan / UBSan build of rb_align_main.cpp
TMP = os.environ.get('FUZZ_TMP', '/tmp/rbg_fuzz')
os.makedirs(TMP, exist_ok=True)

def bgzf(data, block):
    out = bytearray()
    chunks = [data[a:a + block] for a in range(0, len(data), block)] + [b""]
    for chunk in chunks:
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(cd) + 8 - 1)
        out += cd + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out)

def gen(rng):
    recs = []
    n = rng.randint(1, 400)
    for i in range(n):
        m = rng.randint(1, 300)
        seq = bytes(rng.choice(b"ACGTN") for _ in range(m))
        qual = bytes(rng.choice(b"@+I!~>") for _ in range(m))
        name = b"r%d" % i + (b" " + bytes(rng.choice(b"abc @+>") for _ in range(rng.randint(0, 8))) if rng.random() < 0.3 else b"")
        kind = rng.random()
        if kind < 0.80:
            rec = b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n"
        elif kind < 0.84:   # multi-line sequence / quality
            h = m // 2
            rec = b"@" + name + b"\n" + seq[:h] + b"\n" + seq[h:] + b"\n+\n" + qual[:h] + b"\n" + qual[h:] + b"\n"
        elif kind < 0.87:   # CRLF
            rec = b"@" + name + b"\r\n" + seq + b"\r\n+\r\n" + qual + b"\r\n"
        elif kind < 0.90:   # FASTA record
            rec = b">" + name + b"\n" + seq + b"\n"
        elif kind < 0.92:   # blank lines
            rec = b"@" + name + b"\n\n" + seq + b"\n+\n" + qual + b"\n\n"
        elif kind < 0.94:   # '+' line repeating the name
            rec = b"@" + name + b"\n" + seq + b"\n+" + name + b"\n" + qual + b"\n"
        elif kind < 0.96:   # empty sequence
            rec = b"@" + name + b"\n\n+\n\n"
        elif kind < 0.98:   # embedded NUL
            rec = b"@" + name + b"\n" + seq[:m // 2] + b"\x00" + seq[m // 2:] + b"\n+\n" + qual + b"I\n"
        else:               # garbage line between records
            rec = b"garbage " + seq[:10] + b"\n@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n"
        recs.append(rec)
    data = b"".join(recs)
    t = rng.random()
    if t < 0.1:
        data = data[:rng.randint(0, len(data))]          # truncated anywhere
    elif t < 0.2:
        data = data.rstrip(b"\n")                         # no final newline
    return data

def run(path, *extra, env=None):
    p = subprocess.run([RB, "--parse-only", *extra, path], capture_output=True, env=env)
    err = [l for l in p.stderr.decode(errors="replace").splitlines() if l.startswith("ERROR")]
    return p.returncode, p.stdout, err

def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    bad = 0
    for it in range(iters):
        rng = random.Random(seed0 * 100000 + it)
        data = gen(rng)
        open(TMP + "/p.fq", "wb").write(data)
        want = run(TMP + "/p.fq", "--threads", "1")
        for trial in range(3):
            chunk = str(rng.choice([64, 200, 777, 1500, 5000, 40000]))
            got = run(TMP + "/p.fq", "--threads", "4", "--chunk-bytes", chunk)
            if got != want:
                bad += 1; print("MISMATCH plain seed", seed0, "it", it, "chunk", chunk, want[0], got[0], want[2], got[2], len(want[1]), len(got[1])); open(TMP + "/bad_%d_%d.fq" % (seed0, it), "wb").write(data)
            if len(data) > 0:
                block = rng.choice([1, 13, 100, 1000, 65280])
                if block == 1 and len(data) > 3000: block = 13
                open(TMP + "/z.fq.gz", "wb").write(bgzf(data, block))
                env = dict(os.environ)
                if rng.random() < 0.7: env["RBG_VIEW_MARGIN"] = str(rng.choice([0, 1, 50, 300, 2000]))
                got = run(TMP + "/z.fq.gz", "--threads", "4", "--chunk-bytes", chunk, env=env)
                if got != want:
                    bad += 1; print("MISMATCH bgzf seed", seed0, "it", it, "chunk", chunk, "block", block, env.get("RBG_VIEW_MARGIN"), want[0], got[0], want[2], got[2], len(want[1]), len(got[1])); open(TMP + "/badz_%d_%d.fq" % (seed0, it), "wb").write(data)
    print("done seed", seed0, "iters", iters, "bad", bad)


if __name__ == "__main__":
    main()
