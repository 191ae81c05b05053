#!/usr/bin/env python3
"""Adds the .ftab fixtures to tests/golden/expected (run in the build container, where
oracle/_ref has been built from /root/reference): the UNMODIFIED reference
`rb_build --ftab-only -k K` (include/rowbowt_io.hpp:128-144 -> RowBowt::build_ftab ->
FTab::serialize) on the committed .rbwt fixtures.  Small k: the text itself; k = 10 (the
reference default): its sha256."""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref")
CASES = [("toy", "small.fa", (4, 6, 10)), ("tiny", "tiny", (5, 10)), ("greedy", "ref.fa", (7,))]
KEEP_TEXT_UP_TO = 6


def ref_ftab(rbwt, k):
    with tempfile.TemporaryDirectory() as td:
        pre = os.path.join(td, "x")
        shutil.copy(rbwt, pre + ".rbwt")
        subprocess.run([os.path.join(BIN, "rb_build"), "-a", "-k", str(k), "-o", pre, pre], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return open(pre + ".ftab", "rb").read()


def main():
    exp = os.path.join(HERE, "expected")
    sums = {}
    for d, pre, ks in CASES:
        for k in ks:
            txt = ref_ftab(os.path.join(HERE, d, pre + ".rbwt"), k)
            name = "%s.k%d.ftab" % (d, k)
            sums[name] = {"sha256": hashlib.sha256(txt).hexdigest(), "lines": txt.count(b"\n"), "bytes": len(txt)}
            if k <= KEEP_TEXT_UP_TO:
                open(os.path.join(exp, name), "wb").write(txt)
    json.dump(sums, open(os.path.join(exp, "ftab_sha256.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(sums, indent=1))


if __name__ == "__main__":
    main()
