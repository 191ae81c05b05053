#!/bin/bash
# Runs ON THE GPU BOX: parity tests, search-kernel occupancy variants, binary-level end to end (two GPU workers per device).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2h}
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 900 python tools/exp_r2d.py c2 10000000 ${2:-0:4,0:5,0:9} > $O/${T}_variants.jsonl 2> $O/${T}_variants.err || tail -5 $O/${T}_variants.err
python - <<PY
import json
for ln in open("$O/${T}_variants.jsonl"):
    d = json.loads(ln)
    print("%-8s %-6s pair %d minb %d: search %.2f ms (min %.2f)  digest %s" % (d["kind"], d["reads_set"], d["pair"], d["minb"], d["ms_search"], d["ms_min"], "same" if d["same_digest"] else "DIFFERENT"))
PY
RBG_HOST_STATS=1 timeout 1200 python tools/e2e_binaries.py --config c2 --reads 10000000 --ref-reads 20000 --skip-locate --out $O/${T}_e2e_binaries.json 2>&1 | cut -c1-900 | tail -12
