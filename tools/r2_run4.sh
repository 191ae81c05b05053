#!/bin/bash
# Runs ON THE GPU BOX: parity tests with the lane-pair search kernel (default), then the pair / single sweep.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2d}
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1200 python tools/exp_r2d.py c2 10000000 ${2:-0:4,1:4,1:5,1:6,1:8} > $O/${T}_pair.jsonl 2> $O/${T}_pair.err || tail -5 $O/${T}_pair.err
python - <<PY
import json
for ln in open("$O/${T}_pair.jsonl"):
    d = json.loads(ln)
    print("%-8s %-6s pair %d minb %d: search %.2f ms (min %.2f)  lines/step %.3f  digest %s" % (d["kind"], d["reads_set"], d["pair"], d["minb"], d["ms_search"], d["ms_min"], d["lf_lines"] / max(1, d["lf_steps"]), "same" if d["same_digest"] else "DIFFERENT"))
PY
