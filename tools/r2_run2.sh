#!/bin/bash
# Runs ON THE GPU BOX: parity tests (new layout, refill kernel), the kernel sweeps of tools/exp_r2b.py (main and alt builds).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2b}
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1500 python tools/exp_r2b.py c2 10000000 all > $O/${T}_sweep_main.jsonl 2> $O/${T}_sweep_main.err || tail -5 $O/${T}_sweep_main.err
RBG_LIB=$PWD/rowbowt_b200/librowbowt_gpu_alt.so timeout 900 python tools/exp_r2b.py c2 10000000 search > $O/${T}_sweep_alt.jsonl 2> $O/${T}_sweep_alt.err || tail -5 $O/${T}_sweep_alt.err
RBG_SEARCH_MINB=5 timeout 900 python tools/exp_r2b.py c2 10000000 search > $O/${T}_sweep_minb5.jsonl 2> $O/${T}_sweep_minb5.err
RBG_SEARCH_MINB=3 timeout 900 python tools/exp_r2b.py c2 10000000 search > $O/${T}_sweep_minb3.jsonl 2> $O/${T}_sweep_minb3.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_sweep_*.jsonl")):
    print("==", f)
    for ln in open(f):
        d = json.loads(ln)
        if d["kind"] == "locate":
            print("locate tile %4d ctas %d narrow %d: %.2f ms  %.1f G phi/s" % (d["tile"], d["ctas"], d["narrow"], d["ms_phi"], d["g_phi_per_s"]))
        else:
            print("%-8s %-6s layout %d W %d dir %.0f MB: search %.2f ms  steps %.3g lines/step %.3f" % (d["kind"], d["reads_set"], d["layout"], d["window"], d["dir_MB"], d["ms_search"], d["lf_steps"], d["lf_lines"] / max(1, d["lf_steps"])))
PY
ls -la $O | tail -8
