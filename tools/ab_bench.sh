#!/bin/bash
# A/B of the two directory layouts on one GPU box: tools/ab_bench.sh <tag> [bench args...]
TAG=$1; shift
for L in 1 2; do
  RBG_LAYOUT=$L timeout 900 python bench.py --no-cpu-baseline --no-gather "$@" > gpurun_out/${TAG}_layout$L.json 2> gpurun_out/${TAG}_layout$L.err
  tail -2 gpurun_out/${TAG}_layout$L.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_layout$L.json"))
print("layout $L", d["config"]["index"], "search ms", d["kernel_ms"]["ms_search"], "LF/s %.3g" % d["roofline"]["lf_steps_per_s"], "lines/step", d["roofline"]["lines_per_lf_step"], "reads/s %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "checksum", d["checksum"])
PY
done
