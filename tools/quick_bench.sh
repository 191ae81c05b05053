#!/bin/bash
# Runs ON THE GPU BOX: bench (no CPU baseline, no gather) for the given tag and configs, prints a summary.
#   tools/quick_bench.sh <tag> <mode> <cfg[:reads]>...
TAG=$1; MODE=$2; shift; shift
for spec in "$@"; do
  CFG=${spec%%:*}; READS=${spec##*:}; [ "$READS" = "$CFG" ] && READS=10000000
  OUT=gpurun_out/${TAG}_${CFG}_${MODE}
  timeout 900 python bench.py --no-cpu-baseline --no-gather --config $CFG --reads $READS --mode $MODE --steps 3 --warmup 3 > $OUT.json 2> $OUT.err || tail -5 $OUT.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT.json"))
    print("$CFG $MODE", d["config"]["index"], "kernel_ms", {k: round(v,3) for k,v in d["kernel_ms"].items()}, "LF/s %.3g" % d["roofline"]["lf_steps_per_s"], "lines/step %.3f" % d["roofline"]["lines_per_lf_step"], "reads/s %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "checksum", d["checksum"])
except Exception as e:
    print("$CFG $MODE failed", e)
PY
done
