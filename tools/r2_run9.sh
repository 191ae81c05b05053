#!/bin/bash
# Runs ON THE GPU BOX: parity tests (c2 + c5w), the default bench (all legs incl. the config-5 leg) with its wall time.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2i}
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
s=$(date +%s)
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
echo "bench.py default run: $(( $(date +%s) - s )) s wall" | tee $O/${T}_box.txt
python tools/bench_summary.py $O/${T}_bench.json
