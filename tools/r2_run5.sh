#!/bin/bash
# Runs ON THE GPU BOX: parity tests (c2 + c5s), the default bench (all legs incl. c5) with its wall time, binary-level end to end with host stage clocks.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2e}
nproc > $O/${T}_box.txt; free -g | head -2 >> $O/${T}_box.txt; nvidia-smi -L >> $O/${T}_box.txt
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
s=$(date +%s)
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
echo "bench.py default run: $(( $(date +%s) - s )) s wall" | tee -a $O/${T}_box.txt
python tools/bench_summary.py $O/${T}_bench.json
RBG_HOST_STATS=1 timeout 1200 python tools/e2e_binaries.py --config c2 --reads 10000000 --ref-reads 20000 --out $O/${T}_e2e_binaries.json 2>&1 | tail -12
