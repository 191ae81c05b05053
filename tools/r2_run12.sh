#!/bin/bash
# Runs ON THE GPU BOX: parity tests, default bench, the TLB-cliff experiment.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2n}
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
python tools/bench_summary.py $O/${T}_bench.json
timeout 900 python tools/exp_tlb_cliff.py > $O/${T}_tlb_cliff.jsonl 2> $O/${T}_tlb_cliff.err || tail -5 $O/${T}_tlb_cliff.err
cut -c1-330 $O/${T}_tlb_cliff.jsonl
