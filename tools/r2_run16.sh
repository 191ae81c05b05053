#!/bin/bash
# Runs ON THE GPU BOX: windows larger than the ladder's choice (smaller directory, more collapsed stretches), new cache test.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2t}
( timeout 300 python -m pytest tests/test_round2_gpu.py -x -q -k "layout_cache" 2>&1 | tail -4 ) | tee $O/${T}_pytest.log
timeout 600 python tools/exp_tlb_cliff.py 5:0 5:1024 5:1280 5:1408 5:1536 5:1792 5:2048 5:2560 > $O/${T}_windows.jsonl 2> $O/${T}_windows.err || tail -5 $O/${T}_windows.err
cut -c1-420 $O/${T}_windows.jsonl
