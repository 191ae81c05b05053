#!/bin/bash
# Runs ON THE GPU BOX: the chunk-boundary race (test + experiment), racecheck again over the pipelined tests, initcheck with its details condensed.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2s}
( timeout 600 python -m pytest tests/test_round2_gpu.py -x -q -k "chunk_boundary or layout_cache or packed_batch_equals" 2>&1 | tail -8 ) | tee $O/${T}_pytest.log
timeout 200 python tools/exp_pack_race.py 2>&1 | tee $O/${T}_pack_race.jsonl
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_round2_gpu.py -x -q -m gpu \
    -k "pipelined_locate or edge_reads or chunk_boundary or narrow_ranges" 2>&1 | tail -30 ) > $O/${T}_racecheck.log
echo "== racecheck"; tail -5 $O/${T}_racecheck.log
timeout 300 compute-sanitizer --tool initcheck --print-limit 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "query_matches_oracle or pipelined_locate" > $O/${T}_initcheck_full.log 2>&1
grep -E "Uninitialized|=========     at |=========     by thread|ERROR SUMMARY|passed|failed" $O/${T}_initcheck_full.log | sed -e 's/by thread.*//' | sort | uniq -c | sort -rn | head -40 > $O/${T}_initcheck.txt
grep -m 3 -B2 -A14 "Uninitialized" $O/${T}_initcheck_full.log | cut -c1-200 >> $O/${T}_initcheck.txt
head -c 2000000 $O/${T}_initcheck_full.log > $O/${T}_initcheck_head.log; rm -f $O/${T}_initcheck_full.log
cat $O/${T}_initcheck.txt | head -80
