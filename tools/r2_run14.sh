#!/bin/bash
# Runs ON THE GPU BOX: the layout-cache tests, open timings, synccheck / initcheck over the small parity tests, input formats at the binary level.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2r}
( timeout 600 python -m pytest tests/test_round2_gpu.py -x -q -k "layout_cache" 2>&1 | tail -8 ) | tee $O/${T}_pytest_cache.log
timeout 300 python tools/exp_open_cache.py c2 > $O/${T}_open_cache.jsonl 2> $O/${T}_open_cache.err || tail -5 $O/${T}_open_cache.err
cut -c1-300 $O/${T}_open_cache.jsonl
for tool in synccheck initcheck racecheck; do
  ( timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_rb_markers.py -x -q -m gpu \
      -k "query_matches_oracle or edge_reads or pipelined_locate or wide_positions or (markers and toy)" 2>&1 | tail -30 ) > $O/${T}_$tool.log
  echo "== $tool"; tail -6 $O/${T}_$tool.log
done
timeout 600 python tools/e2e_inputs.py --reads 2000000 > $O/${T}_e2e_inputs.jsonl 2> $O/${T}_e2e_inputs.err || tail -5 $O/${T}_e2e_inputs.err
cut -c1-330 $O/${T}_e2e_inputs.jsonl
