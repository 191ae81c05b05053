#!/bin/bash
# Runs ON THE GPU BOX: layout-cache tests + open timings, compute-sanitizer over the small parity tests, search_kernel A/B (RBG_DUP_LOAD).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2q}
( timeout 600 python -m pytest tests/test_round2_gpu.py -x -q -k "layout_cache" 2>&1 | tail -8 ) | tee $O/${T}_pytest_cache.log
timeout 300 python tools/exp_open_cache.py c2 > $O/${T}_open_cache.jsonl 2> $O/${T}_open_cache.err || tail -5 $O/${T}_open_cache.err
cut -c1-400 $O/${T}_open_cache.jsonl
( timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_rb_markers.py tests/test_rb_build.py -x -q -m gpu \
    -k "query_matches_oracle or edge_reads or max_hits or wide_positions or pipelined_locate or ftab_seeded or fbb_index or markers or build" 2>&1 | tail -25 ) > $O/${T}_memcheck.log
tail -12 $O/${T}_memcheck.log
timeout 300 python tools/exp_r2f.py > $O/${T}_ab_main.jsonl 2> $O/${T}_ab_main.err || tail -5 $O/${T}_ab_main.err
RBG_LIB=$PWD/rowbowt_b200/librowbowt_gpu_alt.so timeout 300 python tools/exp_r2f.py > $O/${T}_ab_alt.jsonl 2> $O/${T}_ab_alt.err || tail -5 $O/${T}_ab_alt.err
cat $O/${T}_ab_main.jsonl $O/${T}_ab_alt.jsonl | cut -c1-300
