#!/bin/bash
# Runs ON THE GPU BOX: parity tests, default bench, full ncu captures of the hot kernels at the final sources (traffic for roofline.traffic).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2k}
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
python tools/bench_summary.py $O/${T}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:search_kernel -s 1 -c 1 -o $O/${T}_c2_search_count \
    python bench.py --mode count --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_search.log 2>&1
timeout 900 $NCU -k regex:search_kernel -s 1 -c 1 -o $O/${T}_c2_search_toehold \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_search_toe.log 2>&1
# locate_kernel: launches 0,1 = warm-up + timed step with u64 locations, 2.. = the narrow form (2-bit input leg)
timeout 900 $NCU -k regex:locate_kernel -s 3 -c 1 -o $O/${T}_c2_locate_kernel_narrow \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
timeout 900 $NCU -k regex:locate_draw_kernel -s 3 -c 1 -o $O/${T}_c5w_locate_draw_kernel_narrow \
    python bench.py --mode count --legs c5 --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate_c5.log 2>&1
ls -la $O | tail -12
