#!/bin/bash
# Runs ON THE GPU BOX: parity tests, default bench, ncu launch list of the same command, full ncu captures of the three hot kernels.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2j}
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
s=$(date +%s)
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
echo "bench.py default run: $(( $(date +%s) - s )) s wall" | tee $O/${T}_box.txt
python tools/bench_summary.py $O/${T}_bench.json
# launch list (per-launch durations, cold-cache and serialised) of a short run of the same command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:search_kernel -s 1 -c 1 -o $O/${T}_c2_search_count \
    python bench.py --mode count --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_search.log 2>&1
timeout 900 $NCU -k regex:search_kernel -s 1 -c 1 -o $O/${T}_c2_search_toehold \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_search_toe.log 2>&1
timeout 900 $NCU -k regex:locate_kernel -s 1 -c 1 -o $O/${T}_c2_locate_kernel \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
timeout 900 $NCU -k regex:locate_draw_kernel -s 1 -c 1 -o $O/${T}_c5w_locate_draw_kernel \
    python bench.py --mode count --legs c5 --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate_c5.log 2>&1
ls -la $O | tail -12
