#!/bin/bash
# Runs ON THE GPU BOX: final sources (locate kernels at their measured CTAs per SM): smoke, the whole -m gpu suite, the default bench, then
# one full ncu capture of the narrow locate_kernel for roofline_locate.traffic.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2y}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/${T}_smoke.log
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
python tools/bench_summary.py $O/${T}_bench.json
NCU="ncu --set full --clock-control none --import-source on -f"
# locate_kernel: launches 0,1 = warm-up + timed step with u64 locations, 2.. = the narrow form (2-bit input leg)
timeout 240 $NCU -k regex:locate_kernel -s 3 -c 1 -o $O/${T}_c2_locate_kernel_narrow \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
ls -la $O/${T}_* | tail -8
