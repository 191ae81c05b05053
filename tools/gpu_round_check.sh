#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/gpu_round_check.sh <tag>'): the parity tests, the bench lines of every mode
# and the reference arm; results land in gpurun_out/<tag>_*.  RBG_LIB=<other build of librowbowt_gpu.so> for A/B runs.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-s8}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_c2_count.json 2> $O/${T}_bench_c2_count.err; cut -c1-1500 $O/${T}_bench_c2_count.json
for M in locate markers all; do
  timeout 600 python bench.py --mode $M --steps 3 --warmup 3 --no-gather --cpu-sample 50000 > $O/${T}_bench_c2_$M.json 2> $O/${T}_bench_c2_$M.err
  python -c "import json; d=json.load(open('$O/${T}_bench_c2_$M.json')); print('$M', d['value'], d['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'phi', d['phi_steps_per_step'], 'cpu', d.get('cpu_baseline'))"
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_c2_count_reference.json 2>/dev/null; cat $O/${T}_bench_c2_count_reference.json | cut -c1-600
