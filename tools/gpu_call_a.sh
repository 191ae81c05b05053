#!/bin/bash
# Runs ON THE GPU BOX: parity tests, A/B of the search kernel, the bench lines of every mode, the reference arm.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-s8}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest.log
for L in "" rowbowt_b200/librowbowt_gpu_prev.so; do
  RBG_LIB=$L timeout 300 python bench.py --mode count --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib=$L', d['kernel_ms'], 'e2e_ms', d['e2e']['ms_per_step'], d['checksum'], d['roofline']['lines_per_lf_step'])"
done
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_c2_count.json 2> $O/${T}_bench_c2_count.err; cut -c1-1500 $O/${T}_bench_c2_count.json
for M in locate markers all; do
  timeout 600 python bench.py --mode $M --steps 3 --warmup 3 --no-gather --cpu-sample 50000 > $O/${T}_bench_c2_$M.json 2> $O/${T}_bench_c2_$M.err
  python -c "import json; d=json.load(open('$O/${T}_bench_c2_$M.json')); print('$M', d['value'], d['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'phi', d['phi_steps_per_step'], 'cpu', d.get('cpu_baseline'))"
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_c2_count_reference.json 2>/dev/null; cat $O/${T}_bench_c2_count_reference.json | cut -c1-600
