#!/bin/bash
# A/B: L2 window variants on the count bench; rb_markers / rb_align binaries after the host pipeline changes
mkdir -p gpurun_out; O=gpurun_out; T=${1:-s8d}
timeout 600 python -m pytest tests/test_rb_markers.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for P in 1 2 0; do
  RBG_L2_PIN=$P timeout 300 python bench.py --mode count --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L2_PIN=$P', d['kernel_ms']['ms_search'], 'e2e_ms', d['e2e']['ms_per_step'], d['checksum'], d['config']['index']['l2_persisting_MB'], d['setup_s'])"
done
RBG_L2_PIN=2 timeout 300 python bench.py --mode locate --steps 3 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('locate L2_PIN=2', d['kernel_ms'], d['setup_s'])"
timeout 1200 python tools/e2e_tools.py --markers-config c2 --reads 1000000 --ref-reads 100000 --build-config none --out $O/${T}_e2e_tools.json 2>&1 | tail -4
timeout 1200 python tools/e2e_binaries.py --config c2 --reads 10000000 --ref-reads 20000 --out $O/${T}_e2e_binaries_10m.json 2>&1 | tail -4 | cut -c1-400
