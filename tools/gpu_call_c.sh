#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; T=${1:-s8c}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest.log
for L in "" rowbowt_b200/librowbowt_gpu_prev.so; do
  RBG_LIB=$L timeout 300 python bench.py --mode locate --steps 3 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib=$L', d['kernel_ms'], 'e2e_ms', d['e2e']['ms_per_step'], d['checksum'], d['config']['index']['phi_MB'])"
done
bash tools/profile_gpu.sh ${T}_c2_count search_kernel > $O/${T}_profile.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:locate_kernel -s 1 -c 1 -f -o $O/${T}_c2_locate_kernel \
    python bench.py --mode locate --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
ls -la $O | tail
