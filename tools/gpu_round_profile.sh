#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/gpu_round_profile.sh <tag>'): ncu launch list + full capture of search_kernel and locate_kernel on c2, then the
# binary-level end-to-end records (rb_align / rb_markers / rb_build next to the reference binaries).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-s8}
bash tools/profile_gpu.sh ${T}_c2_count search_kernel > $O/${T}_profile.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:locate_kernel -s 1 -c 1 -f -o $O/${T}_c2_locate_kernel \
    python bench.py --mode locate --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
timeout 1200 python tools/e2e_binaries.py --config c2 --reads 2000000 --ref-reads 40000 --out $O/${T}_e2e_binaries.json 2>&1 | tail -8
timeout 1500 python tools/e2e_tools.py --markers-config c2 --reads 1000000 --ref-reads 100000 --build-config none --out $O/${T}_e2e_tools.json 2>&1 | tail -8
ls -la $O | tail -15
