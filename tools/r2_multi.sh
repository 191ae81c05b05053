#!/bin/bash
# Runs ON THE GPU BOX with N GPUs (gpurun --gpus N): PCIe sweep, the default bench under torchrun, rb_align --gpus N at the binary level.
N=${1:-2}; mkdir -p gpurun_out; O=gpurun_out; T=${2:-r2m}
nvidia-smi -L | wc -l > $O/${T}_${N}gpu_box.txt; nproc >> $O/${T}_${N}gpu_box.txt; free -g | head -2 >> $O/${T}_${N}gpu_box.txt
nvidia-smi topo -m >> $O/${T}_${N}gpu_box.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for n in 1 $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 tools/h2d_sweep.py 2>/dev/null | grep pcie_sweep | tee -a $O/${T}_pcie_sweep_${N}gpu.jsonl
done
timeout 1500 $RUN bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_bench_c2_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err || tail -20 $O/${T}_bench_${N}gpu.err
python tools/bench_summary.py $O/${T}_bench_c2_${N}gpu.json
RBG_HOST_STATS=1 timeout 1200 python tools/e2e_binaries.py --config c2 --reads 10000000 --ref-reads 10000 --gpus $N --skip-locate --out $O/${T}_e2e_binaries_${N}gpu.json 2>&1 | cut -c1-700 | tail -6
