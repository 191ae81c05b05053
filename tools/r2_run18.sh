#!/bin/bash
# Runs ON THE GPU BOX with 2 GPUs: rb_align --gpus 2 --layout-cache (two handles opened at once: both miss and write, then both
# open from the cache), same report as --gpus 1 without it; the default bench under torchrun with 2 ranks (count + locate legs).
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2v}
python - <<'PY' 2>&1 | tee $O/${T}_2gpu_cache.jsonl
import hashlib, json, os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from tools import synth
panel = synth.make_panel(*synth.CONFIGS["c2"])
reads = synth.make_reads(panel, 1_000_000, 150, seed=3)[0]
fq = "/tmp/r1m.fq"; synth.write_fastq(reads, fq)
pre = "data/c2/c2"
if os.path.exists(pre + ".rbgcache"): os.remove(pre + ".rbgcache")
want = None
for flags in (["-m"], ["-s", "-m"]):
    want = None
    for extra in (["--gpus", "1"], ["--gpus", "2", "--layout-cache"], ["--gpus", "2", "--layout-cache"], ["--gpus", "2"]):
        t0 = time.perf_counter()
        p = subprocess.run(["rowbowt_b200/rb_align"] + flags + extra + [pre, fq], capture_output=True)
        wall = time.perf_counter() - t0
        h = hashlib.sha256(p.stdout).hexdigest()[:16]
        want = want or h
        err = p.stderr.decode().strip().splitlines()
        print(json.dumps({"flags": " ".join(flags), "options": " ".join(extra), "rc": p.returncode, "wall_s": round(wall, 3), "load_query_s": err[-1] if err else "",
                          "report_bytes": len(p.stdout), "same_report": h == want, "cache_MB": round(os.path.getsize(pre + ".rbgcache") / 1e6, 1) if os.path.exists(pre + ".rbgcache") else 0}), flush=True)
    if os.path.exists(pre + ".rbgcache"): os.remove(pre + ".rbgcache")
print(json.dumps({"stray_tmp_files": [f for f in os.listdir("data/c2") if ".tmp." in f]}))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --legs locate > $O/${T}_bench_2gpu.json 2> $O/${T}_bench_2gpu.err || tail -20 $O/${T}_bench_2gpu.err
python tools/bench_summary.py $O/${T}_bench_2gpu.json
