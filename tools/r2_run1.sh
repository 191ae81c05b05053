#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/r2_run1.sh <tag>'): parity tests, the default bench line (all legs), an A/B of the
# toehold search (RBG_LIB=alt build), launch list + full ncu captures of locate_kernel / search_kernel<toehold>.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader | tee $O/${T}_gpu.txt
lscpu | egrep "Model name|^CPU\(s\)|NUMA" | tee -a $O/${T}_gpu.txt
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$O/${T}_bench.json"))
    def show(name, r):
        print(name, "dev %.2f ms" % r["ms_per_step"], "packed-in %.2f ms" % r.get("ms_per_step_packed_input", 0), "kernels", {k: round(v, 2) for k, v in r["kernel_ms"].items()},
              "roof %.3f" % r["roofline"]["frac"], ("loc-roof %.3f (%.1f G phi/s)" % (r["roofline_locate"]["frac"], r["roofline_locate"]["phi_steps_per_s"] / 1e9)) if "roofline_locate" in r else "",
              "e2e %.2f ms" % r["e2e"]["ms_per_step"], "ascii %.2f ms" % r["e2e_ascii"]["ms_per_step"] if "e2e_ascii" in r else "", "cs_eq", r.get("checksum_equal_packed_narrow"))
    show("count", dict(d, ms_per_step_packed_input=1e3 * d["config"]["reads_per_gpu"] / d["value_packed_input"]))
    for k, v in d["legs"].items():
        show(k, v)
    print("cpu", d.get("cpu_baseline"), "clocks", d["clocks"], "host_pack", d["host_pack"], "idx", d["config"]["index"])
except Exception as e:
    print("bench summary failed", repr(e))
PY
# A/B: third rank in every toehold step (alt build) vs only when a lane's range shrank
if [ -f rowbowt_b200/librowbowt_gpu_alt.so ]; then
  RBG_LIB=$PWD/rowbowt_b200/librowbowt_gpu_alt.so timeout 600 python bench.py --mode locate --legs '' --steps 5 --warmup 3 --no-cpu-baseline --no-gather > $O/${T}_bench_alt.json 2> $O/${T}_bench_alt.err
  python -c "import json; d=json.load(open('$O/${T}_bench_alt.json')); print('ALT locate', d['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])"
fi
# launch list of one locate step (shares), then the two kernels in full
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_all.csv \
    python bench.py --mode all --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:locate_kernel -s 1 -c 1 -f -o $O/${T}_c2_locate_kernel \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_locate.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -f -o $O/${T}_c2_search_toehold \
    python bench.py --mode locate --legs '' --steps 1 --warmup 1 --no-cpu-baseline --no-gather > $O/${T}_ncu_search_toe.log 2>&1
ls -la $O | tail -12
