#!/usr/bin/env python3
"""Benchmark of the rb_align query path on B200 (BASELINE.json metric: 150bp reads/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode count|locate|markers|all]
                  [--impl reference] [--config c2|medium|small] [--reads R]

One "step" = one pass of the hot path over one batch of synthetic reads (the whole
BASELINE config: 10M x 150bp reads against the 50 Mbp x 64 haplotype pfbwt-f index,
count-only by default = configs[1]).  `value` is measured with the batch resident in HBM;
`e2e` goes through the public C-ABI call rbg_query with pinned HOST buffers (H2D of the
reads and D2H of the results inside the timed region).  With --gpus N (launched under
torchrun) every rank holds an index replica and its own batch: weak scaling, no collective
on the data path; torch.distributed is used only for the barrier and the
max-over-ranks of the device time.

--impl reference times the UNMODIFIED reference rb_align (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on the host cores, same index, same reads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

READ_LEN = 150
DATA = os.path.join(ROOT, "data")
MODES = {"count": 0, "locate": 1, "markers": 2, "all": 3}


# ------------------------------------------------------------------------------------------
def workload(config: str, log):
    """(name, index prefix, panel).  The index is built once by tools/synth.py through the
    unmodified reference builder and cached under data/ (git-ignored, travels with the repo)."""
    order = [config] if config != "auto" else ["c2", "medium", "small"]
    for cfg in order:
        prefix = os.path.join(DATA, cfg, cfg)
        if os.path.exists(prefix + ".rbwt"):
            L, H = synth.CONFIGS[cfg]
            return cfg, prefix, synth.make_panel(L, H)
    # nothing cached: build the largest config that builds in about a minute
    cfg = order[-1] if config != "auto" else "small"
    if cfg == "c2":
        raise SystemExit("data/c2 is not built: run `python tools/synth.py c2 data/c2` (~45 min CPU)")
    L, H = synth.CONFIGS[cfg]
    panel = synth.make_panel(L, H)
    os.makedirs(os.path.join(DATA, cfg), exist_ok=True)
    prefix = os.path.join(DATA, cfg, cfg)
    print("bench: building %s index with the reference builder ..." % cfg, file=log)
    synth.build_index(panel, prefix, markers=True, log=log)
    return cfg, prefix, panel


def describe(cfg, n_reads):
    L, H = synth.CONFIGS[cfg]
    full = "" if (cfg == "c2" and n_reads == 10_000_000) else " [REDUCED: not the BASELINE config]"
    if cfg == "c5s":
        full = " [1/10-scale stand-in for BASELINE config 5 (64 Mbp x 2504): n > 2^32 rows]"
    return "synthetic %d bp reference x %d haplotypes (pfbwt-f index), %d x %dbp exact reads%s" % (L, H, n_reads, READ_LEN, full)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_ev = gpu, [], threading.Event()

    def run(self):
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        self.stop_ev.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, streaming copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def pinned_array(lib, nbytes, dtype):
    p = lib.rbg_host_alloc(nbytes)
    if not p:
        raise MemoryError("rbg_host_alloc")
    buf = (C.c_uint8 * nbytes).from_address(p)
    return np.frombuffer(buf, dtype=dtype), p


# ------------------------------------------------------------------------------------------
def reference_cmd(prefix, fq, mode):
    cmd = [os.path.join(ROOT, "oracle", "_ref", "rb_align")]
    if mode & 1:
        cmd.append("-s")
    if mode & 2:
        cmd.append("-m")
    return cmd + [prefix, fq]


def run_reference_shards(prefix, reads, mode, procs, tmpdir):
    """The reference rb_align is single-threaded (src/rb_align.cpp:162-193); all-core = one process
    per contiguous shard.  Returns (reads/s, slowest shard's own query seconds)."""
    shards = np.array_split(np.arange(len(reads)), procs)
    fqs = []
    for i, idx in enumerate(shards):
        fq = os.path.join(tmpdir, "shard%d.fq" % i)
        synth.write_fastq(reads[idx[0]:idx[-1] + 1], fq, start_id=int(idx[0]))
        fqs.append(fq)
    ps = [subprocess.Popen(reference_cmd(prefix, fq, mode), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE) for fq in fqs]
    qt = []
    for p in ps:
        err = p.communicate()[1].decode().strip().split("\n")
        if p.returncode != 0:
            raise RuntimeError("reference rb_align failed: %s" % err[-1])
        qt.append(float(err[-1].split()[1]))          # "<index_load_time> <total_query_time>"
    return len(reads) / max(qt), max(qt)


def reference_arm(args, rank, world, log):
    if rank != 0:
        return
    mode = MODES[args.mode]
    base = {"impl": "reference", "metric": "150bp reads/s (%s)" % args.mode, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic"}
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rb_align")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/rb_align not built (run make -C oracle ref where /root/reference exists)"}))
        return
    cfg, prefix, panel = workload(args.config, log)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.ref_procs or cores))
    per_proc = args.ref_reads_per_proc
    n = per_proc * procs
    reads, _, _ = synth.make_reads(panel, n, READ_LEN, seed=3)
    vals = []
    with tempfile.TemporaryDirectory() as td:
        for it in range(args.warmup + args.steps):
            v, _ = run_reference_shards(prefix, reads, mode, procs, td)
            if it >= args.warmup:
                vals.append(v)
    v = float(np.mean(vals))
    sample = "%d reads (%d per process x %d processes) of the same workload per step" % (n, per_proc, procs)
    base.update({"value": v, "ms_per_step": 1e3 * n / v,
                 "config": {"workload": describe(cfg, args.reads), "mode": args.mode, "sample": sample},
                 "cpu_baseline": {"value": v, "unit": "reads/s", "cores": procs, "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="count", choices=sorted(MODES))
    ap.add_argument("--config", default="auto", choices=["auto"] + sorted(synth.CONFIGS))
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="reads timed through the reference for cpu_baseline")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-reads-per-proc", type=int, default=50_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--ftab-k", type=int, default=10, help="k of the k-mer seed table built on the GPU at open (0 = none)")
    args = ap.parse_args()
    log = sys.stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world, log)
        return

    import torch
    import torch.distributed as dist
    import rowbowt_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, in front of the one JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    lib = rb.lib()
    mode = MODES[args.mode]

    cfg, prefix, panel = workload(args.config, log)
    t0 = time.time()
    ix = rb.GpuIndex.open(prefix, sa=bool(mode & 1), markers=bool(mode & 2), device=local)
    t_open = time.time() - t0
    t0 = time.time()
    ix.build_ftab(args.ftab_k)
    t_ftab = time.time() - t0
    info = ix.info()
    n_reads = args.reads
    t0 = time.time()
    reads, _, _ = synth.make_reads(panel, n_reads, READ_LEN, seed=3 + rank)     # a different batch on every rank
    t_reads = time.time() - t0

    # pinned host staging for the e2e leg (the caller's buffers in the C-ABI call)
    bases, _p1 = pinned_array(lib, n_reads * READ_LEN, np.uint8)
    offs, _p2 = pinned_array(lib, (n_reads + 1) * 8, np.uint64)
    bases[:] = reads.reshape(-1)
    offs[:] = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    del reads
    batch = rb.binding._Batch(n_reads, bases.ctypes.data, offs.ctypes.data)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident leg: `value` ----------------------------------------------------
    staged = ix.upload((bases, offs))
    for _ in range(args.warmup):
        ix.query_staged(staged, mode)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms, search_ms, launches, steps_lf, lines_lf, phi_steps, mk_words = [], [], 0, 0, 0, 0, 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        ix.query_staged(staged, mode)
        st = ix.stats()
        dev_ms.append(st.ms_total)
        search_ms.append(st.ms_search)
        launches += st.launches
        steps_lf, lines_lf, phi_steps, mk_words = st.lf_steps, st.lf_lines, st.phi_steps, st.marker_words
        last = st.as_dict()
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    checksum = ix.query_staged(staged, mode, checksum=True)

    # ---- end-to-end leg through rbg_query with host buffers: `e2e` ----------------------------
    for _ in range(min(args.warmup, 2)):
        ix.query_raw(batch, mode)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ix.query_raw(batch, mode)
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / args.steps
    e2e_stats = ix.stats().as_dict()
    clocks = sampler.summary()
    staged.free()

    t = torch.tensor([float(np.mean(dev_ms)), wall_ms, e2e_ms, float(np.mean(search_ms))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, wall_step, e2e_step, search_step = [float(x) for x in t.tolist()]
    total_reads = n_reads * world

    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        # roofline of the dominant kernel (search_kernel): algorithmic bytes per launch:
        #   64 B per distinct directory line an LF step loads + 38 B of 2-bit read + 16 B of range per read
        alg_bytes = lines_lf * 64 + n_reads * (38 + 16)
        achieved = alg_bytes / (float(np.mean(search_ms)) * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "lf_steps_per_s": steps_lf / (float(np.mean(search_ms)) * 1e-3),
                "lines_per_lf_step": lines_lf / max(1, steps_lf)}
        # dram bytes of one launch from the committed ncu capture of this workload (profiles/), if any
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            t_ = json.load(open(tpath)).get("%s:%s:%d" % (cfg, args.mode, n_reads))
            if t_:
                roof["traffic"] = t_["dram_bytes"]
                roof["traffic_source"] = t_["source"]
        if not args.no_gather:
            # the ceiling that actually binds this kernel: random 64-byte line reads, measured
            # live over a buffer as large as the rank directory and over 8 GB (beyond the TLB reach)
            g_dir = lib.rbg_gather_roofline(local, max(int(info.dir_bytes), 1 << 20), 64, 256)
            g_big = lib.rbg_gather_roofline(local, 8 << 30, 64, 256)
            line_gbs = lines_lf * 64 / (float(np.mean(search_ms)) * 1e-3) / 1e9
            roof["random_gather_64B_gbs_at_dir_footprint"] = g_dir
            roof["random_gather_64B_gbs_8GB"] = g_big
            roof["line_gbs"] = line_gbs
            roof["frac_of_random_gather"] = line_gbs / g_dir if g_dir > 0 else None
        out = {"metric": "150bp reads/s (%s)" % args.mode, "value": total_reads / (ms_step * 1e-3), "unit": "reads/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
               "wall_ms_per_step": wall_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u64", "data": "synthetic",
               "config": {"workload": describe(cfg, n_reads), "mode": args.mode, "reads_per_gpu": n_reads, "read_len": READ_LEN,
                          "index": {"n": info.n, "r": info.r, "window": info.window, "lines": info.n_lines,
                                    "cluster_windows": info.n_cluster, "dir_MB": info.dir_bytes / 1e6,
                                    "phi_MB": info.phi_bytes / 1e6, "toehold_MB": info.toehold_bytes / 1e6,
                                    "ftab_k": info.ftab_k, "ftab_MB": info.ftab_bytes / 1e6,
                                    "l2_window_MB": info.hot_bytes / 1e6, "l2_persisting_MB": info.l2_pinned_bytes / 1e6},
                          "l2": "inputs larger than L2 (%.0f MB index + %.0f MB reads per step)" % (
                              info.dir_bytes / 1e6, n_reads * READ_LEN / 1e6),
                          "parallelism": "replicated index, reads sharded, no collective"},
               "lf_steps_per_s": steps_lf * world / (ms_step * 1e-3), "lf_steps_per_step": steps_lf,
               "phi_steps_per_step": phi_steps, "marker_words_per_step": mk_words,
               "kernel_ms": {k: last[k] for k in ("ms_pack", "ms_search", "ms_locate", "ms_markers", "ms_total")},
               "checksum": checksum, "gpu_launches": launches, "clocks": clocks, "roofline": roof,
               "e2e": {"value": total_reads / (e2e_step * 1e-3), "unit": "reads/s", "ms_per_step": e2e_step,
                       "h2d_bytes_per_step": n_reads * READ_LEN + (n_reads + 1) * 8,
                       "d2h_bytes_per_step": 16 * n_reads + (8 * n_reads + 8 * (n_reads + 1) + 8 * phi_steps + 8 * n_reads if mode & 1 else 0)
                       + (8 * (n_reads + 1) + 8 * mk_words if mode & 2 else 0),
                       "stages_ms": {k: e2e_stats[k] for k in ("ms_h2d", "ms_pack", "ms_search", "ms_locate", "ms_markers", "ms_d2h", "ms_total")}},
               "setup_s": {"index_open": t_open, "ftab_build": t_ftab, "make_reads": t_reads}}
        # CPU baseline beside it: the unmodified reference on a bounded sample, one process (as shipped)
        if not args.no_cpu_baseline and world == 1 and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rb_align")):
            k = min(args.cpu_sample, n_reads)
            sample_reads = np.asarray(bases[:k * READ_LEN]).reshape(k, READ_LEN)
            with tempfile.TemporaryDirectory() as td:
                v, qt = run_reference_shards(prefix, sample_reads, mode, 1, td)
            out["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": 1, "kind": "reference",
                                   "sample": "first %d reads of the step's batch through oracle/_ref/rb_align (its own total_query_time %.2f s)" % (k, qt)}
        elif world == 1:
            out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not present"}
        print(json.dumps(out))
    ix.close()
    lib.rbg_host_free(_p1)
    lib.rbg_host_free(_p2)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
