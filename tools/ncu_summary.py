#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel, --set full) into the handful of numbers the roofline and the
optimisation log need.  Runs here (no GPU): ncu -i <rep> --page raw --csv.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [--md profiles/x.md] [--title "..."]
"""
import argparse
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by L1TEX"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps/issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving (warps/issue)"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch (warps/issue)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction (warps/issue)"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--md")
    ap.add_argument("--title", default="")
    ap.add_argument("--grep", default="", help="also print every metric whose name contains this")
    a = ap.parse_args()
    kernels, units = load(a.rep)
    lines = []
    for k in kernels:
        lines.append("### %s  `%s`" % (a.title or a.rep, k.get("Kernel Name", "?")))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for name, label in WANT:
            if name in k:
                lines.append("| %s (`%s`) | %s | %s |" % (label, name, k[name], units.get(name, "")))
        if a.grep:
            for name in sorted(k):
                if a.grep in name:
                    lines.append("| `%s` | %s | %s |" % (name, k[name], units.get(name, "")))
        lines.append("")
    text = "\n".join(lines)
    print(text)
    if a.md:
        with open(a.md, "a") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
