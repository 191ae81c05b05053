#!/usr/bin/env python3
"""One line per leg of a bench.py JSON line (for the run logs)."""
import json
import sys

d = json.load(open(sys.argv[1]))


def show(name, r):
    rl = r.get("roofline_locate")
    print("%-12s dev %.2f ms (2-bit in, narrow out %.2f)  search %.2f  phi %.2f  roof %.3f%s  e2e %.2f ms%s" % (
        name, r["ms_per_step"], r.get("ms_per_step_packed_input", 0), r["kernel_ms"]["ms_search"], r["kernel_ms"].get("ms_phi", 0),
        r["roofline"]["frac"], "  locate-roof %.3f (%.1f G phi/s)" % (rl["frac"], rl["phi_steps_per_s"] / 1e9) if rl else "",
        r["e2e"]["ms_per_step"], "  ascii %.2f" % r["e2e_ascii"]["ms_per_step"] if "e2e_ascii" in r else ""))


show(d["config"]["mode"], dict(d, ms_per_step_packed_input=1e3 * d["config"]["reads_per_gpu"] * d["n_gpus"] / d["value_packed_input"]))
for k, v in d["legs"].items():
    show(k, v)
print("n_gpus", d["n_gpus"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "cpu", (d.get("cpu_baseline") or {}).get("value"),
      "clocks", d["clocks"], "host_pack GB/s %.1f" % d["host_pack"]["gb_per_s"], "layout", d["config"]["index"].get("layout"), "launches", d["gpu_launches"])
