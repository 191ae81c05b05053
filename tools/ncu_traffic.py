#!/usr/bin/env python3
"""profiles/ncu_traffic.json from full ncu captures: DRAM bytes (read + written) of ONE launch of a kernel, keyed
"<config>:<leg>:<reads>:<kernel>" and stamped with the hash of the kernel sources they were captured at (bench.py drops an
entry whose hash differs from the current sources instead of reporting a stale figure).

  python tools/ncu_traffic.py c2:count:10000000:search_kernel=gpurun_out/x.ncu-rep [more key=rep ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ncu_summary import load  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    import bench
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out = {k: v for k, v in out.items() if k.count(":") == 3}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        kernels, units = load(rep)
        k = kernels[0]
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(k[m].replace(",", "")) * UNIT[units[m]]
        out[key] = {"dram_bytes": int(total), "kernel_name": k.get("Kernel Name"), "src_hash": bench.kernel_source_hash(),
                    "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum, one launch, ncu --set full)" % os.path.basename(rep)}
        print(key, out[key])
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
