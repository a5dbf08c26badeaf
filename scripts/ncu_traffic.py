#!/usr/bin/env python
"""profiles/<tag>_ncu_full_raw.csv (ncu --set full --page raw --csv) -> profiles/ncu_traffic.json: per kernel and launch the DRAM
bytes (dram__bytes_read.sum + dram__bytes_write.sum), the duration and the pipe counters bench.py quotes next to its live
timings (`roofline.traffic`).  Usage: python scripts/ncu_traffic.py profiles/r2c_ncu_full_raw.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WANT = {"dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write", "gpu__time_duration.sum": "duration",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_active_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct"}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"source": os.path.relpath(path, ROOT), "kernels": {}}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "greedy_ws_kernel" if "greedy_ws" in name else "fe_fused_kernel" if "fe_fused" in name else name
        k = {}
        for metric, short in WANT.items():
            if metric not in col:
                continue
            v, u = float(r[col[metric]].replace(",", "")), units[col[metric]]
            if short.startswith("dram"):
                k[short] = v * SCALE[u]
            elif short == "duration":
                k["duration_ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
            else:
                k[short] = v
        k["dram_bytes"] = k.get("dram_bytes_read", 0.0) + k.get("dram_bytes_write", 0.0)
        out["kernels"][key] = k
    dst = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1])
