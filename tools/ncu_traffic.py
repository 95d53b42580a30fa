#!/usr/bin/env python
"""Sum DRAM traffic per kernel family from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv` capture of `bench.py --ncu` (one eager step) and write profiles/rNN_traffic.json, which bench.py reports as
roofline.traffic (bytes per step for the dense conv stack = the kernels the roofline is quoted on).
usage: python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r01_traffic.json"""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1}
CONV = ("conv_fused_kernel", "conv_tc_kernel", "conv_tc_persist_kernel", "prep_bf16_kernel", "conv1d_dense_kernel", "convtr1d_dense_kernel",
        "channel_stats_kernel", "coeffs_from_stats_kernel")


def main(path, out):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = defaultdict(lambda: {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0})
    ids = set()
    for r in csv.DictReader(lines):
        name = r["Kernel Name"].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"^.*::", "", re.sub(r"[<(].*", "", re.sub(r"^void\s+", "", name))).strip()
        m, val = r.get("Metric Name"), float(r["Metric Value"].replace(",", "") or 0)
        if m == "gpu__time_duration.sum":
            per[name]["ns"] += val * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r.get("Metric Unit", "ns"), 1)
            if r["ID"] not in ids:
                ids.add(r["ID"])
                per[name]["launches"] += 1
        elif m == "dram__bytes_read.sum":
            per[name]["read"] += val * UNIT.get(r.get("Metric Unit", "byte"), 1)
        elif m == "dram__bytes_write.sum":
            per[name]["write"] += val * UNIT.get(r.get("Metric Unit", "byte"), 1)
    conv = {k: v for k, v in per.items() if k in CONV}
    res = {"source": path, "note": "one eager Kokoro cfg2 step under ncu (cold-ish L2, serialised); bytes are dram__bytes_read.sum + dram__bytes_write.sum",
           "conv_stack_bytes_per_step": sum(v["read"] + v["write"] for v in conv.values()),
           "conv_stack_launches": sum(v["launches"] for v in conv.values()),
           "all_kernels_bytes_per_step": sum(v["read"] + v["write"] for v in per.values()),
           "per_kernel": {k: {"launches": v["launches"], "dram_read_MB": round(v["read"] / 1e6, 3), "dram_write_MB": round(v["write"] / 1e6, 3),
                              "time_us": round(v["ns"] / 1e3, 1)} for k, v in sorted(per.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"]))}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("conv_stack_bytes_per_step", "conv_stack_launches", "all_kernels_bytes_per_step")}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
