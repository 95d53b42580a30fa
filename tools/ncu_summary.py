#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit, val = r.get("Metric Unit", "ns"), float(r["Metric Value"].replace(",", ""))
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary ({path})\n")
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms summed kernel time (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / c / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
