#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and average
duration and share per kernel.  usage: ncu_summary.py launches.csv > summary.csv"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        if "k_block_op" in name or "k_packed" in name:  # the same kernel runs on every multigrid level: split fine / coarse by grid
            # ("fine" = the grid is at its cap of 148 x blocks per SM: levels 0 and 1 at n = 215)
            cap = 148 * (3 if "k_packed" in name else 6)
            name += " fine" if int(r["Grid Size"].strip("()").split(",")[0]) >= cap else " coarse"
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "sum_us", "avg_us", "share_pct"])
    for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, round(s, 1), round(s / n, 2), round(100 * s / tot, 1)])


if __name__ == "__main__":
    main(sys.argv[1])
