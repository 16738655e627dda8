#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name:
launch count, total / mean device time and share of the total."""
import csv
import re
import sys
from collections import defaultdict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def main(path, top=40):
    rows = load(path)
    tot = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t, _, _ in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print("launches %d, total device time %.3f ms" % (len(rows), tot / 1e3))
    print("%-70s %7s %12s %10s %7s" % ("kernel", "count", "total_us", "mean_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-70s %7d %12.1f %10.2f %6.1f%%" % (n[:70], c, t, t / c, 100 * t / tot))
    return rows


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
