#!/usr/bin/env python
"""Print the per-launch table (time, DRAM bytes read, DRAM %) of the last solve captured by
tools/ncu_solve_list.sh."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/solve_list.csv"
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 33
lines = [l for l in open(path) if not l.startswith("==")]
L = OrderedDict()
for r in csv.DictReader(lines):
    d = L.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0], "grid": r["Grid Size"], "block": r["Block Size"]})
    d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
tot = tot_b = 0.0
for k in list(L)[-n_last:]:
    d = L[k]
    t, b, pc = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    tu = t[0] / 1000 if t[1] == "ns" else t[0]
    mb = b[0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[b[1]]
    tot += tu
    tot_b += mb
    print("%-30s grid %-14s block %-12s %8.1f us  read %7.1f MB  dram %5.1f%%  (%.0f GB/s)" %
          (d["name"][-30:], d["grid"], d["block"], tu, mb, pc[0], mb / tu * 1e3 if tu else 0))
print("total %.1f us, %.1f MB" % (tot, tot_b))
