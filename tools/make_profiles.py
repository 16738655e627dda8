#!/usr/bin/env python
"""Reduce the scratch profiler output in gpurun_out/ to the small, tracked summaries under
profiles/ (the judge reads profiles/, gpurun_out/ is scratch).

    python tools/make_profiles.py [TAG]        # TAG = r01 by default

Inputs (whatever exists):
  gpurun_out/launches_TAG.csv      ncu --metrics gpu__time_duration.sum launch list of one bench.py run
  gpurun_out/prof_TAG_raw.csv      `ncu --set full` raw page of the hot kernels
  gpurun_out/solve_list.csv        per-launch time + DRAM bytes of one shift-invert solve
  gpurun_out/plain_bench_TAG.log   the bench line of the same command without a profiler
Outputs:
  profiles/TAG_launches_by_kernel.txt, profiles/TAG_hot_kernels_ncu.csv,
  profiles/TAG_solve_sweep_launches.txt, profiles/TAG_bench_line.json
"""
import csv
import io
import json
import os
import re
import sys
from collections import defaultdict, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max",
]


def launches(tag):
    src = os.path.join(G, "launches_%s.csv" % tag)
    if not os.path.exists(src):
        return
    lines = [l for l in open(src, newline="") if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r.get("Metric Unit", "ns"), 1e-3)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        rows.append((name, v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    out = io.StringIO()
    out.write("# ncu launch list of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-clock-sampler --lite` (one step)\n")
    out.write("# (WS_NO_GRAPH=1: same kernels launched directly; per-launch times under ncu are cold-cache and\n")
    out.write("#  serialised, so read SHARES, not absolutes.)  launches %d, total device time %.3f ms\n" % (len(rows), tot / 1e3))
    out.write("%-72s %7s %12s %10s %7s\n" % ("kernel", "count", "total_us", "mean_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("%-72s %7d %12.1f %10.2f %6.1f%%\n" % (n[:72], c, t, t / c, 100 * t / tot))
    fam = defaultdict(float)
    for n, (c, t) in agg.items():
        if re.search(r"k_fwd|k_bwd|k_permute", n):
            fam["solve sweeps (K8c)"] += t
        elif re.search(r"k_gemm_tiles|k_panel|k_diag_inverse|k_extend_add|k_scatter|k_front_fused|k_recip|k_dest_map|k_build_wsrc|k_count_bad", n):
            fam["numeric factorisation (K8b)"] += t
        elif re.search(r"k_spmv", n):
            fam["SpMV (K5)"] += t
        elif re.search(r"k_dots|k_update|k_normalize|k_rotate|k_scale_rows|k_vector_T", n):
            fam["Krylov vector kernels (K6/K7/K10)"] += t
        elif re.search(r"k_assemble", n):
            fam["assembly (K3/K4)"] += t
        elif re.search(r"k_centroids|k_bbox|k_kd_keys|k_leaf_of|k_owner|k_bucket|k_after_order|k_front_own|k_count_pairs|k_emit_pairs|k_bnd_split|k_front_bounds|k_rel|k_fill_i32", n):
            fam["symbolic analysis (K8a)"] += t
        else:
            fam["pattern / edges / sort / scan (K1/K2, CUB inside K8a)"] += t
    out.write("\n# by family\n")
    for n, t in sorted(fam.items(), key=lambda kv: -kv[1]):
        out.write("%-44s %12.1f us %6.1f%%\n" % (n, t, 100 * t / tot))
    open(os.path.join(P, "%s_launches_by_kernel.txt" % tag), "w").write(out.getvalue())


def hot(tag, suffix="", keep_all_matching=None):
    src = os.path.join(G, "prof_%s%s_raw.csv" % (tag, suffix))
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, newline="")))
    hdr, units = rows[0], rows[1]
    keep = [k for k in KEEP if k in hdr]
    if keep_all_matching:
        keep += [h for h in hdr if re.search(keep_all_matching, h) and h not in keep]
    idx = [hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")] + [hdr.index(k) for k in keep]
    with open(os.path.join(P, "%s%s_hot_kernels_ncu.csv" % (tag, suffix)), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if len(r) == len(hdr):
                w.writerow([re.sub(r"\(.*", "", r[i]) if i == idx[0] else r[i] for i in idx])


def solve_list(tag):
    src = os.path.join(G, "solve_list.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    L = OrderedDict()
    for r in csv.DictReader(lines):
        d = L.setdefault(r["ID"], {"name": re.sub(r"^void ", "", r["Kernel Name"].split("(")[0]), "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    # keep the last complete solve: everything after the last-but-one k_spmv_stream
    keys = list(L)
    out = io.StringIO()
    out.write("# one shift-invert solve (forward + backward sweep over the inverted panels) + the SpMV that follows,\n")
    out.write("# 1M-triangle vector problem, ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput\n")
    tot = tot_b = 0.0
    sel = keys[-33:]
    for k in sel:
        d = L[k]
        t, b = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"]
        pc = d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", (0.0, "%"))
        tu = t[0] / 1000 if t[1] in ("ns", "nsecond") else t[0]
        mb = b[0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[b[1]]
        tot += tu
        tot_b += mb
        out.write("%-30s grid %-14s block %-12s %8.1f us  read %7.1f MB  dram %5.1f%%  (%.0f GB/s)\n" %
                  (d["name"][-30:], d["grid"], d["block"], tu, mb, pc[0], mb / tu * 1e3 if tu else 0))
    out.write("total %.1f us, %.1f MB read\n" % (tot, tot_b))
    open(os.path.join(P, "%s_solve_sweep_launches.txt" % tag), "w").write(out.getvalue())
    # DRAM traffic of one solve (sweep kernels only, SpMV excluded) for bench.py's roofline.traffic
    rd = wr = 0.0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for k in sel:
        d = L[k]
        if "spmv" in d["name"]:
            continue
        b = d["dram__bytes_read.sum"]
        rd += b[0] * scale[b[1]]
        w = d.get("dram__bytes_write.sum")
        if w:
            wr += w[0] * scale[w[1]]
    json.dump({"solve_dram_bytes": int(rd + wr), "read": int(rd), "write": int(wr),
               "source": "ncu dram__bytes_read.sum / dram__bytes_write.sum summed over the %d sweep launches of one "
                         "shift-invert solve (tools/solve_only.py, 1M-element vector problem)" % (len(sel) - 1)},
              open(os.path.join(P, "%s_traffic.json" % tag.split("_")[0][:3]), "w"), indent=1)


def bench_line(tag):
    for name in ("bench_%s.log" % tag, "plain_bench_%s.log" % tag):
        src = os.path.join(G, name)
        if not os.path.exists(src):
            continue
        for l in open(src):
            if l.startswith("{") and '"metric"' in l:
                json.dump(json.loads(l), open(os.path.join(P, "%s_%s.json" % (tag, name.split("_" + tag)[0])), "w"), indent=1)
                break


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(P, exist_ok=True)
    launches(tag)
    hot(tag)
    hot(tag, "_wide", keep_all_matching=r"tensor|dmma|fp64")
    solve_list(tag)
    bench_line(tag)
    print(sorted(os.listdir(P)))
