#!/usr/bin/env python
"""Per-kernel SASS instruction mix of libwavesolve_b200.so (cuobjdump -sass; runs without a GPU):
fp64 tensor-core MMAs (DMMA), fp64 FMA/ADD/MUL, global loads / stores by width, asynchronous copies
(LDGSTS = cp.async, UBLKCP / UTMALDG = bulk / TMA), atomics, barriers.  Written to profiles/ as the evidence of
which hardware paths the kernels use."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wavesolve_b200", "libwavesolve_b200.so")
COLS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU.RCP64H", "LDG.E.64", "LDG.E.128", "LDG(other)", "STG", "LDGSTS", "UBLKCP", "UTMALDG",
        "LDS", "STS", "ATOM/RED", "BAR", "SHFL", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    fn = None
    table = collections.OrderedDict()
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            table[fn] = collections.Counter()
            continue
        m = ins.match(line)
        if not m or fn is None:
            continue
        op = m.group(1)
        c = table[fn]
        c["total"] += 1
        if op.startswith("DMMA"): c["DMMA"] += 1
        elif op.startswith("DFMA"): c["DFMA"] += 1
        elif op.startswith("DADD"): c["DADD"] += 1
        elif op.startswith("DMUL"): c["DMUL"] += 1
        elif op.startswith("MUFU.RCP64H"): c["MUFU.RCP64H"] += 1
        elif op.startswith("LDGSTS"): c["LDGSTS"] += 1
        elif op.startswith("UBLKCP"): c["UBLKCP"] += 1
        elif op.startswith("UTMALDG"): c["UTMALDG"] += 1
        elif op.startswith("LDG"):
            if ".128" in op: c["LDG.E.128"] += 1
            elif ".64" in op: c["LDG.E.64"] += 1
            else: c["LDG(other)"] += 1
        elif op.startswith("STG"): c["STG"] += 1
        elif op.startswith("LDS"): c["LDS"] += 1
        elif op.startswith("STS"): c["STS"] += 1
        elif op.startswith(("ATOM", "RED", "ATOMG")): c["ATOM/RED"] += 1
        elif op.startswith("BAR"): c["BAR"] += 1
        elif op.startswith("SHFL"): c["SHFL"] += 1
    names = list(table)
    try:
        dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        demangle = {n: n for n in names}
    print("# SASS instruction mix per kernel of wavesolve_b200/libwavesolve_b200.so (sm_100a), tools/sass_summary.py")
    print("# DMMA = fp64 tensor-core MMA (mma.sync.m8n8k4.f64; tcgen05 has no fp64 kind); LDGSTS = cp.async; UBLKCP/UTMALDG = bulk/TMA copies")
    print("%-72s " % "kernel" + " ".join("%11s" % c for c in COLS))
    ours = [n for n in names if "ws" in demangle[n] or demangle[n].startswith(("k_", "void k_"))]
    for n in sorted(ours, key=lambda x: demangle[x]):
        d = re.sub(r"\(.*", "", demangle[n]).replace("void ", "")
        if "cub::" in d or "thrust::" in d:
            continue
        print("%-72s " % d[:72] + " ".join("%11d" % table[n][c] for c in COLS))
    tot = collections.Counter()
    for n in ours:
        tot.update(table[n])
    print("%-72s " % "TOTAL (own kernels)" + " ".join("%11d" % tot[c] for c in COLS))


if __name__ == "__main__":
    main()
