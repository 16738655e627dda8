#!/usr/bin/env python
"""bench.py -- wavesolve FEM hot path on B200: assembly (+ eigensolve when built).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic problem per GPU:
workload c4 = vector (Whitney edge + P1 node) problem on a ~1M-triangle step-index-fibre
mesh (BASELINE.json configs[3], the configuration the metric is quoted on).  At N>1 every
rank runs its own problem of the same size (a sweep point: one problem per GPU, no data-path
collective) -> weak scaling; value = work of all ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WL = 1.55


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-elements", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def make_mesh(n_elements, seed=0):
    from wavesolve_b200 import meshgen as mg
    return mg.fiber_disc(n_elements, order=1, seed=seed), mg.FIBER_IOR


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_port_run(n_elements, steps=1, warmup=0):
    """Times the numpy/scipy restatement of the reference path (oracle/restated.py: vectorised
    get_unique_edges + construct_AB_vec; the literal reference is O(E^2)/dense and cannot run
    these sizes) on a bounded sample."""
    from oracle import restated as R
    mesh, ior = make_mesh(n_elements, seed=1)
    k = 2 * np.pi / WL
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        R.get_unique_edges(mesh)
        A, B = R.construct_AB_vec(mesh, ior, k)
        dt = time.perf_counter() - t0
        if it >= warmup:
            ts.append(dt)
    t = float(np.mean(ts))
    return {"value": mesh.num_elements / t, "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": "oracle/restated.py get_unique_edges + construct_AB_vec on a %d-triangle fibre mesh "
                      "(numpy/scipy, single-threaded); %.2f s per pass" % (mesh.num_elements, t),
            "seconds": t, "elements": mesh.num_elements}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_run(args.cpu_sample_elements, steps=max(1, args.steps), warmup=min(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements": r["elements"],
                   "note": "bounded CPU sample of the workload"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


METRIC = "FEM elements/s assembled (vector Whitney+P1: edge table + CSR pattern + A,B values)"
WORKLOAD = "c4: vector problem, step-index fibre, ~1M triangles (BASELINE.json configs[3])"


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- wavesolve_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import wavesolve_b200 as ws
    from wavesolve_b200 import _lib, fe_solver as fs, waveguide as wg, device as dv

    mesh, ior = make_mesh(args.elements, seed=rank)
    k = 2 * np.pi / WL
    ne = mesh.num_elements
    tris_h = np.ascontiguousarray(np.asarray(mesh.cells[1].data)[:, :3].astype(np.int32))
    npts = mesh.points.shape[0]

    # ---- device-resident step: inputs (mesh arrays) already in HBM -------------------------
    tris_d = dv.h2d(tris_h, dev)
    # (the mesh object needs the edge table for the material-major marshalling; get it once)
    wg.get_unique_edges(mesh)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    asm_ms = []

    def device_step(prob, record):
        # K1 edge table on device
        edges_d, eidx_d = wg.unique_edges_device(tris_d, npts)
        # K2 pattern + K4 assembly
        pat = fs.DevicePattern(prob.dofs, prob.N)
        old = prob.pattern
        prob.pattern = pat
        old.close()
        if record:
            e0, e1 = ev(), ev()
            e0.record()
        A, B = prob.assemble()
        if record:
            e1.record()
            asm_ms.append((e0, e1))
        return A, B

    prob = fs.VectorProblem(mesh, ior, k, device=dev)
    nnz = prob.pattern.nnz
    N = prob.N

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step(prob, False)
    barrier()
    l0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        t0e, t1e = ev(), ev()
        t0e.record()
        for _ in range(args.steps):
            device_step(prob, True)
        t1e.record()
        barrier()
        dev_ms = t0e.elapsed_time(t1e) / args.steps
    launches = _lib.launch_count() - l0
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in asm_ms]))

    # ---- end to end through the public API with host buffers -------------------------------
    import copy
    e2e_t = []
    h2d_b = d2h_b = 0
    for it in range(max(1, min(args.steps, 3)) + 1):
        m = copy.copy(mesh)
        m.cells = [copy.copy(c) for c in mesh.cells]
        barrier()
        t0 = time.perf_counter()
        edges, eidx = ws.get_unique_edges(m)
        A, B = ws.construct_AB_vec(m, ior, k, sparse=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it > 0:
            e2e_t.append(dt)
        h2d_b = tris_h.nbytes + (mesh.points[:, :2].nbytes + tris_h.nbytes * 3 + 8 * ne)
        d2h_b = edges.nbytes // 2 + eidx.nbytes // 2 + A.data.nbytes + A.indices.nbytes + A.indptr.nbytes \
            + B.data.nbytes + B.indices.nbytes + B.indptr.nbytes
    e2e_s = float(np.mean(e2e_t))

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, k_ms = [float(x) for x in t.tolist()]
        tot = torch.tensor([float(ne)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_elems = float(tot.item())
    else:
        total_elems = float(ne)

    peak, peak_src = peaks()
    # algorithmic bytes of the vector assembly kernel: connectivity 12 B +
    # k^2n^2 8 B per element, vertex coordinates 16 B per vertex, 8 B per stored value of A and B
    algo_bytes = ne * (12 + 8) + 16 * npts + 8 * 2 * nnz
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": total_elems / (dev_ms * 1e-3), "unit": "elements/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": ne, "N": N, "nnz": nnz,
                   "l2": "inputs_larger_than_l2 (%.0f MB touched per step)" % ((algo_bytes + 96 * ne) / 1e6),
                   "parallelism": "one problem per GPU"},
        "roofline": {"bound": "hbm", "kernel": "k_assemble_rows<VectorP1>", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": algo_bytes, "kernel_ms": k_ms, "traffic": None},
        "e2e": {"value": total_elems / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d_b),
                "d2h_bytes_per_step": int(d2h_b), "seconds": e2e_s,
                "api": "get_unique_edges + construct_AB_vec(sparse=True) with host numpy in/out"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            r = cpu_port_run(args.cpu_sample_elements)
            line["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
