#!/usr/bin/env python
"""bench.py -- wavesolve FEM hot path on B200: assembly + generalized sparse eigensolve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload c4 (BASELINE.json configs[3], the configuration the metric is quoted on): vector
modes (Whitney edge + P1 node elements) of a step-index fibre on a ~1M-triangle mesh, 10 modes.
One "step" = one complete eigensolve from the mesh arrays: edge table (K1), CSR pattern (K2),
assembly of A/B and of the quasi-definite shift-invert matrix (K4), nested-dissection analysis,
multifrontal LDL^T (K8), Krylov iteration (K5-K10).  `value` = eigensolves/s with the mesh
arrays resident in HBM; `e2e` = the same through the public API (get_unique_edges +
solve_waveguide_vec) with host numpy in/out.  At N>1 every rank solves its own problem of the
same size (sweep point: one problem per GPU, no data-path collective) -> weak scaling.

Prints ONE JSON line (rank 0).  DESIGN.md "Measurement" states the byte/flop accounting.
"""
from __future__ import annotations

import argparse
import copy
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
NMAX = 10
METRIC = "eigensolves/s (10 vector modes, ~1M-triangle mesh, assembly + factorisation + Krylov iteration)"
WORKLOAD = "c4: solve_waveguide_vec, step-index fibre r_core=5 r_clad=10, ~1M P1 triangles, Nmax=10 (BASELINE.json configs[3])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-elements", type=int, default=80_000)
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("WS_REF_BUDGET_S", "1500")),
                    help="--impl reference at N=1: the same-configuration CPU solve (the full --elements mesh, one timed "
                         "step) runs when a calibration solve predicts it fits this many seconds; otherwise the largest "
                         "size that does")
    ap.add_argument("--cal-elements", type=int, default=0, help="size of the calibration solve (default 40k triangles)")
    ap.add_argument("--sweep-problems", type=int, default=256, help="problems of the timed c5 sweep slice (0: skip)")
    ap.add_argument("--sweep-elements", type=int, default=50_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lite", action="store_true",
                    help="time the steps only: skip the kernel micro-measurements and the end-to-end pass "
                         "(the command the ncu launch list under profiles/ is taken from)")
    ap.add_argument("--no-clock-sampler", action="store_true",
                    help="do not spawn nvidia-smi during the timed region (profiler runs)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: ONE `nvidia-smi -lms 200`
    process (B200_PROFILING.md's clocks line) started before and terminated after -- spawning a
    fresh nvidia-smi per sample re-initialises NVML every time and stalls the driver calls of the
    step being measured."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.th = None

    def _run(self):
        try:
            for line in self.proc.stdout:
                line = line.strip()
                if line:
                    self.samples.append([s.strip() for s in line.split(",")])
        except Exception:
            pass

    def __enter__(self):
        if self.index is not None:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.th = threading.Thread(target=self._run, daemon=True)
                self.th.start()
                # NVML initialisation of the sampler must be over before the timed region starts (it stalls
                # the driver calls of the step under test): wait for its first sample, not for a fixed time --
                # on a fresh box nvidia-smi and libnvidia-ml are still being paged in
                t0 = time.time()
                while not self.samples and time.time() - t0 < 10.0 and self.proc.poll() is None:
                    time.sleep(0.02)
                time.sleep(0.1)
            except Exception:
                self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                pass
            if self.th is not None:
                self.th.join(timeout=5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_traffic():
    """DRAM bytes of one shift-invert solve from the committed ncu capture (None if absent)."""
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["solve_dram_bytes"])
    except Exception:
        return None


def make_mesh(n_elements, seed=0):
    from wavesolve_b200 import meshgen as mg
    return mg.fiber_disc(n_elements, order=1, seed=seed), mg.FIBER_IOR


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def host_info():
    """Core count and thread environment of the host the CPU arm runs on."""
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = None
    return {"cpu_count": os.cpu_count(), "sched_affinity": aff,
            "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS"), "OPENBLAS_NUM_THREADS": os.environ.get("OPENBLAS_NUM_THREADS"),
            "MKL_NUM_THREADS": os.environ.get("MKL_NUM_THREADS"),
            "threads_used": 1, "why": "scipy's SuperLU (splu) and ARPACK (eigs) are single-threaded; the reference's path has no other parallelism"}


# Same-configuration CPU measurement made once in the build container (oracle/make_golden_at_size.py c4, kept in
# tests/golden/c4_vec_1M.npz): 1545 s for ONE c4 eigensolve with oracle/restated.py; a 40k-triangle solve of the same
# fibre took 8.9 s on the same host the same hour.  The ratio predicts the full-size time on another host from a
# calibration solve there.
CAL_ELEMENTS = 40_000
CAL_SECONDS_BUILD_HOST = 8.9
SIZE_EXPONENT = 1.83          # t ~ Ne^1.83 between 80k (24.0 s) and 1M (1545 s) triangles on the build host


def same_config_record():
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "c4_vec_1M.npz"))
        return {"seconds": float(z["seconds"]), "elements": int(z["elements"]), "n_op": int(z["n_op"]),
                "eigensolves_per_s": 1.0 / float(z["seconds"]), "cores": 1,
                "where": "build container (8 cores), oracle/make_golden_at_size.py c4 -> tests/golden/c4_vec_1M.npz; "
                         "the eigenvalues of that run are what tests/test_gpu_at_size.py asserts"}
    except Exception:
        return None


def cpu_port_run(n_elements, steps=1, warmup=0, budget_s=None):
    """The reference's algorithm for this path restated in numpy/scipy (oracle/restated.py:
    vectorised get_unique_edges + construct_AB_vec + matrix-free 'transform' eigs with SuperLU),
    timed on the host.  The literal reference cannot run the vector path beyond ~5k elements
    (O(E^2) edge search, dense N x N matrices), and the 1M-element workload would
    take well over ten minutes per solve on the CPU, so this is a BOUNDED sample: the same
    geometry at `n_elements` triangles."""
    from oracle import restated as R
    mesh, ior = make_mesh(n_elements, seed=0)        # the mesh family (and, at full size, the very mesh) of our arm's rank 0
    ts, info = [], None
    for it in range(warmup + steps):
        m = copy.copy(mesh)
        m.cells = [copy.copy(c) for c in mesh.cells]
        t0 = time.perf_counter()
        R.get_unique_edges(m)
        w, v, cnt, info = R.solve_waveguide_vec(m, WL, ior, Nmax=NMAX, return_info=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            ts.append(dt)
            if budget_s is not None and sum(ts) + dt > budget_s:
                break                     # the next step would overrun the time budget of the reference arm
    t = float(np.mean(ts))
    return {"value": 1.0 / t, "unit": "eigensolves/s", "cores": 1, "kind": "port",
            "sample": "oracle/restated.py (numpy + scipy SuperLU/ARPACK, single-threaded like the reference's "
                      "scipy path) on the same fibre at %d triangles: %.2f s per solve, %d OP applications"
                      % (mesh.num_elements, t, info["n_op"]),
            "seconds": t, "elements": mesh.num_elements, "steps_timed": len(ts), "host": host_info()}


def run_reference(args):
    """--impl reference: the reference's algorithm for this path on the host cores (oracle/restated.py -- the
    literal reference cannot run the vector path beyond ~5k elements).

    N = 1: the SAME configuration as our arm -- the full --elements mesh, one timed solve, no warm-up -- when a
    40k-triangle calibration solve predicts that it fits --ref-budget-s (it took 1545 s in the build container);
    otherwise the largest mesh predicted to fit, with same_config=false and the extrapolation stated separately.
    N > 1 (the scaling runs): rank 0 runs a bounded sample sized for ~150 s -- the CPU path does not scale with
    the number of GPUs, the same-configuration number is the N = 1 line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_arm = time.perf_counter()
    cal = cpu_port_run(args.cal_elements or CAL_ELEMENTS)
    rec = same_config_record()
    full_build = rec["seconds"] if rec else 1545.0
    cal_build = CAL_SECONDS_BUILD_HOST * (cal["elements"] / 39910.0) ** SIZE_EXPONENT
    pred_full = full_build * cal["seconds"] / cal_build * (args.elements / 1.0e6) ** SIZE_EXPONENT
    budget = args.ref_budget_s if args.gpus == 1 else min(args.ref_budget_s, 150.0)
    if pred_full <= budget:
        n_el = args.elements
    else:
        n_el = int(args.elements * (budget / pred_full) ** (1.0 / SIZE_EXPONENT))
        n_el = max(min(n_el, args.elements), args.cpu_sample_elements)
    r = cpu_port_run(n_el, steps=1, warmup=0)
    same = abs(r["elements"] - args.elements) <= 0.02 * args.elements
    note = ("same configuration as the GPU arm: one timed solve of the full mesh, no warm-up (a CPU solve has no "
            "warm-up effects worth %.0f s)" % r["seconds"]) if same else \
        ("BOUNDED sample: %d of ~%d triangles (the full-size solve was predicted at %.0f s from a %.1f s calibration "
         "solve at %d triangles, budget %.0f s); not the same configuration -- see extrapolated_full_size"
         % (r["elements"], args.elements, pred_full, cal["seconds"], cal["elements"], budget))
    cfg = {"workload": WORKLOAD, "elements": r["elements"], "elements_per_gpu": r["elements"], "Nmax": NMAX,
           "same_config": bool(same), "note": note,
           "calibration": {"elements": cal["elements"], "seconds": cal["seconds"], "predicted_full_size_seconds": pred_full,
                           "build_host_seconds": CAL_SECONDS_BUILD_HOST, "size_exponent": SIZE_EXPONENT}}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "eigensolves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_timed": r["steps_timed"],
        "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "same_config": bool(same),
        "config": cfg,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host")},
        "e2e": {"value": r["value"], "unit": "eigensolves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "arm_seconds": None,
    }
    if not same:
        est = r["seconds"] * (args.elements / float(r["elements"])) ** SIZE_EXPONENT
        line["extrapolated_full_size"] = {"seconds": est, "eigensolves_per_s": 1.0 / est, "exponent": SIZE_EXPONENT,
                                          "basis": "t ~ Ne^%.2f fitted on the build host between 80k and 1M triangles; an "
                                                   "estimate, NOT a measurement" % SIZE_EXPONENT}
    if rec:
        line["same_config_build_host"] = rec
    line["arm_seconds"] = time.perf_counter() - t_arm
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# c5: the sweep slice (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------
def run_sweep_slice(args, dev, rank, world):
    """256 problems = 16 lantern z-slices (multimode-core radius 9 .. 16 as in lantern-demo.ipynb cell 4, one
    ~50k-triangle mesh each) x 16 wavelengths (1.0 .. 1.8), vector modes, Nmax = 10, partitioned over the ranks by
    sweep.partition, eigenpairs gathered on rank 0 over NCCL -- the gather and the device -> host copy of the
    gathered eigenvectors are INSIDE the timed region.  Mesh generation (host, stays on the host like gmsh in the
    reference) and the edge tables are outside it."""
    import torch
    import torch.distributed as dist
    import wavesolve_b200 as ws
    from wavesolve_b200 import meshgen as mg, sweep

    n_wl = 16
    n_mesh = max(1, args.sweep_problems // n_wl)
    radii = np.linspace(9.0, 16.0, n_mesh)
    meshes = [mg.lantern_cross_section(args.sweep_elements, order=1, seed=100 + i, r_clad_bundle=float(r))
              for i, r in enumerate(radii)]
    for m in meshes:
        ws.get_unique_edges(m)
    wls = np.linspace(1.0, 1.8, n_wl)
    probs = sweep.problem_grid(n_mesh, wls, mg.LANTERN_IOR, tags=[float(r) for r in radii])
    # warm-up: kernels, memory pools and the NCCL point-to-point connections of EVERY rank to rank 0 (set up lazily on
    # first use, ~0.5 s per peer: a one-off per process, like the communicator itself) -- one problem per rank
    sweep.solve_sweep(meshes[:1], probs[:max(2, world)], Nmax=NMAX, kind="vector", dst=0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    gs = {}
    t0 = time.perf_counter()
    res = sweep.solve_sweep(meshes, probs, Nmax=NMAX, kind="vector", dst=0, gather_stats=gs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, gs.get("seconds", 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    out = None
    if rank == 0:
        ok = all(r.rc == 0 for r in res)
        for i in (0, len(probs) // 2 + 3, len(probs) - 1):
            pr = probs[i]
            w, v, cnt = ws.solve_waveguide_vec(meshes[pr.mesh_id], pr.wl, pr.IOR_dict, Nmax=NMAX, verbose=False)
            k = 2 * np.pi / pr.wl
            ne_s, ne_i = np.sqrt(res[i].w) / k, np.sqrt(w.real) / k
            ok = ok and bool(np.max(np.abs(ne_s - ne_i) / ne_i) < 1e-10) and res[i].mode_count == cnt and res[i].v.shape == v.shape
        sizes = sorted(set(int(r.v.shape[1]) for r in res))
        out = {"workload": "c5: %d problems = %d lantern z-slices (~%dk P1 triangles each, N = %d..%d) x %d wavelengths, "
                           "solve_waveguide_vec semantics, Nmax=%d (BASELINE.json configs[4])"
                           % (len(probs), n_mesh, args.sweep_elements // 1000, sizes[0], sizes[-1], n_wl, NMAX),
               "problems": len(probs), "problems_per_s": len(probs) / dt, "seconds": dt,
               "gather_ms": gs.get("seconds", 0.0) * 1e3, "gathered_bytes": int(gs.get("bytes", 0)),
               "gather": "NCCL point-to-point to rank 0 + pinned device->host copy, inside the timed region"
                         if world > 1 else "single rank: pinned device->host copy of the packed results, inside the timed region",
               "ranks_that_solved": sorted(set(r.rank for r in res)),
               "failed_problems": int(sum(1 for r in res if r.rc != 0)),
               "parity_ok": bool(ok),
               "parity": "3 gathered problems re-solved through solve_waveguide_vec: neff within 1e-10, same mode count"}
    del res
    return out


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
    from wavesolve_b200 import _lib, fe_solver as fs, waveguide as wg, device as dv, solver as sv

    mesh, ior = make_mesh(args.elements, seed=rank)
    k = 2 * np.pi / WL
    sigma = float(np.power(k * max(ior.values()), 2))
    ne = mesh.num_elements
    tris_h = np.ascontiguousarray(np.asarray(mesh.cells[1].data)[:, :3].astype(np.int32))
    npts = mesh.points.shape[0]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- resident inputs (mesh arrays in HBM before the timed region) ------------------------
    mesh_e = copy.copy(mesh)
    mesh_e.cells = [copy.copy(c) for c in mesh.cells]
    wg.get_unique_edges(mesh_e)
    base = fs.VectorProblem(mesh_e, ior, k, device=dev)
    tris_d = dv.h2d(tris_h, dev)
    N, Ntt = base.N, base.Ntt
    stats = {}

    def device_step():
        wg.unique_edges_device(tris_d, npts)                       # K1
        prob = base.fresh_pattern()                                # K2 (built inside the solve)
        wd, Vd, info = fs.solve_vector_resident(prob, sigma, NMAX)  # K2, K4, K8, K5-K10
        prob.pattern.close()
        stats.update(info)
        return wd, Vd

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    barrier()
    l0 = _lib.launch_count()
    with ClockSampler(None if args.no_clock_sampler else local) as clk:
        t0e, t1e = ev(), ev()
        t0e.record()
        for _ in range(args.steps):
            if os.environ.get("WS_BENCH_TRACE"):
                torch.cuda.synchronize()
                _t = time.perf_counter()
            wd, Vd = device_step()
            if os.environ.get("WS_BENCH_TRACE"):
                torch.cuda.synchronize()
                print("[bench step] rank %d: %.1f ms" % (rank, (time.perf_counter() - _t) * 1e3), file=sys.stderr, flush=True)
        t1e.record()
        barrier()
        dev_ms = t0e.elapsed_time(t1e) / args.steps
    launches = (_lib.launch_count() - l0) // max(1, args.steps)
    nnz = base.pattern.nnz
    fstat = stats["factor"]

    if args.lite:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * 1e3 / dev_ms, "unit": "eigensolves/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "lite": True,
                              "gpu_launches": int(launches), "eig": {"op_applications": int(stats.get("n_op", 0))}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- dominant kernel family: the two solve sweeps (k_fwd_* + k_bwd_*), timed live -----------
    prob = base.fresh_pattern()
    prob.build_pattern()
    A, B = prob.assemble()
    Mq = prob.assemble_qd(sigma)
    sym = sv.Symbolic.on_device(prob.dofs, prob.xy, N, Ntt)
    _ex = sym.export()
    _p, _q = _ex["p"].astype(np.int64), _ex["q"].astype(np.int64)
    # entries of the inverted panels one sweep actually reads: the lower triangle of the pivot
    # block and the whole boundary block of every front (the stored rectangle m x p is larger)
    sweep_entries = int((_p * (_p + 1) // 2 + _q * _p).sum())
    S = sv.Solver(prob.pattern, sym)
    e0, e1 = ev(), ev()
    e0.record()
    S.factor(Mq)                 # first factorisation of a solver also emits the tile-task plan (host)
    e1.record()
    torch.cuda.synchronize()
    factor_first_ms = e0.elapsed_time(e1)
    e0.record()
    S.factor(Mq)
    e1.record()
    torch.cuda.synchronize()
    factor_ms = e0.elapsed_time(e1)
    b = torch.randn(N, dtype=torch.float64, device=dev)
    x = S.solve(b)
    nrep = 20
    e0.record()
    for _ in range(nrep):
        S.solve(b, out=x)
    e1.record()
    torch.cuda.synchronize()
    solve_ms = e0.elapsed_time(e1) / nrep
    y = torch.empty_like(b)
    prob.pattern.spmv(B, b, out=y)
    e0.record()
    for _ in range(nrep):
        prob.pattern.spmv(B, b, out=y)
    e1.record()
    torch.cuda.synchronize()
    spmv_ms = e0.elapsed_time(e1) / nrep
    AB = prob.assemble()
    e0.record()
    for _ in range(nrep):
        prob.assemble(out=AB)
    e1.record()
    torch.cuda.synchronize()
    asm_ms = e0.elapsed_time(e1) / nrep
    S.close()
    prob.pattern.close()

    # ---- end to end through the public API with host buffers ---------------------------------
    e2e_t = []
    h2d_b = d2h_b = 0
    # untimed passes first: the result arrays land in page-locked memory that torch's caching host
    # allocator only recycles once the previous call's arrays are dropped, so steady state (what a
    # sweep sees) is reached after two calls
    n_warm = max(1, min(args.warmup, 3))
    for it in range(max(1, min(args.steps, 3)) + n_warm):
        m = copy.copy(mesh)
        m.cells = [copy.copy(c) for c in mesh.cells]
        barrier()
        t0 = time.perf_counter()
        edges, eidx = ws.get_unique_edges(m)
        w, v, cnt = ws.solve_waveguide_vec(m, WL, ior, Nmax=NMAX, verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if os.environ.get("WS_BENCH_TRACE"):
            print("[bench e2e] pass %d: %.1f ms" % (it, dt * 1e3), file=sys.stderr, flush=True)
        if it >= n_warm:
            e2e_t.append(dt)
        h2d_b = tris_h.nbytes + base.h2d_bytes
        d2h_b = edges.nbytes + eidx.nbytes + m.edge_flips.nbytes + w.nbytes + v.nbytes
    e2e_s = float(np.median(e2e_t))          # host-side passes jitter (page faults, allocator); median of the timed passes
    e2e_mean = float(np.mean(e2e_t))
    # the same call with real_output=True (extension: float64 results instead of scipy's complex128 container, whose
    # imaginary part is zero): reported beside the headline, which keeps the reference dtype
    e2e_real = []
    for it in range(3):
        m = copy.copy(mesh)
        m.cells = [copy.copy(c) for c in mesh.cells]
        barrier()
        t0 = time.perf_counter()
        ws.get_unique_edges(m)
        w_, v_, cnt_ = ws.solve_waveguide_vec(m, WL, ior, Nmax=NMAX, verbose=False, real_output=True)
        torch.cuda.synchronize()
        if it:
            e2e_real.append(time.perf_counter() - t0)
        del w_, v_
    e2e_real_s = float(np.median(e2e_real))

    total = float(world)
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, solve_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, solve_ms = [float(x_) for x_ in t.tolist()]

    peak, peak_src = peaks()
    # algorithmic bytes of one solve (both sweeps): the inverted panels are read once per sweep
    solve_bytes = 2 * 8 * sweep_entries
    achieved = solve_bytes / (solve_ms * 1e-3) / 1e9
    n_op = int(stats.get("n_op", 0))
    line = {
        "metric": METRIC, "value": total / (dev_ms * 1e-3), "unit": "eigensolves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements": ne, "elements_per_gpu": ne, "N": N, "nnz": nnz, "Nmax": NMAX,
                   "ncv": max(2 * NMAX + 4, 20), "tol": "eps (scipy default tol=0)",
                   "l2": "inputs_larger_than_l2 (factor panels %.1f GB, basis %.2f GB per step)"
                         % (8 * fstat["nnzL"] / 1e9, 8 * 22 * N / 1e9),
                   "parallelism": "one problem per GPU"},
        "assembly": {"elements_per_s": ne / (asm_ms * 1e-3), "kernel_ms": asm_ms,
                     "algorithmic_bytes": ne * 20 + 16 * npts + 16 * nnz,
                     "GBps": (ne * 20 + 16 * npts + 16 * nnz) / (asm_ms * 1e-3) / 1e9,
                     "frac_of_peak": (ne * 20 + 16 * npts + 16 * nnz) / (asm_ms * 1e-3) / 1e9 / peak},
        "spmv": {"kernel_ms": spmv_ms, "GBps": (12 * nnz + 20 * N) / (spmv_ms * 1e-3) / 1e9,
                 "frac_of_peak": (12 * nnz + 20 * N) / (spmv_ms * 1e-3) / 1e9 / peak},
        "factor": {"ms": factor_ms, "first_call_ms": factor_first_ms, "flops": fstat["flops"], "TFLOPs": fstat["flops"] / (factor_ms * 1e-3) / 1e12,
                   "nnzL": fstat["nnzL"], "sweep_entries": sweep_entries, "fronts": fstat["fronts"], "max_front": fstat["maxfront"],
                   "accuracy_guard": {"probe_backward_error": fstat.get("backward_error"), "perturbed_pivots": fstat.get("perturbed_pivots"),
                                      "refine_steps_per_solve": fstat.get("refine_steps"), "pivot_threshold": fstat.get("pivot_threshold")}},
        "eig": {"op_applications": n_op, "restarts": int(stats.get("restarts", 0)), "converged": int(stats.get("nconv", 0))},
        "roofline": {"bound": "hbm", "kernel": "k_fwd_front/k_fwd_tile + k_bwd_tile/k_bwd_front (one shift-invert solve = both sweeps "
                                               "over the inverted panels, one launch per tree level and direction; %d solves per "
                                               "eigensolve, the largest share of the step)" % n_op,
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": solve_bytes, "kernel_ms": solve_ms, "traffic": measured_traffic(),
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one solve, "
                                       "profiles/r01_traffic.json (ncu capture of tools/solve_only.py)"},
        "e2e": {"value": total / e2e_s, "unit": "eigensolves/s", "h2d_bytes_per_step": int(h2d_b),
                "d2h_bytes_per_step": int(d2h_b), "seconds": e2e_s, "seconds_mean": e2e_mean,
                "stat": "median of %d passes after %d untimed ones" % (len(e2e_t), n_warm),
                "seconds_real_output": e2e_real_s,
                "api": "get_unique_edges(mesh) + solve_waveguide_vec(mesh, wl, IOR_dict, Nmax=10) with host numpy in/out"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if args.sweep_problems > 0:
        sw = run_sweep_slice(args, dev, rank, world)
        if rank == 0:
            line["sweep"] = sw
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            r = cpu_port_run(args.cpu_sample_elements)
            line["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample", "host")}
            rec = same_config_record()
            if rec:
                line["cpu_baseline"]["same_config_build_host"] = rec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
