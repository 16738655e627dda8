#!/usr/bin/env python
"""Kernel micro-benchmarks on the 1M-element problems (CUDA events, inputs > L2): assembly,
SpMV, factorisation, one shift-invert solve.  Diagnostic companion of bench.py."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(torch, fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--sorted", type=int, default=0)
    ap.add_argument("--solver", type=int, default=1)
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv

    peak = 6545.6
    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    out = {"elements": a.elements, "sorted": a.sorted}
    # ---- scalar P2
    mesh = mg.fiber_disc(a.elements, order=2, seed=0, spatial_sort=bool(a.sorted))
    prob = fs.ScalarProblem(mesh, mg.FIBER_IOR, k)
    ne, nnz, N, npts = mesh.num_elements, prob.pattern.nnz, prob.N, mesh.points.shape[0]
    AB = prob.assemble()
    ms = timeit(torch, lambda: prob.assemble(out=AB))
    by = ne * (24 + 8) + 16 * (N - 0) * 0 + 16 * (npts) + 16 * nnz
    out["scalar_assemble_ms"] = ms
    out["scalar_assemble_GBs"] = by / ms / 1e6
    out["scalar_assemble_frac"] = by / ms / 1e6 / peak
    out["scalar_elements_per_s"] = ne / ms * 1e3
    A, B = prob.assemble()
    x = torch.randn(N, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = timeit(torch, lambda: prob.pattern.spmv(B, x, out=y))
    out["scalar_spmv_ms"] = ms
    out["scalar_spmv_frac"] = (12 * nnz + 20 * N) / ms / 1e6 / peak
    ms = timeit(torch, lambda: fs.DevicePattern(prob.dofs, prob.N).close(), n=5, warm=1)
    out["scalar_pattern_ms"] = ms
    del prob, A, B, AB
    # ---- vector
    mesh = mg.fiber_disc(a.elements, order=1, seed=0, spatial_sort=bool(a.sorted))
    ws.get_unique_edges(mesh)
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    ne, nnz, N, npts = mesh.num_elements, prob.pattern.nnz, prob.N, mesh.points.shape[0]
    AB = prob.assemble()
    ms = timeit(torch, lambda: prob.assemble(out=AB))
    del AB
    by = ne * 20 + 16 * npts + 16 * nnz
    out["vector_assemble_ms"] = ms
    out["vector_assemble_GBs"] = by / ms / 1e6
    out["vector_assemble_frac"] = by / ms / 1e6 / peak
    out["vector_elements_per_s"] = ne / ms * 1e3
    sigma = (k * max(mg.FIBER_IOR.values())) ** 2
    out["vector_qd_ms"] = timeit(torch, lambda: prob.assemble_qd(sigma))
    A, B = prob.assemble()
    x = torch.randn(N, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = timeit(torch, lambda: prob.pattern.spmv(B, x, out=y))
    out["vector_spmv_ms"] = ms
    out["vector_spmv_frac"] = (12 * nnz + 20 * N) / ms / 1e6 / peak
    tris_d = prob.tris
    from wavesolve_b200 import waveguide as wg
    t32 = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.cells[1].data)[:, :3].astype(np.int32))).cuda()
    out["edges_K1_ms"] = timeit(torch, lambda: wg.unique_edges_device(t32, npts), n=5, warm=1)
    out["vector_pattern_ms"] = timeit(torch, lambda: fs.DevicePattern(prob.dofs, prob.N).close(), n=5, warm=1)
    if a.solver:
        Mq = prob.assemble_qd(sigma)
        sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
        out["symbolic_ms"] = [round(v * 1e3, 1) for v in (sym.t_kd, sym.t_order, sym.t_bnd)]
        S = sv.Solver(prob.pattern, sym)
        st = S.factor(Mq)
        ms = timeit(torch, lambda: S.factor(Mq), n=3, warm=0)
        out["factor_ms"] = ms
        out["factor_TFLOPs"] = st["flops"] / ms / 1e9
        xs = torch.empty_like(x)
        ms = timeit(torch, lambda: S.solve(x, out=xs))
        out["solve_ms"] = ms
        out["solve_GBs"] = 16 * st["nnzL"] / ms / 1e6
        out["solve_frac"] = 16 * st["nnzL"] / ms / 1e6 / peak
        out["nnzL"] = st["nnzL"]
    print(json.dumps({k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in out.items()}))


if __name__ == "__main__":
    main()
