#!/usr/bin/env python
"""What a sweep pays per MESH before its first wavelength: device copies of the mesh, pattern, symbolic analysis, one solver
per lane (front storage + gather maps) and each solver's first factorisation (which emits its tile-task plan), on a
~50k-triangle lantern mesh with 8 lanes.  Wall-clock per phase with a synchronize after each.  Diagnostic only."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv
    mg = ws.meshgen
    ior = mg.LANTERN_IOR
    k = 2 * np.pi / 1.55
    meshes = [mg.lantern_cross_section(50000, order=1, seed=10 + i, r_clad_bundle=float(r)) for i, r in enumerate(np.linspace(9, 16, 3))]
    for m in meshes:
        ws.get_unique_edges(m)
    dev = torch.device("cuda", 0)

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep, mesh in enumerate(meshes):
        T = {}
        t0 = sync()
        prob = fs.VectorProblem(mesh, ior, k, dev, build_pattern=False)
        t1 = sync(); T["problem (H2D, ids)"] = t1 - t0; t0 = t1
        prob.build_pattern()
        t1 = sync(); T["pattern"] = t1 - t0; t0 = t1
        sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt, sv.DEFAULT_LEAF_ELEMS)
        t1 = sync(); T["symbolic"] = t1 - t0; t0 = t1
        prob.apply_T(torch.zeros(prob.N, dtype=torch.float64, device=dev))
        t1 = sync(); T["T tables"] = t1 - t0; t0 = t1
        solvers = [sv.Solver(prob.pattern, sym) for _ in range(8)]
        t1 = sync(); T["8 x solver create"] = t1 - t0; t0 = t1
        bufs = [(torch.empty_like(prob.k2n2), torch.empty(prob.pattern.nnz, dtype=torch.float64, device=dev),
                 torch.empty(prob.pattern.nnz, dtype=torch.float64, device=dev)) for _ in range(8)]
        t1 = sync(); T["8 x value arrays"] = t1 - t0; t0 = t1
        sigma = float((k * max(ior.values())) ** 2)
        Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
        t1 = sync(); T["assemble"] = t1 - t0; t0 = t1
        for S in solvers:
            S.factor(Mq)
        t1 = sync(); T["8 x first factor (plan)"] = t1 - t0; t0 = t1
        for S in solvers:
            S.factor(Mq)
        t1 = sync(); T["8 x second factor"] = t1 - t0; t0 = t1
        for S in solvers:
            S.close()
        sym.close(); prob.pattern.close()
        t1 = sync(); T["close"] = t1 - t0
        print("mesh %d (N = %d): " % (rep, prob.N) + ", ".join("%s %.1f ms" % (k_, v * 1e3) for k_, v in T.items()), flush=True)


if __name__ == "__main__":
    main()
