#!/usr/bin/env python
"""Factor once and run a few shift-invert solves on the 1M-element vector problem: a short
command for ncu captures of the factorisation / solve-sweep kernels."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--solves", type=int, default=3)
    ap.add_argument("--factors", type=int, default=1)
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--leaf", type=int, default=0)
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv

    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    mesh = mg.fiber_disc(a.elements, order=1, seed=0)
    ws.get_unique_edges(mesh)
    sigma = (k * max(mg.FIBER_IOR.values())) ** 2
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    A, B = prob.assemble()
    Mq = prob.assemble_qd(sigma)
    sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt, a.leaf or sv.DEFAULT_LEAF_ELEMS)
    S = sv.Solver(prob.pattern, sym)
    for _ in range(a.factors):
        st = S.factor(Mq)
    b = torch.randn(prob.N, dtype=torch.float64, device="cuda") if a.nrhs == 1 else \
        torch.randn((a.nrhs, prob.N), dtype=torch.float64, device="cuda")
    x = torch.empty_like(b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    S.solve(b, out=x)
    e0.record()
    for _ in range(a.solves):
        S.solve(b, out=x)
    e1.record()
    torch.cuda.synchronize()
    y = prob.pattern.spmv(B, x if a.nrhs == 1 else x[0])
    print("N %d nnzL %d solve %.3f ms  (%.0f GB/s)  depth %d" % (prob.N, st["nnzL"], e0.elapsed_time(e1) / a.solves,
          16 * st["nnzL"] / (e0.elapsed_time(e1) / a.solves * 1e-3) / 1e9, st["depth"]))


if __name__ == "__main__":
    main()
