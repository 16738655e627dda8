#!/usr/bin/env python
"""Cost of every level of the solve sweeps INSIDE the overlapped pipeline: the solve is cut after its first n launches
(WS_SWEEP_STOP=n, re-read per solve) for n = 1 .. 32 and timed with CUDA events over 30 back-to-back repetitions; the
difference between consecutive cuts is what level n adds.  (ncu serialises the launches: its per-launch times sum to
1.3 ms where the solve takes 0.95 ms.)  Diagnostic only."""
import os
import sys

import numpy as np

os.environ["WS_TUNE_RELOAD"] = "1"          # WS_SWEEP_STOP is re-read per solve
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    mesh = mg.fiber_disc(ne, order=1, seed=0)
    ws.get_unique_edges(mesh)
    sigma = (k * max(mg.FIBER_IOR.values())) ** 2
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    Mq = prob.assemble_qd(sigma)
    sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt, sv.DEFAULT_LEAF_ELEMS)
    S = sv.Solver(prob.pattern, sym)
    st = S.factor(Mq)
    b = torch.randn(prob.N, dtype=torch.float64, device="cuda")
    x = torch.empty_like(b)
    S.solve(b, out=x)
    nl = 2 * (st["depth"] + 1)
    prev = 0.0
    for n in list(range(1, nl + 1)):
        os.environ["WS_SWEEP_STOP"] = str(n)
        for _ in range(3):
            S.solve(b, out=x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            S.solve(b, out=x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print("first %2d launches: %.1f us   (+%.1f us)" % (n, ms * 1e3, (ms - prev) * 1e3), flush=True)
        prev = ms


if __name__ == "__main__":
    main()
