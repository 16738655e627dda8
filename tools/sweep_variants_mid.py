#!/usr/bin/env python
"""Per-level pipeline depth of the one-CTA-per-front sweep kernels: levels with at most WS_MID_FRONTS fronts (about one
wave of CTAs, where residency does not matter) get their own WS_FWD_KB_MID / WS_BWD_CB_MID / WS_BWD_RJ_MID.  One process,
one factorisation; the knobs are re-read per solve (WS_TUNE_RELOAD=1).  Diagnostic only."""
import os
import sys

import numpy as np

os.environ["WS_TUNE_RELOAD"] = "1"
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
    xref = x.clone()
    keys = ("WS_MID_FRONTS", "WS_FWD_KB_MID", "WS_BWD_CB_MID", "WS_BWD_RJ_MID", "WS_FWD_KB", "WS_BWD_CB", "WS_BWD_RJ")
    combos = [dict()]
    for mf in (300, 1100, 2100, 4200):
        for kb, cb, rj in ((8, 1, 4), (16, 1, 4), (4, 2, 4), (4, 4, 4), (8, 2, 4), (8, 4, 4), (16, 4, 4), (16, 8, 4)):
            combos.append(dict(WS_MID_FRONTS=mf, WS_FWD_KB_MID=kb, WS_BWD_CB_MID=cb, WS_BWD_RJ_MID=rj))
    combos.append(dict())
    for c in combos:
        for kk in keys:
            os.environ.pop(kk, None)
        for kk, v in c.items():
            os.environ[kk] = str(v)
        for _ in range(3):
            S.solve(b, out=x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            S.solve(b, out=x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print(c, "solve %.4f ms  max|dx| %.1e" % (ms, float((x - xref).abs().max())), flush=True)


if __name__ == "__main__":
    main()
