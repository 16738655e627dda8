#!/usr/bin/env python
"""Wall-clock breakdown (with a synchronize after every phase) of bench.py's device-resident
step on the 1M-element vector problem; diagnostic only."""
import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv, waveguide as wg, device as dv

    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    mesh = mg.fiber_disc(int(os.environ.get("WS_PROFILE_ELEMENTS", "1000000")), order=1, seed=0)
    ior = mg.FIBER_IOR
    sigma = float(np.power(k * max(ior.values()), 2))
    tris_h = np.ascontiguousarray(np.asarray(mesh.cells[1].data)[:, :3].astype(np.int32))
    npts = mesh.points.shape[0]
    wg.get_unique_edges(mesh)
    base = fs.VectorProblem(mesh, ior, k)
    tris_d = dv.h2d(tris_h, base.dev)

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(4):
        T = {}
        t00 = t0 = sync()
        wg.unique_edges_device(tris_d, npts)
        t1 = sync(); T["K1 edges"] = t1 - t0; t0 = t1
        prob = base.fresh_pattern()
        prob.build_pattern()
        t1 = sync(); T["K2 pattern"] = t1 - t0; t0 = t1
        Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
        t1 = sync(); T["K4 assemble"] = t1 - t0; t0 = t1
        sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
        t1 = sync(); T["K8a symbolic"] = t1 - t0; t0 = t1
        S = sv.Solver(prob.pattern, sym)
        t1 = sync(); T["solver create (plan)"] = t1 - t0; t0 = t1
        st = S.factor(Mq)
        t1 = sync(); T["K8b factor"] = t1 - t0; t0 = t1
        wd, Vd, info = fs._eig_device(prob.pattern, S, B, sigma, 1, 10, prob.N, prob.dev, edges=prob.edges, xy=prob.xy, Ntt=prob.Ntt)
        t1 = sync(); T["Krylov"] = t1 - t0; t0 = t1
        S.close(); sym.close(); prob.pattern.close()
        del S, sym, Mq, B, wd, Vd
        t1 = sync(); T["close"] = t1 - t0
        T["total"] = t1 - t00
        T["n_op"] = info["n_op"]
        print(json.dumps({k_: (round(v * 1e3, 2) if k_ != "n_op" else v) for k_, v in T.items()}), flush=True)


if __name__ == "__main__":
    main()
