"""Operator applications / time of the vector eigensolve against the Krylov basis width on SMALL problems (C5 size:
~50k triangles, lantern and fibre cross-sections); diagnostic for the default ncv."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavesolve_b200 as ws
from wavesolve_b200 import fe_solver as fs, solver as sv
mg = ws.meshgen
for name, mesh, ior in (("lantern", mg.lantern_cross_section(50000, order=1, seed=7), mg.LANTERN_IOR),
                        ("lantern r=16", mg.lantern_cross_section(50000, order=1, seed=8, r_clad_bundle=16.0), mg.LANTERN_IOR),
                        ("fibre", mg.fiber_disc(50000, order=1, seed=3), mg.FIBER_IOR)):
    ws.get_unique_edges(mesh)
    for wl in (1.0, 1.55):
        k = 2 * np.pi / wl
        sigma = (k * max(ior.values())) ** 2
        prob = fs.VectorProblem(mesh, ior, k)
        Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
        sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
        S = sv.Solver(prob.pattern, sym)
        S.factor(Mq)
        for ncv in [int(x) for x in os.environ.get("WS_NCVS", "24,28,32,40,48").split(",")]:
            best = 1e9
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                w, V, info = fs._eig_device(prob.pattern, S, B, sigma, 1, 10, prob.N, prob.dev, ncv=ncv, edges=prob.edges, xy=prob.xy, Ntt=prob.Ntt)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
            print("%-12s wl %.2f ncv %d: ops %d restarts %d  %.1f ms" % (name, wl, ncv, info["n_op"], info["restarts"], best * 1e3), flush=True)
        S.close(); prob.pattern.close()
