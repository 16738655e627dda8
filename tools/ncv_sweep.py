"""Operator applications of the vector eigensolve as a function of the Krylov subspace size (ncv); diagnostic."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavesolve_b200 as ws
from wavesolve_b200 import fe_solver as fs, solver as sv
mg = ws.meshgen
k = 2 * np.pi / 1.55
for seed in [int(x) for x in os.environ.get("WS_SEEDS", "0,1,2").split(",")]:
    mesh = mg.fiber_disc(1_000_000, order=1, seed=seed)
    ws.get_unique_edges(mesh)
    sigma = (k * max(mg.FIBER_IOR.values())) ** 2
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
    sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
    S = sv.Solver(prob.pattern, sym)
    S.factor(Mq)
    for ncv in [int(x) for x in os.environ.get('WS_NCVS', '21,24,28,32,40,48').split(',')]:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        w, V, info = fs._eig_device(prob.pattern, S, B, sigma, 1, 10, prob.N, prob.dev, ncv=ncv, edges=prob.edges, xy=prob.xy, Ntt=prob.Ntt)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("seed %d ncv %d: ops %d restarts %d  %.1f ms  w0 %.12f" % (seed, ncv, info["n_op"], info["restarts"], dt * 1e3, float(w[0])), flush=True)
    S.close(); prob.pattern.close()
