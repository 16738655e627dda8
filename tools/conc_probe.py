#!/usr/bin/env python
"""How many independent shift-invert solves of a ~100k-unknown problem a B200 overlaps when ONE host thread enqueues them
round-robin on several streams (NL solvers, ACT = numbers of streams to try): 0.355 ms per solve alone, 0.112 ms with 4
streams, 0.070 ms with 8 (then bound by the host's launch rate).  The measurement behind sweep.default_lanes."""
import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import wavesolve_b200 as ws
from wavesolve_b200 import fe_solver as fs, sweep
mg = ws.meshgen
mesh = mg.lantern_cross_section(50000, order=1, seed=1)
ws.get_unique_edges(mesh)
ior = mg.LANTERN_IOR
k = 2*np.pi/1.55
dev = torch.device("cuda", 0)
nl = int(os.environ.get("NL", "4"))
ctxs = [sweep._MeshContext(mesh, "vector", ior, k, dev, None) for _ in range(nl)]
sigma = float((k*max(ior.values()))**2)
N = ctxs[0].prob.N
keep = []
for c in ctxs:
    Mq, B = c.prob.assemble_qd(sigma, with_B=True, fast=True)
    keep.append((Mq, B))
    c.solver.factor(Mq)
streams = [torch.cuda.Stream() for _ in range(nl)]
bs = [torch.randn(N, dtype=torch.float64, device=dev) for _ in range(nl)]
xs = [torch.empty_like(b) for b in bs]
torch.cuda.synchronize()
for n_act in [int(x) for x in os.environ.get("ACT", "1,2,4").split(",")]:
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(100):
            for l in range(n_act):
                with torch.cuda.stream(streams[l]):
                    ctxs[l].solver.solve(bs[l], out=xs[l])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("streams %d: %.3f ms per solve-round, %.3f ms per solve" % (n_act, dt/100*1e3, dt/100/n_act*1e3))
