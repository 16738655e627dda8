#!/usr/bin/env python
"""Where one small sweep problem (~50k triangles, vector, 10 modes) spends its time: phases of the per-wavelength
work on a resident mesh context, then problems/s against the number of concurrent lanes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wavesolve_b200 as ws  # noqa: E402
from wavesolve_b200 import fe_solver as fs, sweep  # noqa: E402

mg = ws.meshgen
ne = int(os.environ.get("NE", "50000"))
mesh = mg.lantern_cross_section(ne, order=1, seed=1)
ws.get_unique_edges(mesh)
ior = mg.LANTERN_IOR
ctx = sweep._MeshContext(mesh, "vector", ior, 2 * np.pi / 1.55, torch.device("cuda", 0), None)
N = ctx.prob.N
out = {"N": N, "elements": mesh.num_elements}
ph = {}
for rep, wl in enumerate(np.linspace(1.0, 1.8, 6)):
    k = 2 * np.pi / wl
    sigma = float((k * max(ior.values())) ** 2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.prob.set_materials(mesh, ior, k)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    Mq, B = ctx.prob.assemble_qd(sigma, with_B=True, fast=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = ctx.solver.factor(Mq)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    wd, Vd, info = fs._eig_device(ctx.prob.pattern, ctx.solver, B, sigma, 1, 10, N, ctx.prob.dev, edges=ctx.prob.edges,
                                  xy=ctx.prob.xy, Ntt=ctx.prob.Ntt)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    if rep >= 2:
        for name, a, b in (("set_materials", t0, t1), ("assemble", t1, t2), ("factor", t2, t3), ("eig", t3, t4)):
            ph.setdefault(name, []).append((b - a) * 1e3)
        ph.setdefault("n_op", []).append(info["n_op"])
        ph.setdefault("restarts", []).append(info["restarts"])
out["phases_ms"] = {k_: float(np.mean(v)) for k_, v in ph.items()}
b = torch.randn(N, dtype=torch.float64, device="cuda")
x = ctx.solver.solve(b)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50):
    ctx.solver.solve(b, out=x)
torch.cuda.synchronize()
out["solve_ms"] = (time.perf_counter() - t0) / 50 * 1e3
ctx.close()
nm = int(os.environ.get("MESHES", "2"))
meshes = [mg.lantern_cross_section(ne, order=1, seed=10 + i, r_clad_bundle=float(r)) for i, r in enumerate(np.linspace(9, 16, nm))]
for m in meshes:
    ws.get_unique_edges(m)
probs = sweep.problem_grid(nm, np.linspace(1.0, 1.8, 16), ior)
sweep.solve_sweep(meshes[:1], probs[:16], Nmax=10, kind="vector", lanes=16, gather=False)
for lanes in [int(x) for x in os.environ.get("LANES", "1,2,4,8").split(",")]:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = sweep.solve_sweep(meshes, probs, Nmax=10, kind="vector", lanes=lanes, gather=False)
    torch.cuda.synchronize()
    out["lanes_%d_problems_per_s" % lanes] = len(probs) / (time.perf_counter() - t0)
print(json.dumps(out))
