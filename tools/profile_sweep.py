#!/usr/bin/env python
"""cProfile of the host side of a multi-lane sweep (4 lantern meshes of ~50k triangles x 16 wavelengths, 8 lanes):
where the driving thread spends its time while the device works.  Diagnostic only."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavesolve_b200 as ws
from wavesolve_b200 import sweep
mg = ws.meshgen
nm = 4
meshes = [mg.lantern_cross_section(50000, order=1, seed=10 + i, r_clad_bundle=float(r)) for i, r in enumerate(np.linspace(9, 16, nm))]
for m in meshes:
    ws.get_unique_edges(m)
ior = mg.LANTERN_IOR
probs = sweep.problem_grid(nm, np.linspace(1.0, 1.8, 16), ior)
sweep.solve_sweep(meshes[:1], probs[:16], Nmax=10, kind="vector", lanes=8, gather=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
res = sweep.solve_sweep(meshes, probs, Nmax=10, kind="vector", lanes=8, gather=False)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("problems/s", len(probs) / dt, "seconds", dt)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue())
