#!/usr/bin/env python
"""One small pass over the whole hot path for compute-sanitizer (tools/sanitize.sh): edge table, pattern, scalar and
vector assembly, symbolic analysis, factorisation (fused + tiled fronts), flag-synchronised solve sweeps, both Krylov
iterations (narrow and wide basis), a two-lane sweep.  Sizes are small: the sanitizer slows kernels by 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wavesolve_b200 as ws  # noqa: E402
from wavesolve_b200 import sweep  # noqa: E402

mg = ws.meshgen
ne = int(os.environ.get("NE", "6000"))
m2 = mg.fiber_disc(ne, seed=1)
w, v, c = ws.solve_waveguide(m2, 1.55, mg.FIBER_IOR, Nmax=4)
print("scalar", np.sqrt(w) / (2 * np.pi / 1.55), c)
m1 = mg.fiber_disc(ne, order=1, seed=2)
ws.get_unique_edges(m1)
w, v, c = ws.solve_waveguide_vec(m1, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False)
print("vector", np.sqrt(w.real) / (2 * np.pi / 1.55), c)
w, v, c = ws.solve_waveguide_vec(m1, 1.55, mg.FIBER_IOR, Nmax=40, verbose=False)      # wide basis: DMMA rotation
print("vector Nmax=40", c)
w, v, c = ws.solve_waveguide(mg.holey_fiber(ne, seed=3), 1.55, mg.HOLEY_IOR, Nmax=3, target_neff=1.40)   # refinement path
print("interior shift", np.sqrt(w) / (2 * np.pi / 1.55), ws.fe_solver.last_info["factor"]["refine_steps"])
res = sweep.solve_sweep([m1], [sweep.SweepProblem(0, wl, mg.FIBER_IOR) for wl in (1.5, 1.6)], Nmax=3, kind="vector", lanes=2)
print("sweep", [r.rc for r in res])
torch.cuda.synchronize()
