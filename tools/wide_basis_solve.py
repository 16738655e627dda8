#!/usr/bin/env python
"""One vector eigensolve with a wide Krylov basis (Nmax = 80 -> ncv = 164, lantern-demo.ipynb cell 4) on a ~50k-triangle
lantern cross-section: the command the ncu capture of the DMMA restart rotation (k_rotate_mma) and of the CGS kernels is
taken from (tools/gpu_profile.sh)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wavesolve_b200 as ws  # noqa: E402

mg = ws.meshgen
mesh = mg.lantern_cross_section(int(os.environ.get("NE", "50000")), order=1, seed=7)
ws.get_unique_edges(mesh)
nmax = int(os.environ.get("NMAX", "80"))
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w, v, cnt = ws.solve_waveguide_vec(mesh, 1.55, mg.LANTERN_IOR, Nmax=nmax, verbose=False, sparse_solve_mode="pardiso")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
info = ws.fe_solver.last_info
print("Nmax %d N %d: %.1f ms, %d OP applications, %d restarts, %d modes counted" % (nmax, info["N"], dt * 1e3, info["n_op"], info["restarts"], cnt))
