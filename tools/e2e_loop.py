import sys, os, time, copy, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import wavesolve_b200 as ws
mg = ws.meshgen
mesh = mg.fiber_disc(1_000_000, order=1, seed=0); ior = mg.FIBER_IOR
keep = None
for it in range(6):
    m = copy.copy(mesh); m.cells = [copy.copy(c) for c in mesh.cells]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if it == 4:
        pr = cProfile.Profile(); pr.enable()
    edges, eidx = ws.get_unique_edges(m)
    t1 = time.perf_counter()
    w, v, cnt = ws.solve_waveguide_vec(m, 1.55, ior, Nmax=10, verbose=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if it == 4:
        pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:2600])
    print("iter %d: edges %.1f ms, solve %.1f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
