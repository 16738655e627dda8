#!/usr/bin/env python
"""cProfile of the public-API path (get_unique_edges + solve_waveguide_vec with host numpy in/out)
on the 1M-element vector problem: where the host-side time of bench.py's `e2e` goes."""
import argparse
import cProfile
import copy
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws

    mg = ws.meshgen
    mesh = mg.fiber_disc(a.elements, order=1, seed=0)
    ior = mg.FIBER_IOR

    def once():
        m = copy.copy(mesh)
        m.cells = [copy.copy(c) for c in mesh.cells]
        ws.get_unique_edges(m)
        w, v, cnt = ws.solve_waveguide_vec(m, 1.55, ior, Nmax=10, verbose=False)
        torch.cuda.synchronize()
        return w

    for _ in range(2):
        t0 = time.perf_counter()
        once()
        print("e2e %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    once()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(a.top)
    print(s.getvalue())


if __name__ == "__main__":
    main()
