#!/usr/bin/env python
"""Mode resampling (K13-K15) on the device vs the reference's own path on the host cores.

    python tools/bench_interp.py [--elements 200000] [--grid 1000]

Device: P2 scalar mode of a lantern cross-section resampled on a grid x grid raster (locator build,
point location, weights, gather-sum), CUDA events, inputs resident.  Host: the numpy restatement of
the same functions (oracle/restated.py) on a small raster -- the literal reference needs one KD-tree
query per try and per point in a Python loop and is slower still."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=int, default=200_000)
    ap.add_argument("--grid", type=int, default=1000)
    ap.add_argument("--cpu-grid", type=int, default=100)
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import device as dv
    from oracle import restated as R

    mg = ws.meshgen
    mesh = mg.lantern_cross_section(a.elements, seed=3)
    R0 = float(np.abs(mesh.points[:, :2]).max())
    v = np.random.default_rng(0).standard_normal(len(mesh.points))
    xa = np.linspace(-R0, R0, a.grid)
    ya = np.linspace(-R0, R0, a.grid)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    t0 = time.perf_counter()
    loc = ws.construct_meshtree(mesh)
    torch.cuda.synchronize()
    t_build_first = time.perf_counter() - t0
    xd, yd, nx, ny = loc._points(xa, ya, True)
    vd = dv.h2d(v, loc.dev)
    res = {}
    for name, fn in (("locate", lambda: loc.locate(xd, yd, nx, ny, 0.95 * R0, 0)),):
        td = fn()
    wd = loc.weights(xd, yd, nx, ny, td, 2)

    def apply_dev():
        out = torch.empty((nx * ny, 1, 1), dtype=torch.float64, device=loc.dev)
        from wavesolve_b200 import _lib
        _lib.check(_lib.lib().ws_interp_apply(dv.ptr(td), nx * ny, loc.Ne, dv.ptr(loc.conn), loc.stride, 6, 1, dv.ptr(wd),
                                              dv.ptr(vd), 1, dv.ptr(out), loc.dev.index, dv.stream_ptr(loc.dev)))
        return out

    stages = {"locate": lambda: loc.locate(xd, yd, nx, ny, 0.95 * R0, 0),
              "weights": lambda: loc.weights(xd, yd, nx, ny, td, 2),
              "apply": apply_dev}
    for name, fn in stages.items():
        fn()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / 10
    npts = nx * ny
    inside = int((td >= 0).sum())
    # end to end through the drop-in (host arrays in / out)
    ws.interpolate(v, mesh, xa, ya, meshtree=loc, maxr=0.95 * R0)
    t0 = time.perf_counter()
    out = ws.interpolate(v, mesh, xa, ya, meshtree=loc, maxr=0.95 * R0)
    t_e2e = time.perf_counter() - t0
    # host baseline on a small raster
    xc = np.linspace(-R0, R0, a.cpu_grid)
    t0 = time.perf_counter()
    ref = R.interpolate_kdtree_loop(v, mesh, xc, xc, maxr=0.95 * R0)
    t_cpu = time.perf_counter() - t0
    chk = ws.interpolate(v, mesh, xc, xc, meshtree=loc, maxr=0.95 * R0)
    assert np.array_equal(np.isnan(ref), np.isnan(chk)) and np.nanmax(np.abs(ref - chk)) <= 1e-12 * (1 + np.nanmax(np.abs(ref)))
    dev_ms = sum(res.values())
    # algorithmic bytes per point: 16 (point, via two 1-D axes: ~0) + 8 (tri index) + 48 (weights) written and
    # read once, 24 + 48 (connectivity row + three vertices) + 48 (six field values) gathered, 8 out
    bytes_pp = 8 * 2 + 48 * 2 + 24 + 48 + 48 + 8
    print(json.dumps({
        "workload": "P2 mode of a lantern cross-section (%d triangles) on a %d x %d raster" % (mesh.num_elements, nx, ny),
        "points": npts, "inside": inside, "locator_build_ms_first_call": t_build_first * 1e3,
        **{k: round(v_, 4) for k, v_ in res.items()},
        "device_points_per_s": npts / (dev_ms * 1e-3), "device_GBps_algorithmic": npts * bytes_pp / (dev_ms * 1e-3) / 1e9,
        "e2e_points_per_s": npts / t_e2e, "e2e_seconds": t_e2e,
        "cpu_baseline": {"kind": "port", "cores": 1, "points_per_s": a.cpu_grid ** 2 / t_cpu,
                         "sample": "oracle/restated.py interpolate_kdtree_loop() (the reference's per-point KD-tree search + weight loop) "
                                   "on a %d x %d raster of the same mesh: %.2f s"
                                   % (a.cpu_grid, a.cpu_grid, t_cpu)},
        "nan_outside": int(np.isnan(out).sum())}))


if __name__ == "__main__":
    main()
