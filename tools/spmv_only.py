#!/usr/bin/env python
"""SpMV (K5) alone on the 1M-triangle vector problem: time per product (CUDA events, 50 back-to-back launches) with a checksum
of the result bits (the variants of profiles/r02c_spmv_pipeline_experiment.txt were compared through it).  `python tools/spmv_only.py`; each variant runs in its own process (the knobs are read once)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs
    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    mesh = mg.fiber_disc(int(os.environ.get("WS_SPMV_ELEMENTS", "1000000")), order=1, seed=0)
    ws.get_unique_edges(mesh)
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    A, B = prob.assemble()
    N, nnz = prob.N, prob.pattern.nnz
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty_like(x)
    for _ in range(3):
        prob.pattern.spmv(B, x, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        prob.pattern.spmv(B, x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    by = 12 * nnz + 20 * N
    h = int(torch.frombuffer(bytearray(y.cpu().numpy().tobytes()), dtype=torch.int64).sum().item())
    print("%-28s %.4f ms  %.0f GB/s  (%.3f of 6545.6, %.3f of 8000)  checksum %d" % (
        os.environ.get("WS_SPMV_TAG", ""), ms, by / ms / 1e6, by / ms / 1e6 / 6545.6, by / ms / 1e6 / 8000, h), flush=True)


if __name__ == "__main__":
    if os.environ.get("WS_SPMV_CHILD"):
        child()
    else:
        variants = [("k_spmv_stream", dict())]
        for tag, env in variants:
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, WS_SPMV_CHILD="1", WS_SPMV_TAG=tag, **env))
