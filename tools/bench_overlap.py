#!/usr/bin/env python
"""Overlap matrix G = U B V^T (K16/K17) on the device vs scipy on the host cores.

    python tools/bench_overlap.py [--elements 1000000] [--modes 10]

Device: B of the scalar P2 problem resident as (pattern, values), modes resident as rows; CUDA events
around `MassOperator.overlap` (SpMV for two columns at a time + Gram reduction).  Also timed through the
public call with host numpy modes in and the numpy matrix out.  Host: the reference user's expression
`U @ (B @ V.T)` with scipy CSR on the same data."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=int, default=1_000_000)
    ap.add_argument("--modes", type=int, default=10)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws

    mesh = ws.meshgen.fiber_disc(a.elements, seed=0)
    op = ws.mass_operator(mesh)
    N, nnz = op.N, op.pattern.nnz
    rng = np.random.default_rng(0)
    U = rng.standard_normal((a.modes, N))
    Ud = torch.from_numpy(U).cuda()
    op.overlap(Ud)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        G = op.overlap(Ud)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / a.reps
    # algorithmic bytes: per PAIR of columns the (value, column) stream + row pointers, two x gathers and two
    # y writes, then the nu rows of U and the two y chunks for the Gram partials
    pairs = (a.modes + 1) // 2
    alg = pairs * (12 * nnz + 4 * N + 4 * 8 * N + (a.modes + 2) * 8 * N)
    op.overlap(U)
    t0 = time.perf_counter()
    for _ in range(3):
        Gh = op.overlap(U)
    api_ms = (time.perf_counter() - t0) / 3 * 1e3
    Bs = ws.construct_B(mesh, sparse=True)
    t0 = time.perf_counter()
    Gref = U @ (Bs @ U.T)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    err = float(np.abs(Gh - Gref).max() / np.abs(Gref).max())
    out = {"elements": a.elements, "N": int(N), "nnz": int(nnz), "modes": a.modes,
           "device_ms": dev_ms, "algorithmic_GB": alg / 1e9, "device_GBps": alg / dev_ms / 1e6,
           "api_host_in_out_ms": api_ms, "scipy_host_ms": cpu_ms, "max_rel_err_vs_scipy": err,
           "pairs_per_s_device": a.modes * a.modes / dev_ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
