#!/usr/bin/env python
"""Stage-by-stage timing of the vector pipeline (synchronising between stages; diagnostic only,
bench.py is the measurement of record)."""
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
    ap.add_argument("--kind", default="vector", choices=["vector", "scalar"])
    ap.add_argument("--leaf", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--nsolves", type=int, default=20)
    ap.add_argument("--host-symbolic", action="store_true")
    a = ap.parse_args()
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs, solver as sv, waveguide as wg, _lib

    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    order = 1 if a.kind == "vector" else 2
    mesh = mg.fiber_disc(a.elements, order=order, seed=0)
    ior = mg.FIBER_IOR
    sigma = (k * max(ior.values())) ** 2
    leaf = a.leaf or sv.DEFAULT_LEAF_ELEMS

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(a.reps):
        T = {}
        t0 = sync()
        if a.kind == "vector":
            wg.get_unique_edges(mesh)
            T["edges(api)"] = sync() - t0
            t0 = sync()
            prob = fs.VectorProblem(mesh, ior, k)
        else:
            prob = fs.ScalarProblem(mesh, ior, k)
        T["h2d+pattern"] = sync() - t0
        t0 = sync()
        A, B = prob.assemble()
        T["assemble"] = sync() - t0
        t0 = sync()
        if a.kind == "vector":
            Mq = prob.assemble_qd(sigma)
            T["assemble_qd"] = sync() - t0
        t0 = sync()
        sym = (sv.Symbolic if a.host_symbolic else sv.Symbolic.on_device)(*((prob.h_dofs, prob.h_xy) if a.host_symbolic else (prob.dofs, prob.xy)), prob.N, prob.first_class_start, leaf)
        T["symbolic"] = sync() - t0
        T["sym.kd(1)"], T["sym.order(1)"], T["sym.bnd(1)"] = sym.t_kd, sym.t_order, sym.t_bnd
        t0 = sync()
        S = sv.Solver(prob.pattern, sym)
        T["solver_create"] = sync() - t0
        t0 = sync()
        st = S.factor(Mq) if a.kind == "vector" else S.factor(A, B, sigma)
        T["factor"] = sync() - t0
        b = torch.randn(prob.N, dtype=torch.float64, device="cuda")
        x = S.solve(b)
        t0 = sync()
        for _ in range(a.nsolves):
            x = S.solve(b, out=x)
        T["solve(1)"] = (sync() - t0) / a.nsolves
        y = torch.empty_like(b)
        prob.pattern.spmv(B, b, out=y)
        t0 = sync()
        for _ in range(a.nsolves):
            prob.pattern.spmv(B, b, out=y)
        T["spmv(1)"] = (sync() - t0) / a.nsolves
        t0 = sync()
        if a.kind == "vector":
            w, V, info = fs._eig_device(prob.pattern, S, B, sigma, 1, 10, prob.N, prob.dev,
                                        edges=prob.edges, xy=prob.xy, Ntt=prob.Ntt)
        else:
            w, V, info = fs._eig_device(prob.pattern, S, B, sigma, 0, 10, prob.N, prob.dev)
        T["eig"] = sync() - t0
        out = {"rep": rep, "kind": a.kind, "Ne": mesh.num_elements, "N": prob.N, "nnz": prob.pattern.nnz,
               "nnzL": st["nnzL"], "flops": st["flops"], "depth": st["depth"], "maxfront": st["maxfront"],
               "bytes_solver": st["bytes"], "n_op": info["n_op"], "restarts": info["restarts"],
               "factor_tflops": st["flops"] / T["factor"] / 1e12,
               "solve_GBs": 2 * 8 * st["nnzL"] / T["solve(1)"] / 1e9,
               "spmv_GBs": (12 * prob.pattern.nnz + 20 * prob.N) / T["spmv(1)"] / 1e9,
               "cpus": os.cpu_count(), "stages_ms": {k_: round(v * 1e3, 3) for k_, v in T.items()},
               "total_ms": round(sum(v for k_, v in T.items() if "(1)" not in k_) * 1e3, 2)}
        print(json.dumps(out), flush=True)
        S.close()
        del S, sym, prob


if __name__ == "__main__":
    main()
