#!/usr/bin/env python
"""C5-like sweep (BASELINE.json configs[4]): wavelengths x z-slices, one problem per GPU at a time,
meshes of ~50k triangles, vector modes, eigenpairs gathered with one NCCL collective (K11/K12).

    python tools/bench_sweep.py                       # 1 GPU
    torchrun --nproc-per-node N tools/bench_sweep.py  # N GPUs (NCCL)

Prints one JSON line on rank 0: problems/s (max-over-ranks wall time of the whole sweep incl. the
gather) and a parity check of a few gathered eigenpairs against individual solve_waveguide_vec calls."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=4)
    ap.add_argument("--wavelengths", type=int, default=8)
    ap.add_argument("--elements", type=int, default=50_000)
    ap.add_argument("--nmax", type=int, default=10)
    ap.add_argument("--lanes", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import wavesolve_b200 as ws
    from wavesolve_b200 import sweep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mg = ws.meshgen
    # z-slices: the multimode core radius grows along the taper (lantern-demo cell 4)
    radii = np.linspace(4.0, 6.5, a.meshes)
    meshes = [mg.fiber_disc(a.elements, r_core=float(r), order=1, seed=10 + i) for i, r in enumerate(radii)]
    for m in meshes:
        ws.get_unique_edges(m)
    wls = np.linspace(1.0, 1.8, a.wavelengths)
    probs = [sweep.SweepProblem(mi, float(wl), mg.FIBER_IOR) for mi in range(a.meshes) for wl in wls]
    # warm-up (kernels, allocator pools, NCCL communicator)
    sweep.solve_sweep(meshes[:1], probs[:2], Nmax=a.nmax, kind="vector")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = sweep.solve_sweep(meshes, probs, Nmax=a.nmax, kind="vector", lanes=a.lanes or None)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        for i in (0, len(probs) // 2, len(probs) - 1):
            pr = probs[i]
            w, v, cnt = ws.solve_waveguide_vec(meshes[pr.mesh_id], pr.wl, pr.IOR_dict, Nmax=a.nmax, verbose=False)
            k = 2 * np.pi / pr.wl
            ne_s, ne_i = np.sqrt(res[i].w) / k, np.sqrt(w.real) / k
            ok = ok and bool(np.max(np.abs(ne_s - ne_i) / ne_i) < 1e-10) and res[i].mode_count == cnt
            ok = ok and res[i].v.shape == v.shape
        owners = sorted(set(r.rank for r in res))
        print(json.dumps({"metric": "sweep problems/s (vector, %d modes, %d meshes x %d wavelengths, ~%dk triangles each)"
                                    % (a.nmax, a.meshes, a.wavelengths, a.elements // 1000),
                          "value": len(probs) / float(dt), "n_gpus": world, "seconds": float(dt), "problems": len(probs),
                          "N": int(len(meshes[0].cells[0].data) + len(meshes[0].points)), "ranks_that_solved": owners,
                          "gathered_bytes": int(sum(r.v.nbytes + r.w.nbytes for r in res)), "parity_ok": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
