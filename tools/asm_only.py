#!/usr/bin/env python
"""Assembly kernels alone on the 1M-element meshes (for ncu launch lists / --set full captures):
`python tools/asm_only.py [vector|scalar] [reps]`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "vector"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ne = int(os.environ.get("WS_ASM_ELEMENTS", "1000000"))
    import torch
    import wavesolve_b200 as ws
    from wavesolve_b200 import fe_solver as fs
    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    if kind == "scalar":
        mesh = mg.fiber_disc(ne, order=2, seed=0)
        prob = fs.ScalarProblem(mesh, mg.FIBER_IOR, k)
    else:
        mesh = mg.fiber_disc(ne, order=1, seed=0)
        ws.get_unique_edges(mesh)
        prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    AB = prob.assemble()
    big = torch.empty(64 << 20, dtype=torch.float64, device="cuda")   # 512 MB: flushes L2 between repetitions
    for _ in range(reps):
        big.zero_()
        prob.assemble(out=AB)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prob.assemble(out=AB)
    e1.record()
    torch.cuda.synchronize()
    print(kind, "assemble ms (back to back, CUDA events):", e0.elapsed_time(e1) / 20, "Ne", mesh.num_elements, "nnz", prob.pattern.nnz)


if __name__ == "__main__":
    main()
