"""TEST INFRASTRUCTURE ONLY.  Golden eigenvalues of the BASELINE configs at their STATED sizes.

Run in the build container (one-off; the 1M case takes tens of minutes and ~15 GB):

    python -m oracle.make_golden_at_size c3      # ~200k P2 triangles, lantern, LITERAL reference
    python -m oracle.make_golden_at_size c4      # ~1M P1 triangles, fibre, oracle/restated.py
    python -m oracle.make_golden_at_size c2      # ~50k P1 triangles, fibre, oracle/restated.py

The fixtures are tiny: the mesh is NOT stored, only the generator arguments (``meshgen`` is
deterministic for a seed on this image) and SHA-256 digests of the mesh arrays, which the test
re-checks before it compares anything; then the eigenvalues, the mode count and the timing of
the CPU run (that timing is also the same-configuration CPU baseline quoted in DESIGN.md).

  c3: ``solve_waveguide`` of the unmodified reference (fe_solver.py:292-348) through
      oracle/refshim.py -- the literal code can run the scalar path at any size.
  c4, c2: ``oracle.restated.solve_waveguide_vec`` -- the literal vector path cannot run beyond
      ~5k elements; the restatement is pinned on the literal code by
      tests/test_oracle_golden.py and tests/test_oracle_live_reference.py.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from wavesolve_b200 import meshgen as mg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WL = 1.55


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).data).hexdigest()


def mesh_digests(mesh):
    out = {"sha_points": digest(np.asarray(mesh.points, dtype=np.float64)),
           "sha_tris": digest(np.asarray(mesh.cells[1].data).astype(np.int64))}
    for i, lab in enumerate(mesh.cell_sets.keys()):
        out["sha_set_%d" % i] = digest(np.asarray(mesh.cell_sets[lab][1]).astype(np.int64))
    return out


def save(name, **kw):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in kw.items()})
    print("wrote", path, flush=True)


def c3(target=200_000, seed=5, Nmax=10):
    from oracle import refshim
    fe, sf, wg = refshim.load()
    mesh = mg.lantern_cross_section(target, seed=seed)
    k = 2 * np.pi / WL
    t0 = time.perf_counter()
    w, v, cnt = fe.solve_waveguide(mesh, WL, mg.LANTERN_IOR, sparse=True, Nmax=Nmax, order=2)
    dt = time.perf_counter() - t0
    print("c3 literal: Ne", mesh.num_elements, "neff", np.sqrt(w) / k, "count", cnt, "%.1f s" % dt, flush=True)
    save("c3_lantern_200k", generator="lantern_cross_section", target=target, seed=seed, order=2, Nmax=Nmax, wl=WL,
         elements=mesh.num_elements, w=np.asarray(w), count=cnt, seconds=dt, source="literal reference fe_solver.solve_waveguide",
         labels=np.array(list(mesh.cell_sets.keys())), **mesh_digests(mesh))


def vec(name, target, seed, Nmax=10):
    from oracle import restated as R
    mesh = mg.fiber_disc(target, order=1, seed=seed)
    dg = mesh_digests(mesh)
    k = 2 * np.pi / WL
    t0 = time.perf_counter()
    R.get_unique_edges(mesh)
    t1 = time.perf_counter()
    A, B = R.construct_AB_vec(mesh, mg.FIBER_IOR, k)
    t2 = time.perf_counter()
    w, v, cnt, info = R.solve_waveguide_vec(mesh, WL, mg.FIBER_IOR, Nmax=Nmax, AB=(A, B), return_info=True)
    t3 = time.perf_counter()
    order = np.argsort(-w.real, kind="stable")
    print(name, "restated: Ne", mesh.num_elements, "N", A.shape[0], "neff", np.sqrt(w.real[order]) / k, "count", cnt,
          "edges %.1f s, assembly %.1f s, solve %.1f s, n_op %d" % (t1 - t0, t2 - t1, t3 - t2, info["n_op"]), flush=True)
    save(name, generator="fiber_disc", target=target, seed=seed, order=1, Nmax=Nmax, wl=WL, elements=mesh.num_elements,
         N=A.shape[0], nnzA=A.nnz, nnzB=B.nnz, w=w.real[order], w_imag_max=np.abs(w.imag).max(), count=cnt,
         seconds=t3 - t0, seconds_edges=t1 - t0, seconds_assembly=t2 - t1, seconds_solve=t3 - t2, n_op=info["n_op"],
         nnz_lu=info["nnz_lu"], cores=1, source="oracle/restated.py solve_waveguide_vec", sha_edges=digest(np.asarray(mesh.cells[0].data).astype(np.int64)),
         **dg)


if __name__ == "__main__":
    what = sys.argv[1:] or ["c3", "c2", "c4"]
    if "c3" in what:
        c3()
    if "c2" in what:
        vec("c2_vec_50k", 50_000, 22)
    if "c4" in what:
        vec("c4_vec_1M", 1_000_000, 0)
