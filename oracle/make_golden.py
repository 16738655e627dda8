"""TEST INFRASTRUCTURE ONLY.  Generate ``tests/golden/*.npz`` from the LITERAL reference.

Run in the build container (where ``/root/reference`` exists):

    python -m oracle.make_golden

Every fixture stores the input mesh arrays (so nothing depends on the mesh generator or on
qhull being bit-reproducible elsewhere) and the outputs of the unmodified reference
functions ``construct_AB`` (fe_solver.py:131), ``construct_B`` (fe_solver.py:101),
``get_unique_edges`` (waveguide.py:10), ``construct_AB_vec`` (fe_solver.py:148),
``solve_waveguide`` (fe_solver.py:292) and ``solve_waveguide_vec`` (fe_solver.py:350,
sparse_solve_mode='transform').  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so these reference-generated outputs are what pins the oracle.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from wavesolve_b200 import meshgen as mg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WL = 1.55


def mesh_arrays(mesh):
    labels = list(mesh.cell_sets.keys())
    d = {"points": mesh.points, "tris": np.asarray(mesh.cells[1].data),
         "labels": np.array(labels)}
    for i, lab in enumerate(labels):
        d["set_%d" % i] = np.asarray(mesh.cell_sets[lab][1])
    return d


def csx(prefix, M):
    M = M.copy()
    M.sort_indices()
    return {prefix + "_indptr": M.indptr, prefix + "_indices": M.indices, prefix + "_data": M.data,
            prefix + "_shape": np.array(M.shape), prefix + "_format": np.array(M.format)}


def scalar_case(fe, name, mesh, ior, Nmax=10, target_neff=None, store_v=True, with_B=False):
    k = 2 * np.pi / WL
    out = mesh_arrays(mesh)
    out["ior"] = np.array([ior[l] for l in mesh.cell_sets.keys()])
    out["wl"] = np.array(WL)
    A, B = fe.construct_AB(mesh, ior, k, sparse=True, order=2)
    out.update(csx("A", A))
    out.update(csx("B", B))
    if with_B:
        out.update(csx("Bfull", fe.construct_B(mesh, sparse=True)))
    if Nmax:
        w, v, cnt = fe.solve_waveguide(mesh, WL, ior, sparse=True, Nmax=Nmax, order=2,
                                       target_neff=target_neff)
        out["w"] = np.array(w)
        out["count"] = np.array(cnt)
        out["Nmax"] = np.array(Nmax)
        out["target_neff"] = np.array(np.nan if target_neff is None else target_neff)
        if store_v:
            out["v"] = np.ascontiguousarray(v)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "Ne", mesh.num_elements, "N", A.shape[0], "nnz", A.nnz,
          "neff", (np.sqrt(out["w"]) / k)[:4] if Nmax else None, "count", out.get("count"))


def scalar_p1_case(fe, name, mesh, ior, Nmax=6):
    """order=1 scalar path: construct_AB_order1 (fe_solver.py:103-129, lil matrices) and
    solve_waveguide(order=1) (fe_solver.py:292-348)."""
    k = 2 * np.pi / WL
    out = mesh_arrays(mesh)
    out["ior"] = np.array([ior[l] for l in mesh.cell_sets.keys()])
    out["wl"] = np.array(WL)
    A, B = fe.construct_AB(mesh, ior, k, sparse=True, order=1)
    assert A.format == "lil"
    out.update(csx("A", A.tocsr()))
    out.update(csx("B", B.tocsr()))
    w, v, cnt = fe.solve_waveguide(mesh, WL, ior, sparse=True, Nmax=Nmax, order=1)
    out["w"] = np.array(w)
    out["v"] = np.ascontiguousarray(v)
    out["count"] = np.array(cnt)
    out["Nmax"] = np.array(Nmax)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "Ne", mesh.num_elements, "N", A.shape[0], "nnz", A.nnz, "neff", np.sqrt(out["w"]) / k, "count", cnt)


def vector_case(fe, wg, name, mesh, ior, Nmax=6, target_neff=None, solve=True):
    k = 2 * np.pi / WL
    out = mesh_arrays(mesh)
    out["ior"] = np.array([ior[l] for l in mesh.cell_sets.keys()])
    out["wl"] = np.array(WL)
    m = copy.deepcopy(mesh)
    edges, eidx = wg.get_unique_edges(m)
    out["edges"] = edges
    out["edge_indices"] = eidx
    out["edge_flips"] = m.edge_flips
    A, B = fe.construct_AB_vec(m, ior, k, sparse=True)
    out.update(csx("A", A))
    out.update(csx("B", B))
    if solve:
        w, v, cnt = fe.solve_waveguide_vec(m, WL, ior, sparse=True, Nmax=Nmax,
                                           target_neff=target_neff,
                                           sparse_solve_mode="transform", verbose=False)
        out["w"] = np.array(w)
        out["v"] = np.ascontiguousarray(v)
        out["count"] = np.array(cnt)
        out["Nmax"] = np.array(Nmax)
        out["target_neff"] = np.array(np.nan if target_neff is None else target_neff)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "Ne", mesh.num_elements, "N", A.shape[0], "nnzA", A.nnz, "nnzB", B.nnz,
          "neff", np.sqrt(out["w"].real) / k if solve else None, "count", out.get("count"))


def interp_cases(fe, wg):
    """Mode resampling (fe_solver.py:594-872, mesher.py:235): KD-tree point location, brute-force
    point location, P2 / P1 / Whitney weights, interpolate, interpolate_vec, mesh_interpolate."""
    import importlib
    ms = importlib.import_module("wavesolve.mesher")
    rng = np.random.default_rng(7)
    # ---- scalar P2
    mesh = mg.fiber_disc(600, seed=1)
    out = mesh_arrays(mesh)
    x, y = mesh.points[:, 0], mesh.points[:, 1]
    v = np.sin(0.3 * x) * np.cos(0.2 * y) + 0.1 * x - 0.05 * y * y
    vc = v * (1.0 + 0.5j) + 0.2j * x
    xa, ya = np.linspace(-10.5, 10.5, 23), np.linspace(-9.7, 10.9, 19)
    tree = ms.construct_meshtree(mesh)
    out.update(v=v, vc=vc, xa=xa, ya=ya)
    # (without maxr the reference raises IndexError for a point outside a P2 mesh: it asks the tree for
    #  up to Npoints-1 > Ne neighbours, fe_solver.py:853-857; its callers always pass maxr)
    out["tri_idxs"] = fe.get_tri_idxs_KDtree(mesh, tree, xa, ya, maxr=9.9)
    out["tri_idxs_maxr"] = fe.get_tri_idxs_KDtree(mesh, tree, xa, ya, maxr=7.5)
    out["tri_idxs_brute"] = fe.get_tri_idxs(mesh, xa[::2], ya[::2])
    out["weights"] = fe.get_interp_weights(mesh, xa, ya, out["tri_idxs"], order=2)
    out["interp"] = fe.interpolate(v, mesh, xa, ya, meshtree=tree, order=2, maxr=9.9)
    out["interp_c"] = fe.interpolate(vc, mesh, xa, ya, tri_idxs=out["tri_idxs"], interp_weights=out["weights"])
    out["interp_maxr"] = fe.interpolate(v, mesh, xa, ya, meshtree=tree, order=2, maxr=7.5)
    outmesh = mg.fiber_disc(150, r_clad=9.0, seed=9)
    out["out_points"] = outmesh.points
    out["mesh_interp"] = fe.mesh_interpolate(v, mesh, outmesh, tree, max_tries=40)
    out["point_interp"] = np.array([fe.unstructured_interpolate(v, mesh, p, tree, max_tries=40)
                                    for p in [[0.3, -0.2], [4.99, 0.01], [9.2, 3.0]]])
    np.savez_compressed(os.path.join(GOLD, "interp_p2_disc600.npz"), **out)
    print("interp_p2_disc600", "inside", int((out["tri_idxs"] >= 0).sum()), "of", out["tri_idxs"].size,
          "nan in mesh_interp", int(np.isnan(out["mesh_interp"]).sum()))
    # ---- vector (Whitney edge functions) and P1 weights
    mesh = mg.fiber_disc(500, order=1, seed=2)
    m = copy.deepcopy(mesh)
    edges, eidx = wg.get_unique_edges(m)
    out = mesh_arrays(mesh)
    out["edges"], out["edge_indices"] = edges, eidx
    N = len(edges) + len(mesh.points)
    vv = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xa, ya = np.linspace(-9.3, 9.1, 17), np.linspace(-10.2, 8.8, 15)
    tree = ms.construct_meshtree(m)
    out.update(v=vv, xa=xa, ya=ya)
    out["tri_idxs"] = fe.get_tri_idxs_KDtree(m, tree, xa, ya)
    out["weights_vec"] = fe.get_interp_weights_vec(m, xa, ya, out["tri_idxs"])
    out["weights_p1"] = fe.get_interp_weights(m, xa, ya, out["tri_idxs"], order=1)
    out["interp_vec"] = fe.interpolate_vec(vv, m, xa, ya, meshtree=tree)
    out["interp_vec_real"] = fe.interpolate_vec(vv.real, m, xa, ya, tri_idxs=out["tri_idxs"],
                                                interp_weights=out["weights_vec"])
    vz = rng.standard_normal(len(mesh.points))
    out["vz"] = vz
    out["interp_p1"] = fe.interpolate(vz, m, xa, ya, meshtree=tree, order=1)
    np.savez_compressed(os.path.join(GOLD, "interp_vec_disc500.npz"), **out)
    print("interp_vec_disc500", "inside", int((out["tri_idxs"] >= 0).sum()), "of", out["tri_idxs"].size)


def fake_meshio_mesh():
    """What meshio.read returns for a gmsh .msh with two triangle6 blocks: one index array per cell block
    in every cell set, the last set being gmsh's bounding entities (waveguide.py:91-93)."""
    rng = np.random.default_rng(12)
    pts = rng.standard_normal((40, 3))
    blocks = [mg.CellBlock("line", rng.integers(0, 40, (7, 2)).astype(np.uint64)),
              mg.CellBlock("triangle6", rng.integers(0, 40, (5, 6)).astype(np.uint64)),
              mg.CellBlock("triangle6", rng.integers(0, 40, (4, 6)).astype(np.uint64)),
              mg.CellBlock("vertex", rng.integers(0, 40, (3, 1)).astype(np.uint64))]
    e = np.zeros(0, dtype=np.int64)
    sets = {"core": [e, np.array([0, 2]), np.array([1]), e],
            "cladding": [e, np.array([1, 3, 4]), np.array([0, 2, 3]), e],
            "jacket": [e, e, e, e],
            "ring": [e, e, np.array([3]), e],
            "gmsh:bounding_entities": [np.array([0, 1]), e, e, np.array([0])]}
    m = mg.Mesh(pts, blocks[1].data, sets)
    m.cells = blocks
    return m


def meshio_case(wg):
    """load_meshio_mesh (waveguide.py:91-126) run literally, with meshio.read returning the fake above."""
    wg.meshio.read = lambda name: fake_meshio_mesh()
    out = wg.load_meshio_mesh("whatever")
    d = {"tris": np.asarray(out.cells[1].data), "labels": np.array(list(out.cell_sets.keys())),
         "others_none": np.array([out.cells[i] is None for i in range(len(out.cells)) if i != 1])}
    for i, (k, v) in enumerate(out.cell_sets.items()):
        assert v[0] is None and v[2] is None and len(v) == 3
        d["set_%d" % i] = np.asarray(v[1], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "meshio_regroup.npz"), **d)
    print("meshio_regroup", d["tris"].shape, [len(d["set_%d" % i]) for i in range(len(d["labels"]))])


def solve_sparse_case(fe):
    """solve_sparse (fe_solver.py:262-290) with its default num_modes=6 on the matrices of two
    existing fixtures (P2 disc, P1 disc: lil -> csr of construct_AB_order1), and compute_IOR_arr
    (fe_solver.py:446-452) on the ragged material sets.  Only results are stored: the meshes and
    matrices are the ones already in scalar_disc600 / scalar_p1_disc800 / scalar_sq6_ragged."""
    k = 2 * np.pi / WL
    ior = mg.FIBER_IOR
    out = {}
    for tag, mesh, order in (("p2", mg.fiber_disc(600, seed=1), 2), ("p1", mg.fiber_disc(800, order=1, seed=6), 1)):
        A, B = fe.construct_AB(mesh, ior, k, sparse=True, order=order)
        w, v, cnt = fe.solve_sparse(A, B, mesh, k, ior)
        out[tag + "_w"], out[tag + "_v"], out[tag + "_count"] = np.array(w), np.ascontiguousarray(v), np.array(cnt)
        print("solve_sparse", tag, "neff", np.sqrt(np.array(w)) / k, "count", cnt)
    m = mg.structured_square(6)
    core, clad = m.cell_sets["core"][1], m.cell_sets["cladding"][1]
    m.cell_sets = {"core": [None, core, None], "void": [None, np.zeros(0, dtype=np.int64), None],
                   "cladding": [None, np.concatenate([clad[:-5], core[:1]]), None]}
    out["ior_arr_ragged"] = fe.compute_IOR_arr(m, {"core": 1.46, "void": 1.0, "cladding": 1.444})
    np.savez_compressed(os.path.join(GOLD, "solve_sparse.npz"), **out)


def main():
    fe, sf, wg = refshim.load()
    if "--solve-sparse-only" in sys.argv:   # added later; leaves the earlier fixtures untouched
        solve_sparse_case(fe)
        return
    if "--meshio-only" in sys.argv:      # added later; leaves the earlier fixtures untouched
        meshio_case(wg)
        return
    os.makedirs(GOLD, exist_ok=True)
    ior = mg.FIBER_IOR
    if "--interp-only" in sys.argv:      # added later; leaves the earlier fixtures untouched
        interp_cases(fe, wg)
        return
    if "--p1-only" in sys.argv:          # added later; leaves the earlier fixtures untouched
        scalar_p1_case(fe, "scalar_p1_sq8", mg.structured_square(8, order=1), ior, Nmax=4)
        scalar_p1_case(fe, "scalar_p1_disc800", mg.fiber_disc(800, order=1, seed=6), ior, Nmax=6)
        return
    # ---- scalar P2 ----
    scalar_case(fe, "scalar_sq8", mg.structured_square(8), ior, with_B=True)
    scalar_case(fe, "scalar_disc600", mg.fiber_disc(600, seed=1), ior)
    scalar_case(fe, "scalar_disc600_target", mg.fiber_disc(600, seed=1), ior, Nmax=4,
                target_neff=1.447)
    # uint64 connectivity as pygmsh/gmsh delivers it
    scalar_case(fe, "scalar_sq5_uint64", mg.structured_square(5, dtype=np.uint64), ior, Nmax=3)
    # ragged: one material empty, some elements in no set, one element in two sets
    m = mg.structured_square(6)
    cs = m.cell_sets
    core = cs["core"][1]
    clad = cs["cladding"][1]
    m.cell_sets = {"core": [None, core, None],
                   "void": [None, np.zeros(0, dtype=np.int64), None],
                   "cladding": [None, np.concatenate([clad[:-5], core[:1]]), None]}
    scalar_case(fe, "scalar_sq6_ragged", m, {"core": 1.46, "void": 1.0, "cladding": 1.444}, Nmax=0)
    # config C1: ~5k elements, 10 modes (eigenvalues + count only, vectors are large)
    scalar_case(fe, "scalar_c1_disc5k", mg.fiber_disc(5000, seed=0), ior, store_v=False)
    # lantern-like three-material cross section
    scalar_case(fe, "scalar_lantern2k", mg.lantern_cross_section(2000, seed=3), mg.LANTERN_IOR, Nmax=8,
                store_v=False)
    # ---- vector P1 + Whitney ----
    vector_case(fe, wg, "vec_sq6", mg.structured_square(6, order=1), ior, Nmax=6)
    vector_case(fe, wg, "vec_sq12", mg.structured_square(12, order=1), ior, Nmax=6)
    vector_case(fe, wg, "vec_disc500", mg.fiber_disc(500, order=1, seed=2), ior, Nmax=6)
    vector_case(fe, wg, "vec_disc1500", mg.fiber_disc(1500, order=1, seed=4), ior, Nmax=10)
    vector_case(fe, wg, "vec_disc500_target", mg.fiber_disc(500, order=1, seed=2), ior, Nmax=4,
                target_neff=1.448)
    m = mg.structured_square(5, order=1, dtype=np.uint64)
    vector_case(fe, wg, "vec_sq5_uint64", m, ior, solve=False)
    # ---- scalar P1 ----
    scalar_p1_case(fe, "scalar_p1_sq8", mg.structured_square(8, order=1), ior, Nmax=4)
    scalar_p1_case(fe, "scalar_p1_disc800", mg.fiber_disc(800, order=1, seed=6), ior, Nmax=6)


if __name__ == "__main__":
    main()
