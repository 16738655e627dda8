"""CPU: the host symbolic phase of the sparse solver (csrc/ws_symbolic.cu) -- structural
invariants, and a numpy multifrontal factorisation (oracle/mf_model.py) run on exactly the
elimination tree the library produced must solve the shift-invert systems."""
import copy

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mf_model as MF
from oracle import restated as R
from wavesolve_b200 import meshgen as mg
from wavesolve_b200 import solver as sv

K0 = 2 * np.pi / 1.55


def scalar_setup(ne, seed=7):
    mesh = mg.fiber_disc(ne, order=2, seed=seed)
    conn, _, _ = R.material_major(mesh)
    cxy = np.ascontiguousarray(mesh.points[:, :2])
    return mesh, conn.astype(np.int32), cxy, int(conn.max()) + 1


def vector_setup(ne, seed=7, gen="disc"):
    mesh = mg.fiber_disc(ne, order=1, seed=seed) if gen == "disc" else mg.structured_square(int(np.sqrt(ne / 2)), order=1)
    R.get_unique_edges(mesh)
    conn, _, _ = R.material_major(mesh)
    e, _, _ = R.material_major(mesh, np.asarray(mesh.edge_indices))
    Ntt = len(mesh.cells[0].data)
    dofs = np.hstack([e.astype(np.int64), conn.astype(np.int64) + Ntt]).astype(np.int32)
    cxy = np.ascontiguousarray(mesh.points[:, :2])
    return mesh, dofs, cxy, Ntt, Ntt + len(mesh.points)


def check_structure(S, ex, dofs, N, cls_start):
    perm, p, q, start, bndoff, bnd, rel = (ex[k] for k in ("perm", "p", "q", "start", "bndoff", "bnd", "rel"))
    assert np.array_equal(np.sort(perm), np.arange(N))
    iperm = np.empty(N, dtype=np.int64)
    iperm[perm] = np.arange(N)
    nf = S.nfronts
    assert p[1:].sum() == N and p[0] == 0
    # own ranges tile [0, N) in level-major bottom-up order
    order = [t for l in range(S.depth, -1, -1) for t in range(1 << l, 1 << (l + 1))]
    pos = 0
    for t in order:
        assert start[t] == pos
        pos += p[t]
    # inside a front: first-class unknowns first, each class ascending in old index
    for t in order[:: max(1, len(order) // 50)]:
        old = perm[start[t]:start[t] + p[t]].astype(np.int64)
        cls = (old < cls_start).astype(int)
        assert np.all(np.diff(cls) >= 0)
        for c in (0, 1):
            assert np.all(np.diff(old[cls == c]) > 0)
    # every element's unknowns live in its leaf front; coupled unknowns share a front chain
    owner = np.empty(N, dtype=np.int64)
    for t in range(1, nf):
        owner[start[t]:start[t] + p[t]] = t
    new = iperm[dofs.astype(np.int64)]
    own_e = owner[new]                                   # (Ne, 6) owners
    deepest = own_e.max(axis=1)                           # heap id grows with depth
    for a in range(6):
        o = own_e[:, a]
        # o must be an ancestor-or-self of `deepest`
        d = deepest.copy()
        while True:
            m = d > o
            if not m.any():
                break
            d[m] >>= 1
        assert np.array_equal(d, o)
    # boundary lists sorted, above own range, rel consistent with the parent's front
    for t in range(2, nf):
        b = bnd[bndoff[t]:bndoff[t + 1]]
        assert np.all(np.diff(b) > 0)
        if len(b):
            assert b[0] >= start[t] + p[t]
        par = t // 2
        front = np.concatenate([np.arange(start[par], start[par] + p[par]), bnd[bndoff[par]:bndoff[par + 1]]])
        assert np.array_equal(front[rel[bndoff[t]:bndoff[t + 1]]], b)
    assert q[1] == 0


@pytest.mark.parametrize("ne,leaf", [(700, 16), (5000, 48), (20000, 32)])
def test_symbolic_invariants_scalar(ne, leaf):
    mesh, dofs, cxy, N = scalar_setup(ne)
    S = sv.Symbolic(dofs, cxy, N, 0, leaf)
    check_structure(S, S.export(), dofs, N, 0)
    assert S.nnzL > N and S.maxp > 0


def test_symbolic_invariants_vector_nodes_first():
    mesh, dofs, cxy, Ntt, N = vector_setup(6000)
    S = sv.Symbolic(dofs, cxy, N, Ntt, 32)
    check_structure(S, S.export(), dofs, N, Ntt)


def test_symbolic_rejects_empty_rows():
    mesh, dofs, cxy, N = scalar_setup(700)
    from wavesolve_b200 import _lib
    with pytest.raises(_lib.WavesolveB200Error, match="singular"):
        sv.Symbolic(dofs, cxy, N + 3, 0, 16)


def _model_from_lib(S):
    ex = S.export()
    return MF.structure_from_arrays(S.N, S.depth, ex["perm"], ex["p"], ex["start"], ex["bndoff"], ex["bnd"])


def test_model_factorisation_on_library_tree_scalar():
    mesh, dofs, cxy, N = scalar_setup(5000)
    A, B = R.construct_AB_order2(mesh, mg.FIBER_IOR, K0)
    sigma = (K0 * max(mg.FIBER_IOR.values())) ** 2
    M = (A - sigma * B).tocsr()
    S = sv.Symbolic(dofs, cxy, N, 0, 32)
    st = MF.factor(_model_from_lib(S), M)
    b = np.random.default_rng(0).standard_normal(N)
    x = MF.solve(st, b)
    assert np.linalg.norm(M @ x - b) <= 1e-12 * np.linalg.norm(b)
    assert (st.d < 0).all()                      # sigma*B - A is SPD at the default shift


@pytest.mark.parametrize("gen", ["disc", "sq"])
def test_model_factorisation_on_library_tree_vector_quasidefinite(gen):
    """The vector matrix is factorised after the congruence T = [[I,-G],[0,I]] (G = discrete
    gradient), which makes it quasi-definite; nodes are eliminated before edges in every front."""
    mesh, dofs, cxy, Ntt, N = vector_setup(5000, gen=gen)
    A, B = R.construct_AB_vec(mesh, mg.FIBER_IOR, K0)
    sigma = (K0 * max(mg.FIBER_IOR.values())) ** 2
    M = (A - sigma * B).tocsr()
    edges = np.asarray(mesh.cells[0].data).astype(np.int64)
    Nzz = N - Ntt
    L = np.linalg.norm(mesh.points[edges[:, 1], :2] - mesh.points[edges[:, 0], :2], axis=1)
    G = sp.csr_matrix((np.r_[1 / L, -1 / L], (np.r_[np.arange(Ntt), np.arange(Ntt)], np.r_[edges[:, 1], edges[:, 0]])),
                      shape=(Ntt, Nzz))
    T = sp.bmat([[sp.identity(Ntt), -G], [None, sp.identity(Nzz)]]).tocsr()
    Mp = (T.T @ M @ T).tocsr()
    S = sv.Symbolic(dofs, cxy, N, Ntt, 32)
    st = MF.factor(_model_from_lib(S), Mp)
    b = np.random.default_rng(1).standard_normal(N)
    x = T @ MF.solve(st, T.T @ b)
    assert np.linalg.norm(M @ x - b) <= 1e-11 * np.linalg.norm(b)
    # inertia = (Ntt negative, Nzz positive): no eigenvalue above the default shift
    assert (st.d < 0).sum() == Ntt and (st.d > 0).sum() == Nzz
