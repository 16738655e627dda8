"""CPU, build container only: the oracle's restatement against the LITERAL reference run live (imported
unmodified from /root/reference behind the stub modules of oracle/refshim.py) on randomly seeded small
meshes -- beyond the fixed fixtures of tests/golden.  Skipped where the reference tree is absent (GPU box)."""
import copy

import numpy as np
import pytest

from oracle import refshim, restated as R
from tests import util

pytestmark = pytest.mark.skipif(not refshim.available(), reason="literal reference not present (GPU box)")

K = 2 * np.pi / util.WL


@pytest.fixture(scope="module")
def ref():
    return refshim.load()


def _same_sparse(M, Rm, exact=True, tol=0.0):
    M, Rm = M.tocsr().copy(), Rm.tocsr().copy()
    M.sort_indices()
    Rm.sort_indices()
    assert M.shape == Rm.shape
    assert np.array_equal(M.indptr, Rm.indptr) and np.array_equal(M.indices, Rm.indices)
    if exact:
        assert np.array_equal(M.data, Rm.data)
    else:
        assert np.all(np.abs(M.data - Rm.data) <= tol * np.abs(Rm.data).max())


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_scalar_assembly_and_solve_random_meshes(ref, seed):
    import wavesolve_b200.meshgen as mg
    fe, sf, wg = ref
    rng = np.random.default_rng(seed)
    mesh = mg.fiber_disc(int(rng.integers(150, 500)), seed=seed)
    ior = {"core": 1.444 + float(rng.uniform(5e-3, 2e-2)), "cladding": 1.444}
    A, B = R.construct_AB_order2(mesh, ior, K)
    Ar, Br = fe.construct_AB(mesh, ior, K, sparse=True, order=2)
    _same_sparse(A, Ar)
    _same_sparse(B, Br)
    _same_sparse(R.construct_B_order2(mesh), fe.construct_B(mesh, sparse=True))
    w, v, cnt = R.solve_waveguide(mesh, util.WL, ior, Nmax=5)
    wr, vr, cntr = fe.solve_waveguide(mesh, util.WL, ior, sparse=True, Nmax=5)
    assert cnt == cntr
    assert np.allclose(util.neff(w), util.neff(np.asarray(wr)), rtol=1e-10, atol=0)
    assert np.array_equal(R.compute_IOR_arr(mesh, ior), fe.compute_IOR_arr(mesh, ior))


@pytest.mark.parametrize("seed", [201, 202])
def test_edges_and_vector_assembly_random_meshes(ref, seed):
    import wavesolve_b200.meshgen as mg
    fe, sf, wg = ref
    rng = np.random.default_rng(seed)
    mesh = mg.fiber_disc(int(rng.integers(120, 260)), order=1, seed=seed)
    m_ref, m_ours = copy.deepcopy(mesh), copy.deepcopy(mesh)
    er, ir = wg.get_unique_edges(m_ref)
    e, i = R.get_unique_edges(m_ours)
    assert np.array_equal(np.asarray(e), np.asarray(er)) and np.array_equal(np.asarray(i), np.asarray(ir))
    assert m_ours.num_edges == m_ref.num_edges and m_ours.num_points == m_ref.num_points
    ior = {"core": 1.444 + float(rng.uniform(5e-3, 2e-2)), "cladding": 1.444}
    A, B = R.construct_AB_vec(m_ours, ior, K)
    Ar, Br = fe.construct_AB_vec(m_ref, ior, K, sparse=True)
    # identical patterns (pruned zeros included); values to the last bit except where the literal code
    # squares lengths with libm pow
    _same_sparse(A, Ar, exact=False, tol=1e-15)
    _same_sparse(B, Br, exact=False, tol=1e-15)


def test_ragged_material_sets_live(ref):
    """Elements in no set, an empty set, one element in two sets, sets out of order."""
    import wavesolve_b200.meshgen as mg
    fe, sf, wg = ref
    m = mg.structured_square(7)
    core, clad = m.cell_sets["core"][1], m.cell_sets["cladding"][1]
    m.cell_sets = {"cladding": [None, np.concatenate([clad[3:-4][::-1], core[:2]]), None],
                   "void": [None, np.zeros(0, dtype=np.int64), None],
                   "core": [None, core, None]}
    ior = {"core": 1.46, "void": 1.0, "cladding": 1.444}
    A, B = R.construct_AB_order2(m, ior, K)
    Ar, Br = fe.construct_AB(m, ior, K, sparse=True, order=2)
    _same_sparse(A, Ar)
    _same_sparse(B, Br)
    assert np.array_equal(R.compute_IOR_arr(m, ior), fe.compute_IOR_arr(m, ior))
