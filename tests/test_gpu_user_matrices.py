"""GPU: solve_sparse (fe_solver.py:262-290) -- the eigen-solve on matrices the caller brings.  The
matrices are the ones the LITERAL reference assembled (tests/golden), the expected eigenpairs the ones
its own solve_sparse returned for them (tests/golden/solve_sparse.npz)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from tests import util
from tests.test_gpu_eig import NEFF_TOL, check_fields

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

K = 2 * np.pi / util.WL


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(util.GOLD, "solve_sparse.npz"), allow_pickle=False)


@pytest.mark.parametrize("tag,name", [("p2", "scalar_disc600"), ("p1", "scalar_p1_disc800")])
def test_solve_sparse_on_reference_matrices(ws, gold, tag, name):
    mesh, ior, z = util.load_golden(name)
    A, B = util.golden_matrix(z, "A"), util.golden_matrix(z, "B")
    A0, B0 = A.copy(), B.copy()
    w, v, cnt = ws.solve_sparse(A, B, mesh, K, ior)                   # default: six modes
    wr, vr = gold[tag + "_w"], gold[tag + "_v"]
    assert w.shape == wr.shape and v.shape == vr.shape and w.dtype == np.float64
    assert cnt == int(gold[tag + "_count"])
    assert np.all(np.diff(w) <= 0)
    assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0)
    Bc = B.tocsr()
    assert np.allclose(np.einsum("ij,ij->i", v, (Bc @ v.T).T), 1.0, atol=1e-10)   # B-orthonormal like eigsh
    check_fields(w, v, wr, vr, Bc)
    # the caller's matrices are not touched
    for M, M0 in ((A, A0), (B, B0)):
        assert np.array_equal(M.indptr, M0.indptr) and np.array_equal(M.indices, M0.indices)
        assert np.array_equal(M.data, M0.data)


def test_formats_pruned_zeros_and_own_matrices(ws, gold):
    mesh, ior, z = util.load_golden("scalar_disc600")
    A, B = util.golden_matrix(z, "A"), util.golden_matrix(z, "B")
    wr = gold["p2_w"]
    # CSC / COO input, explicit zeros pruned (entries then sit on a sub-pattern of the mesh pattern)
    Ap, Bp = A.copy(), B.copy()
    Ap.eliminate_zeros()
    Bp.eliminate_zeros()
    assert Bp.nnz < B.nnz
    for Ax, Bx in ((A.tocsc(), B.tocoo()), (Ap, Bp)):
        w, v, cnt = ws.solve_sparse(Ax, Bx, mesh, K, ior, num_modes=6)
        assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0) and cnt == int(gold["p2_count"])
    # matrices assembled by this package, more modes: same as solve_waveguide on the mesh
    A2, B2 = ws.construct_AB(mesh, ior, K, sparse=True)
    w10, v10, c10 = ws.solve_sparse(A2, B2, mesh, K, ior, num_modes=10)
    assert np.allclose(util.neff(w10), util.neff(z["w"]), rtol=NEFF_TOL, atol=0) and c10 == int(z["count"])


def test_misuse_is_refused(ws):
    mesh, ior, z = util.load_golden("scalar_disc600")
    A, B = util.golden_matrix(z, "A"), util.golden_matrix(z, "B")
    N = A.shape[0]
    with pytest.raises(ValueError, match="shape"):
        ws.solve_sparse(A[:-1, :-1], B, mesh, K, ior)
    with pytest.raises(TypeError):
        ws.solve_sparse(np.eye(N), B, mesh, K, ior)
    # an entry between two unknowns that share no element is not part of this mesh's problem
    dense_row = np.zeros(N, dtype=bool)
    dense_row[A[0].indices] = True
    j = int(np.nonzero(~dense_row)[0][-1])
    X = A.tolil()
    X[0, j] = 1.0
    with pytest.raises(ValueError, match="outside the sparsity pattern"):
        ws.solve_sparse(X.tocsr(), B, mesh, K, ior)
    with pytest.raises(TypeError):
        ws.solve_sparse(A.astype(np.complex128), B, mesh, K, ior)
    assert sp.issparse(A)


def test_legacy_dense_solve_all_eigenpairs(ws):
    """fe_solver.py:220-260: solve(A, B, mesh, k, IOR_dict) = scipy.linalg.eigh(A, B) for ALL eigenpairs, descending."""
    import scipy.linalg as sla
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.fiber_disc(300, seed=4)
    k = 2 * np.pi / 1.55
    A, B = R.construct_AB_order2(mesh, mg.FIBER_IOR, k)
    Ad, Bd = A.toarray(), B.toarray()
    w, v, cnt = ws.solve(Ad.copy(), Bd.copy(), mesh, k, mg.FIBER_IOR)
    wr, vr = sla.eigh(Ad, Bd)
    wr, vr = wr[::-1], vr.T[::-1]
    assert w.shape == wr.shape and v.shape == vr.shape
    assert np.allclose(w, wr, rtol=1e-9, atol=1e-9 * abs(wr).max())
    assert cnt == R.count_modes_scalar(wr, k, mg.FIBER_IOR)
    G = v @ Bd @ v.T
    assert np.allclose(G, np.eye(len(w)), atol=1e-8)                 # B-orthonormal like scipy's eigh
    for i in range(3):                                               # the guided modes, up to sign
        assert abs(v[i] @ Bd @ vr[i]) >= 1 - 1e-8 or abs(w[i] - w[i + 1]) < 1e-6 * abs(w[i]) or abs(w[i] - w[i - 1]) < 1e-6 * abs(w[i])
    w2, v2, c2 = ws.solve(A, B, mesh, k, mg.FIBER_IOR)              # sparse inputs are densified
    assert np.allclose(w2, w)


def test_real_output_option(ws):
    mg = ws.meshgen
    mesh = mg.fiber_disc(3000, order=1, seed=5)
    ws.get_unique_edges(mesh)
    w, v, c = ws.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False)
    wr, vr, cr = ws.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False, real_output=True)
    assert w.dtype == np.complex128 and wr.dtype == np.float64 and vr.dtype == np.float64 and c == cr
    assert np.array_equal(w.real, wr) and np.array_equal(v.real, vr)
