"""CPU: solve_sparse (fe_solver.py:262-290) and compute_IOR_arr (fe_solver.py:446-452) -- the oracle's
restatement and the product's host helper against tests/golden/solve_sparse.npz, which
`oracle/make_golden.py --solve-sparse-only` wrote with the LITERAL reference functions."""
import os

import numpy as np
import pytest

from oracle import restated as R
from tests import util

K = 2 * np.pi / util.WL


def _gold():
    return np.load(os.path.join(util.GOLD, "solve_sparse.npz"), allow_pickle=False)


def _ragged_mesh():
    import wavesolve_b200 as ws
    m = ws.meshgen.structured_square(6)
    core, clad = m.cell_sets["core"][1], m.cell_sets["cladding"][1]
    m.cell_sets = {"core": [None, core, None], "void": [None, np.zeros(0, dtype=np.int64), None],
                   "cladding": [None, np.concatenate([clad[:-5], core[:1]]), None]}
    return m, {"core": 1.46, "void": 1.0, "cladding": 1.444}


def test_oracle_solve_sparse_matches_literal_reference():
    g = _gold()
    for tag, name in (("p2", "scalar_disc600"), ("p1", "scalar_p1_disc800")):
        mesh, ior, z = util.load_golden(name)
        A, B = util.golden_matrix(z, "A"), util.golden_matrix(z, "B")
        w, v, cnt = R.solve_sparse(A, B, mesh, K, ior)
        assert w.shape == (6,) and v.shape == g[tag + "_v"].shape
        assert cnt == int(g[tag + "_count"])
        assert np.allclose(util.neff(w), util.neff(g[tag + "_w"]), rtol=1e-10, atol=0)
        # the fixture was made on the same mesh as the solve_waveguide fixture: its six modes are the
        # first six of the ten (P2) / the six (P1) stored there
        assert np.allclose(util.neff(g[tag + "_w"]), util.neff(z["w"][:6]), rtol=1e-10, atol=0)


def test_compute_IOR_arr_matches_literal_reference():
    import wavesolve_b200 as ws
    g = _gold()
    mesh, ior = _ragged_mesh()
    ref = g["ior_arr_ragged"]
    assert np.array_equal(ws.compute_IOR_arr(mesh, ior), ref)
    assert np.array_equal(R.compute_IOR_arr(mesh, ior), ref)
    assert (ref == 0).sum() == 5           # the five cladding elements left out of every set stay 0
    assert ref[mesh.cell_sets["core"][1][0]] == 1.444     # in two sets: the later set wins


def test_place_on_pattern_index_marshalling():
    """The placement of caller-supplied CSR values on a (larger) pattern -- pure index work in torch,
    device-agnostic, so it is checked here on CPU tensors against scipy."""
    import scipy.sparse as sp
    import torch
    from wavesolve_b200.fe_solver import _place_on_pattern
    rng = np.random.default_rng(4)
    N = 37
    P = sp.random(N, N, density=0.2, random_state=5, format="csr")
    P.data[:] = 1.0
    P = (P + P.T + sp.eye(N)).tocsr()
    P.sort_indices()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
    pp, pi = t(P.indptr, np.int32), t(P.indices, np.int32)
    # identical structure: the data array itself comes back
    d = t(rng.standard_normal(P.nnz), np.float64)
    assert _place_on_pattern(pp, pi, pp.clone(), pi.clone(), d, N) is d
    # a sub-pattern (some entries dropped, incl. whole rows): values land in their slots, the rest is 0
    M = P.copy()
    M.data = rng.standard_normal(P.nnz)
    M.data[rng.random(P.nnz) < 0.4] = 0.0
    M.data[P.indptr[3]:P.indptr[5]] = 0.0
    full = M.data.copy()
    M.eliminate_zeros()
    out = _place_on_pattern(pp, pi, t(M.indptr, np.int32), t(M.indices, np.int32), t(M.data, np.float64), N)
    assert np.array_equal(out.numpy(), full)
    # empty matrix on a non-empty pattern
    E = sp.csr_matrix((N, N))
    out = _place_on_pattern(pp, pi, t(E.indptr, np.int32), t(E.indices, np.int32), t(E.data, np.float64), N)
    assert out.shape == (P.nnz,) and not out.any()
    # an entry outside the pattern (before the first, inside, after the last key) is refused
    Z = P.toarray() == 0
    for (i, j) in (np.argwhere(Z)[0], np.argwhere(Z)[len(np.argwhere(Z)) // 2], np.argwhere(Z)[-1]):
        X = M.tolil()
        X[i, j] = 2.0
        X = X.tocsr()
        with pytest.raises(ValueError, match="1 stored entries outside"):
            _place_on_pattern(pp, pi, t(X.indptr, np.int32), t(X.indices, np.int32), t(X.data, np.float64), N, "A")
    # empty pattern
    z32 = torch.zeros(1, dtype=torch.int32)
    e32 = torch.zeros(0, dtype=torch.int32)
    out = _place_on_pattern(z32, e32, z32.clone(), e32.clone(), torch.zeros(0, dtype=torch.float64), 0)
    assert out.numel() == 0
