"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import scipy.sparse as sp

from wavesolve_b200.meshgen import Mesh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WL = 1.55


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    labels = [str(s) for s in z["labels"]]
    sets = {lab: [None, z["set_%d" % i], None] for i, lab in enumerate(labels)}
    mesh = Mesh(z["points"], z["tris"], sets)
    ior = {lab: float(v) for lab, v in zip(labels, z["ior"])}
    return mesh, ior, z


def load_golden_mesh(name):
    """Fixtures without material data (mode resampling): mesh + arrays."""
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    labels = [str(s) for s in z["labels"]]
    sets = {lab: [None, z["set_%d" % i], None] for i, lab in enumerate(labels)}
    mesh = Mesh(z["points"], z["tris"], sets)
    if "edges" in z.files:
        mesh.cells[0].data = z["edges"]
        mesh.edge_indices = z["edge_indices"]
        mesh.num_edges = len(z["edges"])
        mesh.num_points = mesh.points.shape[0]
    return mesh, z


def golden_matrix(z, prefix):
    fmt = str(z[prefix + "_format"])
    cls = sp.csr_matrix if fmt == "csr" else sp.csc_matrix
    return cls((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]),
               shape=tuple(z[prefix + "_shape"]))


def attach_edges(mesh, z):
    mesh.cells[0].data = z["edges"]
    mesh.edge_indices = z["edge_indices"]
    mesh.edge_flips = z["edge_flips"]
    mesh.num_edges = len(z["edges"])
    mesh.num_points = mesh.points.shape[0]


def assert_same_pattern(M, R):
    """bit-exact sparsity pattern: format, shape, dtype, indptr, indices."""
    M = M.copy(); M.sort_indices()
    R = R.copy(); R.sort_indices()
    assert M.format == R.format
    assert M.shape == R.shape
    assert M.indices.dtype == R.indices.dtype and M.indptr.dtype == R.indptr.dtype
    assert np.array_equal(M.indptr, R.indptr)
    assert np.array_equal(M.indices, R.indices)


def value_errors(M, R):
    """(max entry-wise relative error over entries with |ref| > 1e-3*max|ref|,
        max abs error / max|ref|).  Both matrices must share the pattern.

    Why the entry-wise figure stops at 1e-3*max|ref|: an assembled entry is a sum of up to ~8 element contributions
    of size ~max|ref|, each carrying a rounding error of ~1e-16*max|ref| from the element formulas (k^2 n^2 NN - dNdN
    is a difference of two products, fe_solver.py:51); entries that cancel to below 1e-3*max|ref| therefore have an
    entry-wise relative error of up to 1e-16/1e-3 per ulp of disagreement between two CORRECT summations, and an entry
    that cancels to ~0 has an unbounded one.  The bound that holds for every stored entry, cancelling or not, is the
    absolute one, |err| <= tol * max|ref| (second figure); both are asserted, and the vector path additionally asserts
    bit-identical values (same arithmetic, same summation order) where the oracle can run (test_gpu_assembly.py)."""
    M = M.copy(); M.sort_indices()
    R = R.copy(); R.sort_indices()
    d = np.abs(M.data - R.data)
    scale = np.abs(R.data).max() if R.nnz else 1.0
    big = np.abs(R.data) > 1e-3 * scale
    rel = (d[big] / np.abs(R.data[big])).max() if big.any() else 0.0
    return rel, (d.max() / scale if R.nnz else 0.0)


VAL_TOL = 1e-12   # north_star: assembled A/B values within 1e-12 relative


def assert_values_close(M, R, tol=VAL_TOL):
    rel, nrm = value_errors(M, R)
    assert rel <= tol, "entry-wise relative error %.3e > %g" % (rel, tol)
    assert nrm <= tol, "max-norm relative error %.3e > %g" % (nrm, tol)


def neff(w, wl=WL):
    k = 2 * np.pi / wl
    return np.sqrt(np.real(w)) / k


def b_overlap(u, v, B):
    """|u^T B v| / sqrt(u^T B u * v^T B v)"""
    Bu, Bv = B @ u, B @ v
    return abs(v @ Bu) / np.sqrt(abs(u @ Bu) * abs(v @ Bv))


def subspace_overlap(U, V, B=None):
    """smallest cosine of the principal angles between span(U) and span(V) (rows = vectors),
    in the B inner product if B is given (B SPD) else Euclidean."""
    U = np.asarray(U, dtype=float).T
    V = np.asarray(V, dtype=float).T
    if B is not None:
        def orth(X):
            G = X.T @ (B @ X)
            L = np.linalg.cholesky(G)
            return np.linalg.solve(L, X.T).T
        Qu, Qv = orth(U), orth(V)
        s = np.linalg.svd(Qu.T @ (B @ Qv), compute_uv=False)
    else:
        Qu, _ = np.linalg.qr(U)
        Qv, _ = np.linalg.qr(V)
        s = np.linalg.svd(Qu.T @ Qv, compute_uv=False)
    return s.min()


def group_degenerate(w, rtol=1e-6):
    """indices grouped into clusters of (nearly) equal eigenvalues (w sorted descending)."""
    groups, cur = [], [0]
    for i in range(1, len(w)):
        if abs(w[i] - w[cur[-1]]) <= rtol * abs(w[i]):
            cur.append(i)
        else:
            groups.append(cur)
            cur = [i]
    groups.append(cur)
    return groups
