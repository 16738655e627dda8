"""CPU: the small dense eigen-solvers used for the projected Krylov problem (host C++,
csrc/ws_eig.cu) against numpy.linalg."""
import numpy as np
import pytest

from wavesolve_b200 import _lib


def dense_eig(A, symmetric=False):
    n = A.shape[0]
    A = np.ascontiguousarray(A, dtype=np.float64)
    wr = np.zeros(n); wi = np.zeros(n); V = np.zeros((n, n))
    _lib.check(_lib.lib().ws_host_eig_general(n, A.ctypes.data, wr.ctypes.data, wi.ctypes.data, V.ctypes.data,
                                              1 if symmetric else 0))
    return wr, wi, V


@pytest.mark.parametrize("n", [1, 2, 5, 21, 41, 48, 49, 164, 204, 256])
def test_symmetric_jacobi(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); A = A + A.T
    w, _, V = dense_eig(A, symmetric=True)
    assert np.allclose(np.sort(w), np.linalg.eigvalsh(A), rtol=1e-13, atol=1e-13)
    assert np.allclose(A @ V, V * w, atol=1e-12 * max(1, abs(A).max()))
    assert np.allclose(V.T @ V, np.eye(n), atol=1e-13)


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (6, 2), (21, 3), (21, 4), (41, 5), (33, 6), (164, 7), (204, 8), (256, 9)])
def test_general_real_matrix(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    wr, wi, V = dense_eig(A)
    ref = np.linalg.eigvals(A)
    got = wr + 1j * wi
    assert np.allclose(np.sort_complex(got), np.sort_complex(ref), rtol=1e-10, atol=1e-10)
    j = 0
    while j < n:
        if wi[j] == 0:
            v = V[:, j]
            assert np.linalg.norm(A @ v - wr[j] * v) <= 1e-10 * np.linalg.norm(v) * max(1, abs(A).max())
            j += 1
        else:
            assert wi[j] > 0 and wi[j + 1] == -wi[j]
            v = V[:, j] + 1j * V[:, j + 1]
            assert np.linalg.norm(A @ v - got[j] * v) <= 1e-10 * np.linalg.norm(v) * max(1, abs(A).max())
            j += 2


def test_general_on_arrowhead_plus_hessenberg():
    """the shape the thick-restart produces: dense k x k block, a full row, then Hessenberg."""
    rng = np.random.default_rng(9)
    n, k = 21, 10
    A = np.triu(rng.standard_normal((n, n)), -1)
    A[:k, :k] = np.diag(-np.sort(rng.random(k))[::-1] * 50)
    A[k, :k] = rng.standard_normal(k) * 1e-3
    wr, wi, V = dense_eig(A)
    assert np.allclose(np.sort_complex(wr + 1j * wi), np.sort_complex(np.linalg.eigvals(A)), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n,k", [(164, 80), (204, 100)])
def test_wide_restart_shapes(n, k):
    """Nmax = 80 / 100 (lantern-demo.ipynb cells 4 / 2): ncv = 164 / 204.  The projected matrix after a thick
    restart, symmetric (scalar path: diagonal block + arrow + tridiagonal) and general."""
    rng = np.random.default_rng(n)
    T = np.zeros((n, n))
    T[:k, :k] = np.diag(-np.sort(rng.random(k))[::-1] * 1e3)
    T[k, :k] = T[:k, k] = rng.standard_normal(k) * 1e-4
    for i in range(k, n):
        T[i, i] = -rng.random() * 10
        if i + 1 < n:
            T[i, i + 1] = T[i + 1, i] = rng.random()
    w, _, V = dense_eig(T, symmetric=True)
    assert np.allclose(np.sort(w), np.linalg.eigvalsh(T), rtol=1e-12, atol=1e-12)
    assert np.allclose(T @ V, V * w, atol=1e-10)
    assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)
    H = np.triu(rng.standard_normal((n, n)), -1)
    H[:k, :k] = T[:k, :k] + 1e-3 * rng.standard_normal((k, k))
    H[k, :k] = rng.standard_normal(k) * 1e-3
    wr, wi, V = dense_eig(H)
    got, ref = np.sort_complex(wr + 1j * wi), np.sort_complex(np.linalg.eigvals(H))
    assert np.allclose(got, ref, rtol=1e-8, atol=1e-8)
