"""GPU: overlaps in the mass-matrix inner product (K16/K17, csrc/ws_overlap.cu) through
wavesolve_b200.overlap.  Expected values are the reference's own expression `u.T @ B @ v` (notes
eq. 38) evaluated with scipy on matrices and eigenvectors the LITERAL reference produced
(tests/golden/*.npz: B of construct_AB / construct_B, v of solve_waveguide(_vec))."""

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 1e-12     # relative to the B-norms of the two modes (|u|_B |v|_B = 1 for eigsh output)


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


def _scale(B, U, V):
    au, av = abs(B) @ np.abs(U).T, abs(B) @ np.abs(V).T
    return np.sqrt(np.einsum("in,ni->i", np.abs(U), au))[:, None] * np.sqrt(np.einsum("jn,nj->j", np.abs(V), av))[None, :]


def _check(G, U, V, B, conj=False, tol=TOL):
    ref = (U.conj() if conj else U) @ (B @ V.T)
    assert G.shape == ref.shape
    assert np.all(np.abs(G - ref) <= tol * _scale(B, U, V)), np.abs(G - ref).max()


@pytest.mark.parametrize("name", ["scalar_sq8", "scalar_disc600"])
def test_scalar_overlaps_of_reference_modes(ws, name):
    mesh, ior, z = util.load_golden(name)
    B = util.golden_matrix(z, "Bfull" if name == "scalar_sq8" else "B")
    v = z["v"]                                          # rows, B-orthonormal (eigsh)
    op = ws.mass_operator(mesh)
    G = op.overlap(v)
    assert isinstance(G, np.ndarray) and G.dtype == np.float64
    _check(G, v, v, B)
    assert np.allclose(G, np.eye(len(v)), atol=1e-8)    # eigsh output is B-orthonormal
    # pairs, rectangular stacks, 1-D arguments
    _check(op.overlap(v[:3], v[2:9]), v[:3], v[2:9], B)
    s = op.overlap(v[1], v[1])
    assert np.ndim(s) == 0 and abs(s - v[1] @ (B @ v[1])) <= TOL
    row = op.overlap(v[0], v)
    assert row.shape == (len(v),)
    col = op.overlap(v, v[0])
    assert col.shape == (len(v),)
    assert abs(ws.overlap(v[0], v[2], mesh) - v[0] @ (B @ v[2])) <= TOL      # mesh instead of an operator
    assert np.array_equal(ws.overlap_matrix(v, B=op), G)                     # bitwise reproducible
    op.close()


def test_complex_modes_and_conjugation(ws):
    mesh, ior, z = util.load_golden("scalar_disc600")
    B = util.golden_matrix(z, "B")
    v = z["v"]
    rng = np.random.default_rng(3)
    U = v[:4] + 1j * rng.standard_normal((4, v.shape[1])) * 0.1
    V = rng.standard_normal((3, v.shape[1])) + 1j * v[4:7]
    op = ws.mass_operator(mesh)
    for conj in (False, True):
        G = op.overlap(U, V, conj=conj)
        assert G.dtype == np.complex128
        _check(G, U, V, B, conj=conj)
    _check(op.overlap(U, v[:5]), U, v[:5].astype(complex), B)              # complex x real
    _check(op.overlap(v[:5], V, conj=True), v[:5].astype(complex), V, B, conj=True)
    n2 = op.norms2(U)
    ref = np.einsum("in,ni->i", U.conj(), B @ U.T).real
    assert np.all(np.abs(n2 - ref) <= TOL * np.abs(ref))


def test_vector_modes_of_reference(ws):
    mesh, ior, z = util.load_golden("vec_disc500")
    util.attach_edges(mesh, z)
    k = 2 * np.pi / util.WL
    B = util.golden_matrix(z, "B").tocsr()
    v = z["v"]                                          # complex128 rows from eigs
    op = ws.mass_operator(mesh, ior, k, order="vec")
    assert op.N == B.shape[0]
    _check(op.overlap(v), v, v, B)
    _check(op.overlap(v, conj=True), v, v, B, conj=True)
    vr = np.ascontiguousarray(v.real)
    _check(op.overlap(vr[:5], vr), vr[:5], vr, B)
    with pytest.raises(ValueError):
        ws.mass_operator(mesh, order="vec")


def test_order1_scalar_mass(ws):
    mesh, ior, z = util.load_golden("scalar_p1_disc800")
    B = util.golden_matrix(z, "B").tocsr()
    v = z["v"]
    op = ws.mass_operator(mesh, order=1)
    _check(op.overlap(v), v, v, B)


def test_normalize_and_device_tensors(ws):
    mesh, ior, z = util.load_golden("scalar_disc600")
    B = util.golden_matrix(z, "B")
    N = B.shape[0]
    rng = np.random.default_rng(5)
    U = rng.standard_normal((11, N))                    # more rows than warps per CTA, odd count
    op = ws.mass_operator(mesh)
    ref = U / np.sqrt(np.einsum("in,ni->i", U, B @ U.T))[:, None]
    out = op.normalize(U)
    assert out is not U and np.all(np.abs(out - ref) <= 1e-14 * (1 + np.abs(ref)))
    assert np.all(np.abs(op.norms2(out) - 1.0) <= 1e-12)
    assert np.all(np.abs(ws.normalize_modes(U[2], op) - ref[2]) <= 1e-14 * (1 + np.abs(ref[2])))
    # device in -> device out, real float64 rows are scaled in place
    Ud = torch.from_numpy(U).cuda()
    G = op.overlap(Ud, Ud[:5])
    assert isinstance(G, torch.Tensor) and G.is_cuda and G.shape == (11, 5)
    _check(G.cpu().numpy(), U, U[:5], B)
    r = op.normalize(Ud)
    assert r is Ud and np.all(np.abs(Ud.cpu().numpy() - ref) <= 1e-14 * (1 + np.abs(ref)))
    # complex modes
    Uc = U[:3] + 1j * U[3:6]
    refc = Uc / np.sqrt(np.einsum("in,ni->i", Uc.conj(), B @ Uc.T).real)[:, None]
    assert np.all(np.abs(op.normalize(Uc) - refc) <= 1e-13 * (1 + np.abs(refc)))
    # a mode of zero norm cannot be normalised; wrong lengths are refused
    Z = U.copy()
    Z[4] = 0.0
    with pytest.raises(ValueError):
        op.normalize(Z)
    with pytest.raises(ValueError):
        op.overlap(U[:, :-1])
    # apply: rows of B V
    Y = op.apply(U[:3])
    assert np.all(np.abs(Y - (B @ U[:3].T).T) <= 1e-13 * (1 + np.abs(Y)))
    # empty stacks
    assert op.overlap(U[:0], U).shape == (0, 11)


def test_large_mesh_properties(ws):
    """400k unknowns: symmetry, linearity and agreement with scipy on our own (parity-tested) B."""
    mg = ws.meshgen
    mesh = mg.fiber_disc(200_000, seed=21)
    Bs = ws.construct_B(mesh, sparse=True)
    N = Bs.shape[0]
    rng = np.random.default_rng(9)
    U, V = rng.standard_normal((6, N)), rng.standard_normal((5, N))
    op = ws.mass_operator(mesh)
    G = op.overlap(U, V)
    _check(G, U, V, Bs, tol=1e-12)
    Gt = op.overlap(V, U)
    assert np.all(np.abs(G - Gt.T) <= 1e-12 * _scale(Bs, U, V))                       # B symmetric
    G2 = op.overlap(2.0 * U[:2] - U[2:4], V)
    assert np.all(np.abs(G2 - (2.0 * G[:2] - G[2:4])) <= 4e-12 * _scale(Bs, U[:2], V))   # linear in U
    d = op.norms2(U)
    assert np.all(d > 0) and np.all(np.abs(d - np.diag(op.overlap(U))) <= 1e-12 * d)   # B positive definite
