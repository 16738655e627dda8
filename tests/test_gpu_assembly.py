"""GPU parity tests (through the C ABI) for K1-K5: edge numbering, pattern, scalar and
vector assembly, zero pruning, SpMV.  Compared against the golden fixtures generated from
the literal reference, against the numpy oracle on larger seeded meshes, and -- at the full
1M-element size -- through size-independent properties."""
import copy

import numpy as np
import pytest
import scipy.sparse as sp

from tests import util

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


SCALAR = ["scalar_sq8", "scalar_disc600", "scalar_sq5_uint64", "scalar_sq6_ragged",
          "scalar_c1_disc5k", "scalar_lantern2k"]
VECTOR = ["vec_sq6", "vec_sq12", "vec_disc500", "vec_disc1500", "vec_sq5_uint64"]


@pytest.mark.parametrize("name", SCALAR)
def test_scalar_assembly_vs_reference_golden(ws, name):
    mesh, ior, z = util.load_golden(name)
    k = 2 * np.pi / float(z["wl"])
    A, B = ws.construct_AB(mesh, ior, k, sparse=True, order=2)
    for M, pre in ((A, "A"), (B, "B")):
        G = util.golden_matrix(z, pre)
        util.assert_same_pattern(M, G)          # bit-exact, explicit zeros included
        util.assert_values_close(M, G)          # <= 1e-12 relative
    # element entries are bit-identical to numpy's; only the duplicate-summation order can
    # differ, so the bulk of the values must match exactly
    assert (B.data == util.golden_matrix(z, "B").data).mean() > 0.5


def test_mass_matrix_all_cells(ws):
    mesh, ior, z = util.load_golden("scalar_sq8")
    B = ws.construct_B(mesh, sparse=True)
    G = util.golden_matrix(z, "Bfull")
    util.assert_same_pattern(B, G)
    util.assert_values_close(B, G)
    Bd = ws.construct_B(mesh, sparse=False)
    assert Bd.shape == (len(mesh.points),) * 2
    assert np.allclose(Bd, G.toarray(), rtol=1e-12, atol=1e-18)


def test_dense_return_and_asserts(ws):
    mesh, ior, z = util.load_golden("scalar_sq5_uint64")
    k = 2 * np.pi / float(z["wl"])
    A, B = ws.construct_AB(mesh, ior, k, sparse=False, order=2)
    assert isinstance(A, np.ndarray) and A.shape == tuple(z["A_shape"])
    assert np.allclose(A, util.golden_matrix(z, "A").toarray(), rtol=1e-12, atol=1e-14)
    with pytest.raises(AssertionError):
        ws.construct_AB(mesh, ior, k, sparse=True, order=1)     # fe_solver.py:143
    with pytest.raises(NotImplementedError):
        ws.construct_AB(mesh, ior, k, sparse=True, order=3)     # fe_solver.py:146
    with pytest.raises(KeyError):
        ws.construct_AB(mesh, {"core": 1.45}, k, sparse=True)   # IOR_dict must hold every label


@pytest.mark.parametrize("name", VECTOR)
def test_edges_vs_reference_golden(ws, name):
    mesh, ior, z = util.load_golden(name)
    m = copy.deepcopy(mesh)
    edges, eidx = ws.get_unique_edges(m)
    assert np.array_equal(edges, z["edges"]) and edges.dtype == z["edges"].dtype
    assert np.array_equal(eidx, z["edge_indices"]) and eidx.dtype == z["edge_indices"].dtype
    assert np.array_equal(m.edge_flips, z["edge_flips"])
    assert m.num_edges == len(z["edges"]) and m.num_points == len(mesh.points)
    m2 = copy.deepcopy(mesh)
    ws.get_unique_edges(m2, mutate=False)
    assert not hasattr(m2, "edge_indices")


@pytest.mark.parametrize("name", VECTOR)
def test_vector_assembly_vs_reference_golden(ws, name):
    mesh, ior, z = util.load_golden(name)
    util.attach_edges(mesh, z)
    k = 2 * np.pi / float(z["wl"])
    A, B = ws.construct_AB_vec(mesh, ior, k, sparse=True)
    for M, pre in ((A, "A"), (B, "B")):
        G = util.golden_matrix(z, pre)
        util.assert_same_pattern(M, G)          # CSC, int32, exact zeros pruned like lil->csc
        util.assert_values_close(M, G)
    Ad, Bd = ws.construct_AB_vec(mesh, ior, k, sparse=False)
    assert np.allclose(Bd, util.golden_matrix(z, "B").toarray(), rtol=1e-12, atol=1e-14)


def _oracle():
    from oracle import restated
    return restated


@pytest.mark.parametrize("gen,arg", [("fiber_disc", 50000), ("structured_square", 150),
                                     ("lantern_cross_section", 60000)])
def test_scalar_assembly_vs_oracle_large(ws, gen, arg):
    R = _oracle()
    mg = ws.meshgen
    mesh = getattr(mg, gen)(arg)
    ior = mg.LANTERN_IOR if gen.startswith("lantern") else mg.FIBER_IOR
    k = 2 * np.pi / 1.55
    A, B = ws.construct_AB(mesh, ior, k, sparse=True)
    Ar, Br = R.construct_AB_order2(mesh, ior, k)
    for M, G in ((A, Ar), (B, Br)):
        util.assert_same_pattern(M, G)
        util.assert_values_close(M, G)


@pytest.mark.parametrize("gen,arg", [("fiber_disc", 50000), ("structured_square", 100)])
def test_vector_path_vs_oracle_large(ws, gen, arg):
    R = _oracle()
    mg = ws.meshgen
    mesh = getattr(mg, gen)(arg, order=1)
    m_ref = copy.deepcopy(mesh)
    e_ref, i_ref = R.get_unique_edges(m_ref)
    edges, eidx = ws.get_unique_edges(mesh)
    assert np.array_equal(edges, e_ref) and np.array_equal(eidx, i_ref)
    k = 2 * np.pi / 1.55
    A, B = ws.construct_AB_vec(mesh, mg.FIBER_IOR, k, sparse=True)
    Ar, Br = R.construct_AB_vec(m_ref, mg.FIBER_IOR, k)
    for M, G in ((A, Ar), (B, Br)):
        util.assert_same_pattern(M, G)
        util.assert_values_close(M, G)
        assert np.array_equal(M.data, G.data)   # same arithmetic and same summation order


def test_full_size_properties_scalar_1M(ws):
    """1M-element P2 mesh: properties that do not need the oracle."""
    mg = ws.meshgen
    n = 708                                           # 2*708^2 = 1,002,528 triangles
    mesh = mg.structured_square(n)
    k = 2 * np.pi / 1.55
    A, B = ws.construct_AB(mesh, mg.FIBER_IOR, k, sparse=True)
    N = A.shape[0]
    assert N == (2 * n + 1) ** 2
    assert A.nnz == B.nnz and np.array_equal(A.indices, B.indices)
    one = np.ones(N)
    area = 20.0 * 20.0
    assert abs(one @ (B @ one) - area) <= 1e-10 * area          # sum of the mass matrix = area
    # stiffness rows sum to zero: A 1 = (k^2 n^2 M) 1, so 1^T A 1 = k^2 * int n^2
    cs = mesh.cell_sets
    pts = mesh.points
    tri = mesh.cells[1].data
    def area_of(idx):
        t = tri[idx][:, :3]
        a, b, c = pts[t[:, 0]], pts[t[:, 1]], pts[t[:, 2]]
        return 0.5 * np.sum((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    want = k ** 2 * sum(mg.FIBER_IOR[m] ** 2 * area_of(cs[m][1]) for m in cs)
    assert abs(one @ (A @ one) - want) <= 1e-9 * abs(want)
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    assert abs(B - B.T).max() == 0.0


def test_full_size_properties_vector_1M(ws):
    mg = ws.meshgen
    n = 708
    mesh = mg.structured_square(n, order=1)
    edges, eidx = ws.get_unique_edges(mesh)
    Np = (n + 1) ** 2
    Ne = 2 * n * n
    assert len(edges) == Np + Ne - 1                    # Euler: V - E + F = 1 for a disc
    assert np.all(edges[:, 0] < edges[:, 1])
    # first-appearance numbering: ids are introduced in increasing order along the scan
    flat = eidx.astype(np.int64).ravel()
    first = np.full(len(edges), len(flat), dtype=np.int64)
    np.minimum.at(first, flat, np.arange(len(flat)))
    assert np.all(np.diff(first) > 0)
    k = 2 * np.pi / 1.55
    A, B = ws.construct_AB_vec(mesh, mg.FIBER_IOR, k, sparse=True)
    Ntt = len(edges)
    assert A.shape == (Ntt + Np, Ntt + Np)
    assert A[Ntt:, :].nnz == 0 and A[:, Ntt:].nnz == 0   # A only couples edges (fe_solver.py:206)
    assert abs(B - B.T).max() == 0.0
    assert (A.data == 0).sum() == 0 and (B.data == 0).sum() == 0
    # gradient fields are in the null space of curl-curl: (Att - k^2n^2 Btt) G = 0 for uniform n
    one_ior = {m: 1.0 for m in mesh.cell_sets}
    A1, B1 = ws.construct_AB_vec(mesh, one_ior, k, sparse=True)
    Ktt = (k ** 2) * B1[:Ntt, :Ntt] - A1[:Ntt, :Ntt]      # = curl-curl block
    phi = np.sin(0.3 * mesh.points[:, 0]) * np.cos(0.2 * mesh.points[:, 1])
    # Whitney dof of grad(phi) on edge (a,b), a<b (signed lengths, shape_funcs.py:317-322) is
    # the tangential component of the piecewise-constant gradient: (phi_b - phi_a) / length
    lens = np.linalg.norm(mesh.points[edges[:, 1], :2] - mesh.points[edges[:, 0], :2], axis=1)
    g = (phi[edges[:, 1]] - phi[edges[:, 0]]) / lens
    r = Ktt @ g
    assert np.abs(r).max() <= 1e-9 * np.abs(Ktt).max() * np.abs(g).max()


def test_spmv_matches_scipy(ws):
    from wavesolve_b200 import fe_solver as fs
    mg = ws.meshgen
    mesh = mg.fiber_disc(20000, seed=5)
    k = 2 * np.pi / 1.55
    prob = fs.ScalarProblem(mesh, mg.FIBER_IOR, k)
    A, B = prob.assemble()
    indptr, indices = prob.pattern.export()
    Am = sp.csr_matrix((A.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(prob.N,) * 2)
    x = np.random.default_rng(0).standard_normal(prob.N)
    y = prob.pattern.spmv(A, torch.from_numpy(x).cuda()).cpu().numpy()
    ref = Am @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


def test_empty_and_tiny_meshes(ws):
    mg = ws.meshgen
    mesh = mg.structured_square(1)                       # two elements
    A, B = ws.construct_AB(mesh, mg.FIBER_IOR, 4.0, sparse=True)
    R = _oracle()
    Ar, Br = R.construct_AB_order2(mesh, mg.FIBER_IOR, 4.0)
    util.assert_same_pattern(A, Ar)
    util.assert_values_close(A, Ar)
    m1 = mg.structured_square(1, order=1)
    m1.cells[1].data = m1.cells[1].data[:0]
    edges, eidx = ws.get_unique_edges(m1, mutate=False)
    assert edges.shape[0] == 0 and eidx.shape == (0, 3)


def test_fast_mode_and_fused_qd_assembly(ws):
    """WS_ASM_FAST (reciprocal arithmetic inside the solve pipeline) stays within a few ulp of the
    exact mode; the fused pass returns the same M' and B as the separate ones."""
    from wavesolve_b200 import fe_solver as fs
    mg = ws.meshgen
    k = 2 * np.pi / 1.55
    sp_ = fs.ScalarProblem(mg.fiber_disc(30000, seed=8), mg.FIBER_IOR, k)
    A, B = sp_.assemble()
    Af, Bf = sp_.assemble(fast=True)
    for X, Y in ((A, Af), (B, Bf)):
        X, Y = X.cpu().numpy(), Y.cpu().numpy()
        assert np.abs(X - Y).max() <= 1e-14 * np.abs(X).max()
    mesh = mg.fiber_disc(30000, order=1, seed=9)
    ws.get_unique_edges(mesh)
    vp = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    sigma = (k * 1.4528) ** 2
    A, B = vp.assemble()
    M1 = vp.assemble_qd(sigma)
    M2, B2 = vp.assemble_qd(sigma, with_B=True)
    assert torch.equal(M1, M2) and torch.equal(B, B2)
    M3, B3 = vp.assemble_qd(sigma, with_B=True, fast=True)
    assert (M3 - M1).abs().max().item() <= 1e-13 * M1.abs().max().item()
    assert (B3 - B).abs().max().item() <= 1e-14 * B.abs().max().item()


def test_pinned_result_buffers_are_recycled(ws):
    """Large results come back through page-locked blocks that are recycled once the caller drops the
    array and every view of it (wavesolve_b200.device.d2h_pinned); the cache is bounded across sizes."""
    import gc
    from wavesolve_b200 import device as d
    x = torch.arange(1 << 18, dtype=torch.float64, device="cuda")
    del d._pinned_free[:]
    a = d.d2h_pinned(x)
    assert a.dtype == np.float64 and np.array_equal(a, np.arange(1 << 18, dtype=np.float64))
    p0 = a.ctypes.data
    view = a[:16].view(np.uint64)
    del a
    gc.collect()
    assert not d._pinned_free                             # a view still uses the block
    b = d.d2h_pinned(x * 2)                               # needs a second block
    assert b.ctypes.data != p0 and b[3] == 6.0
    del view
    gc.collect()
    assert len(d._pinned_free) == 1
    c = d.d2h_pinned(x + 1)
    assert c.ctypes.data == p0 and c[0] == 1.0 and b[3] == 6.0   # recycled, and b is untouched
    # a slightly smaller request (the next z-slice of a sweep) is served by the same block; many distinct sizes
    # never park more than _PINNED_KEEP blocks
    del c
    gc.collect()
    e = d.d2h_pinned(x[: (1 << 18) - 4096])
    assert e.ctypes.data == p0 and e[7] == 7.0
    del e, b
    for n in range(8):
        t = d.d2h_pinned(torch.zeros((1 << 17) + 40000 * n, dtype=torch.float64, device="cuda"))
        del t
    gc.collect()
    assert len(d._pinned_free) <= d._PINNED_KEEP
    z = d.d2h_pinned(torch.view_as_complex(torch.ones(1 << 17, 2, dtype=torch.float64, device="cuda")))
    assert z.dtype == np.complex128 and z[5] == 1 + 1j


def test_staged_d2h_of_a_large_result(ws):
    """GB-scale gather payloads come back through two reusable pinned staging blocks into pageable memory."""
    from wavesolve_b200 import device as d
    n = (3 * (64 << 20)) // 8 + 12345
    x = torch.arange(n, dtype=torch.float64, device="cuda") * 0.5
    a = d.d2h_staged(x)
    assert a.dtype == np.float64 and a.shape == (n,) and a[0] == 0.0 and a[-1] == (n - 1) * 0.5
    assert np.array_equal(a[:: 1 << 16], (np.arange(n, dtype=np.float64) * 0.5)[:: 1 << 16])
    z = d.d2h_staged(x[: 2 * (n // 2)].view(-1, 2))
    assert z.shape == (n // 2, 2) and z[5, 1] == 5.5
