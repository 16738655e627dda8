"""GPU parity tests for the eigen-solves behind solve_waveguide / solve_waveguide_vec:
effective indices within 1e-10 relative, mode fields at >= 1 - 1e-8 overlap after sign fixing
(degenerate pairs by subspace angle), against the golden fixtures of the literal reference and
against the numpy/scipy oracle on larger meshes."""
import copy

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NEFF_TOL = 1e-10
OVERLAP_TOL = 1e-8


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


def check_fields(w, v, w_ref, v_ref, B=None, gap_rtol=1e-6):
    """per non-degenerate mode: overlap >= 1-1e-8; per cluster: subspace overlap."""
    for grp in util.group_degenerate(np.real(w_ref), rtol=gap_rtol):
        if grp[-1] == len(w_ref) - 1 and len(w_ref) > 1:
            continue   # the last cluster may be cut by Nmax
        U = np.real(v[grp])
        Vr = np.real(v_ref[grp])
        if len(grp) == 1:
            if B is not None:
                ov = util.b_overlap(U[0], Vr[0], B)
            else:
                ov = abs(U[0] @ Vr[0]) / (np.linalg.norm(U[0]) * np.linalg.norm(Vr[0]))
        else:
            ov = util.subspace_overlap(U, Vr, B)
        assert ov >= 1 - OVERLAP_TOL, (grp, ov)


@pytest.mark.parametrize("name", ["scalar_sq8", "scalar_disc600", "scalar_disc600_target"])
def test_scalar_solve_vs_reference_golden(ws, name):
    mesh, ior, z = util.load_golden(name)
    tn = float(z["target_neff"])
    w, v, cnt = ws.solve_waveguide(mesh, float(z["wl"]), ior, Nmax=int(z["Nmax"]),
                                   target_neff=None if np.isnan(tn) else tn)
    assert cnt == int(z["count"])
    assert w.shape == z["w"].shape and v.shape == z["v"].shape and w.dtype == np.float64
    assert np.all(np.diff(w) <= 0)
    assert np.allclose(util.neff(w), util.neff(z["w"]), rtol=NEFF_TOL, atol=0)
    B = util.golden_matrix(z, "B")
    assert np.allclose(np.einsum("ij,ij->i", v, (B @ v.T).T), 1.0, atol=1e-10)   # B-orthonormal like eigsh
    check_fields(w, v, z["w"], z["v"], B)


@pytest.mark.parametrize("name", ["scalar_c1_disc5k", "scalar_lantern2k"])
def test_scalar_solve_c1_eigenvalues(ws, name):
    mesh, ior, z = util.load_golden(name)
    w, v, cnt = ws.solve_waveguide(mesh, float(z["wl"]), ior, Nmax=int(z["Nmax"]))
    assert cnt == int(z["count"])
    assert np.allclose(util.neff(w), util.neff(z["w"]), rtol=NEFF_TOL, atol=0)


@pytest.mark.parametrize("name", ["vec_sq6", "vec_sq12", "vec_disc500", "vec_disc1500", "vec_disc500_target"])
def test_vector_solve_vs_reference_golden(ws, name):
    mesh, ior, z = util.load_golden(name)
    util.attach_edges(mesh, z)
    tn = float(z["target_neff"])
    w, v, cnt = ws.solve_waveguide_vec(mesh, float(z["wl"]), ior, Nmax=int(z["Nmax"]), verbose=False,
                                       target_neff=None if np.isnan(tn) else tn)
    assert w.dtype == np.complex128 and v.dtype == np.complex128 and v.shape == z["v"].shape
    assert np.all(np.diff(w.real) <= 0) and np.all(w.imag == 0)
    order = np.argsort(-z["w"].real, kind="stable")       # the reference returns ARPACK's order
    w_ref, v_ref = z["w"][order], z["v"][order]
    assert np.allclose(util.neff(w), util.neff(w_ref), rtol=NEFF_TOL, atol=0)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
    check_fields(w, v, w_ref, v_ref, None, gap_rtol=1e-5)
    # mode count: the reference counts in ARPACK order; in sorted order the same rule gives
    k = 2 * np.pi / float(z["wl"])
    from oracle import restated as R
    assert cnt == R.count_modes_vec(w_ref.real, k, ior, None if np.isnan(tn) else tn)


def test_vector_solve_modes_and_dtype_contract(ws):
    mesh, ior, z = util.load_golden("vec_sq12")
    util.attach_edges(mesh, z)
    w, v, cnt = ws.solve_waveguide_vec(mesh, float(z["wl"]), ior, Nmax=4, verbose=False, sparse_solve_mode="straight")
    assert w.dtype == np.float64 and v.dtype == np.float64
    with pytest.raises(AssertionError):
        ws.solve_waveguide_vec(mesh, float(z["wl"]), ior, Nmax=4, verbose=False, sparse_solve_mode="nope")
    m6, ior6, z6 = util.load_golden("scalar_sq8")
    with pytest.raises(AssertionError):
        ws.solve_waveguide_vec(m6, 1.55, ior6)             # needs an order-1 mesh (fe_solver.py:370)
    with pytest.raises(AssertionError):
        ws.solve_waveguide(mesh, 1.55, ior)                # needs an order-2 mesh (fe_solver.py:140)


def test_scalar_solve_vs_oracle_50k(ws):
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.fiber_disc(50000, seed=21)
    w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=10)
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=10)
    assert cnt == cr == 3
    assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0)
    A, B = R.construct_AB_order2(mesh, mg.FIBER_IOR, 2 * np.pi / 1.55)
    check_fields(w, v, wr, vr, B)
    info = ws.fe_solver.last_info
    assert info["nconv"] == 10 and info["factor"]["pos_pivots"] == 0


def test_vector_solve_vs_oracle_50k(ws):
    """config C2: ~50k elements, HE11 x2, TE01, TM01, HE21 x2 guided."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.fiber_disc(50000, order=1, seed=22)
    ws.get_unique_edges(mesh)
    w, v, cnt = ws.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=10, verbose=False)
    wr, vr, cr = R.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=10)
    order = np.argsort(-wr.real, kind="stable")
    assert np.allclose(util.neff(w), util.neff(wr[order]), rtol=NEFF_TOL, atol=0)
    assert cnt == 6
    check_fields(w, v, wr[order], vr[order], None, gap_rtol=1e-5)


def test_lantern_scalar_modes(ws):
    """config C3 geometry (reduced size): three materials, count vs the oracle."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.lantern_cross_section(40000, seed=5)
    w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.LANTERN_IOR, Nmax=10)
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.LANTERN_IOR, Nmax=10)
    assert cnt == cr
    assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0)


def test_full_size_vector_1M_properties(ws):
    """config C4 (~1M triangles, 10 modes): residuals of the returned eigenpairs against the
    assembled matrices, guided-mode count, degeneracy structure of the step-index fibre."""
    import scipy.sparse as sp
    mg = ws.meshgen
    mesh = mg.fiber_disc(1_000_000, order=1, seed=0)
    ws.get_unique_edges(mesh)
    w, v, cnt = ws.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=10, verbose=False)
    info = dict(ws.fe_solver.last_info)
    assert info["nconv"] == 10
    k = 2 * np.pi / 1.55
    A, B = ws.construct_AB_vec(mesh, mg.FIBER_IOR, k, sparse=True)
    w = w.real; v = v.real
    for i in range(10):
        r = A @ v[i] - w[i] * (B @ v[i])
        assert np.linalg.norm(r) <= 1e-9 * np.linalg.norm(A @ v[i]), (i, np.linalg.norm(r))
    assert cnt == 6
    ne = util.neff(w)
    gaps = np.abs(np.diff(ne[:6]))
    assert gaps[0] < 1e-6                                   # HE11 x/y pair
    assert (gaps[1:] < 1e-5).sum() == 1                     # exactly one more pair (HE21) among TE01, TM01, HE21 x2
    assert ne[0] < 1.4528 and ne[5] > 1.444
    ntt = len(mesh.cells[0].data)
    assert info["factor"]["neg_pivots"] == ntt and info["factor"]["pos_pivots"] == len(mesh.points)


@pytest.mark.parametrize("name", ["scalar_p1_sq8", "scalar_p1_disc800"])
def test_scalar_p1_assembly_and_solve_vs_reference_golden(ws, name):
    """order=1 scalar path (fe_solver.py:103-129, solve_waveguide(order=1))."""
    import scipy.sparse as sp
    mesh, ior, z = util.load_golden(name)
    k = 2 * np.pi / float(z["wl"])
    A, B = ws.construct_AB(mesh, ior, k, sparse=True, order=1)
    assert sp.issparse(A) and A.format == "lil" and B.format == "lil"
    for M, pre in ((A, "A"), (B, "B")):
        G = util.golden_matrix(z, pre)
        util.assert_same_pattern(M.tocsr(), G)     # exact zeros pruned like lil does
        util.assert_values_close(M.tocsr(), G)
    Ad, Bd = ws.construct_AB(mesh, ior, k, sparse=False, order=1)
    assert Ad.shape == (len(mesh.points),) * 2 and np.allclose(Bd, util.golden_matrix(z, "B").toarray(), rtol=1e-12, atol=1e-18)
    w, v, cnt = ws.solve_waveguide(mesh, float(z["wl"]), ior, Nmax=int(z["Nmax"]), order=1)
    assert cnt == int(z["count"]) and v.shape == z["v"].shape
    assert np.allclose(util.neff(w), util.neff(z["w"]), rtol=NEFF_TOL, atol=0)
    check_fields(w, v, z["w"], z["v"], util.golden_matrix(z, "B"))
    with pytest.raises(AssertionError):
        ws.solve_waveguide(mesh, float(z["wl"]), ior, order=2)


def test_block2_arnoldi_matches_single_vector(ws):
    """kind 2 (block size 2: two basis vectors share one pair of sweeps) finds the same eigenpairs
    as the single-vector iteration."""
    from wavesolve_b200 import fe_solver as fs, solver as sv
    mg = ws.meshgen
    mesh = mg.fiber_disc(20000, order=1, seed=8)
    ws.get_unique_edges(mesh)
    k = 2 * np.pi / 1.55
    sigma = (k * max(mg.FIBER_IOR.values())) ** 2
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, k)
    Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
    S = sv.Solver(prob.pattern, sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt))
    S.factor(Mq)
    out = {}
    for kind in (1, 2):
        w, V, info = fs._eig_device(prob.pattern, S, B, sigma, kind, 8, prob.N, prob.dev, edges=prob.edges, xy=prob.xy,
                                    Ntt=prob.Ntt)
        out[kind] = (w.cpu().numpy(), V.cpu().numpy(), info)
        assert info["nconv"] == 8
    w1, V1, _ = out[1]
    w2, V2, i2 = out[2]
    assert i2["n_op"] % 2 == 0
    assert np.max(np.abs(np.sqrt(w1) - np.sqrt(w2)) / np.sqrt(w1)) < 1e-10
    for g in util.group_degenerate(w1, rtol=1e-7):
        assert util.subspace_overlap(V1[g], V2[g]) >= 1 - 1e-8
