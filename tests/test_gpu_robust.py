"""GPU parity tests for the two drop-in gaps of round 1:

* wide Krylov bases: the reference's own lantern calls use ``Nmax=100`` (lantern-demo.ipynb cell 2) and
  ``Nmax=80`` (cell 4, the sweep) -> ncv = 204 / 164;
* interior shifts: ``target_neff`` inside (or below) the band of a high-contrast air-hole fibre
  (fe_solver.py:313-316, :375-378; getting-started.ipynb cell 26 uses ``target_neff=0.995``), where
  ``A - sigma B`` is strongly indefinite and the pivot-free LDL^T has to prove its accuracy (static pivots,
  probe solve, iterative refinement, true-residual check) or fail loudly.

Oracle: oracle/restated.py (scipy eigsh / eigs with SuperLU's partial pivoting) on the same mesh arrays.
Tolerances: effective indices 1e-10 relative, fields >= 1 - 1e-8 overlap (degenerate groups by subspace)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NEFF_TOL = 1e-10


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


def _fields_ok(w_ref, v, v_ref, B=None, gap_rtol=1e-6, tol=1e-8):
    for grp in util.group_degenerate(np.real(w_ref), rtol=gap_rtol):
        if grp[-1] == len(w_ref) - 1 and len(w_ref) > 1:
            continue                                   # the last cluster may be cut by Nmax
        ov = util.subspace_overlap(np.real(v[grp]), np.real(v_ref[grp]), B)
        assert ov >= 1 - tol, (grp, ov)


@pytest.mark.parametrize("Nmax", [80, 100])
def test_vector_wide_basis_lantern(ws, Nmax):
    """lantern-demo.ipynb cells 2 / 4: solve_waveguide_vec(m, wl, IOR_dict, Nmax=100 | 80) on a 7-core
    lantern cross-section."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.lantern_cross_section(10000, order=1, seed=7)
    ws.get_unique_edges(mesh)
    w, v, cnt = ws.solve_waveguide_vec(mesh, 1.55, mg.LANTERN_IOR, Nmax=Nmax, verbose=False, sparse_solve_mode="pardiso")
    info = dict(ws.fe_solver.last_info)
    assert w.shape == (Nmax,) and v.shape == (Nmax, len(mesh.cells[0].data) + len(mesh.points))
    assert info["nconv"] == Nmax
    wr, vr, cr = R.solve_waveguide_vec(mesh, 1.55, mg.LANTERN_IOR, Nmax=Nmax)
    order = np.argsort(-wr.real, kind="stable")
    assert np.abs(wr.imag).max() == 0.0
    assert np.allclose(util.neff(w), util.neff(wr[order]), rtol=NEFF_TOL, atol=0)
    assert cnt == R.count_modes_vec(wr.real[order], 2 * np.pi / 1.55, mg.LANTERN_IOR, None)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
    _fields_ok(wr.real[order], v, vr[order], None, gap_rtol=1e-5)


def test_scalar_wide_basis_lantern(ws):
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.lantern_cross_section(10000, order=2, seed=7)
    w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.LANTERN_IOR, Nmax=100)
    assert ws.fe_solver.last_info["nconv"] == 100
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.LANTERN_IOR, Nmax=100)
    assert cnt == cr
    assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0)
    A, B = R.construct_AB_order2(mesh, mg.LANTERN_IOR, 2 * np.pi / 1.55)
    assert np.allclose(np.einsum("ij,ij->i", v, (B @ v.T).T), 1.0, atol=1e-9)
    _fields_ok(wr, v, vr, B)


def test_nmax_above_the_basis_cap_is_served(ws):
    """ncv is clamped at 256: Nmax = 130 still converges (fewer spare directions)."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.fiber_disc(4000, seed=2)
    w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=130)
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=130)
    assert cnt == cr and np.allclose(util.neff(w[w > 0]), util.neff(wr[wr > 0]), rtol=NEFF_TOL, atol=0)
    assert np.allclose(w, wr, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("hollow,target", [(False, 1.40), (False, 1.30), (True, 1.40), (True, 0.995)])
def test_scalar_interior_shift_holey_fibre(ws, hollow, target):
    """getting-started.ipynb cell 26 semantics: target_neff inside / below the band of an air-hole fibre."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.holey_fiber(12000, hollow_core=hollow, seed=3)
    w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.HOLEY_IOR, Nmax=6, target_neff=target)
    info = dict(ws.fe_solver.last_info)
    f = info["factor"]
    assert f["pos_pivots"] > 0 and f["neg_pivots"] > 0             # genuinely indefinite
    assert f["backward_error"] <= 1e-11
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.HOLEY_IOR, Nmax=6, target_neff=target)
    assert cnt == cr
    assert np.allclose(util.neff(w), util.neff(wr), rtol=NEFF_TOL, atol=0)
    A, B = R.construct_AB_order2(mesh, mg.HOLEY_IOR, 2 * np.pi / 1.55)
    _fields_ok(wr, v, vr, B)
    # residuals against the oracle's matrices: the returned pairs solve the pencil
    for i in range(6):
        r = A @ v[i] - w[i] * (B @ v[i])
        assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(A @ v[i])


@pytest.mark.parametrize("hollow,target", [(False, 1.40), (True, 1.40), (True, 0.995)])
def test_vector_interior_shift_holey_fibre(ws, hollow, target):
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.holey_fiber(8000, order=1, hollow_core=hollow, seed=4)
    ws.get_unique_edges(mesh)
    w, v, cnt = ws.solve_waveguide_vec(mesh, 1.55, mg.HOLEY_IOR, Nmax=6, target_neff=target, verbose=False)
    info = dict(ws.fe_solver.last_info)
    assert info["factor"]["backward_error"] <= 1e-11
    wr, vr, cr = R.solve_waveguide_vec(mesh, 1.55, mg.HOLEY_IOR, Nmax=6, target_neff=target)
    order = np.argsort(-wr.real, kind="stable")
    assert np.allclose(util.neff(w), util.neff(wr[order]), rtol=NEFF_TOL, atol=0)
    assert cnt == cr == 6                                           # every w >= 0 counts with target_neff (fe_solver.py:414)
    _fields_ok(wr.real[order], v, vr[order], None, gap_rtol=1e-5)
    A, B = R.construct_AB_vec(mesh, mg.HOLEY_IOR, 2 * np.pi / 1.55)
    for i in range(6):
        r = A @ v[i].real - w[i].real * (B @ v[i].real)
        assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(A @ v[i].real)


def test_default_shift_needs_no_refinement(ws):
    """The guard costs one probe solve on the definite / quasi-definite default-shift matrices: no static pivots,
    no refinement steps, probe backward error at rounding level."""
    mg = ws.meshgen
    mesh = mg.fiber_disc(20000, seed=5)
    ws.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=4)
    f = dict(ws.fe_solver.last_info)["factor"]
    assert f["perturbed_pivots"] == 0 and f["refine_steps"] == 0 and f["backward_error"] <= 1e-13, f
    m1 = mg.fiber_disc(20000, order=1, seed=6)
    ws.get_unique_edges(m1)
    ws.solve_waveguide_vec(m1, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False)
    f = dict(ws.fe_solver.last_info)["factor"]
    assert f["perturbed_pivots"] == 0 and f["refine_steps"] == 0 and f["backward_error"] <= 1e-13, f


def test_shift_on_an_eigenvalue_fails_loudly_or_is_right(ws):
    """sigma placed on an eigenvalue of the pencil (to rounding): A - sigma B is numerically singular.  SuperLU
    happily returns something; here the call must either raise (probe / residual check) or return eigenpairs
    that satisfy the pencil -- never silently wrong ones."""
    from oracle import restated as R
    mg = ws.meshgen
    mesh = mg.fiber_disc(3000, seed=9)
    k = 2 * np.pi / 1.55
    wr, vr, cr = R.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=3)
    target = float(np.sqrt(wr[1]) / k)                    # LP11: sigma = (k target)^2 == wr[1] up to rounding
    A, B = R.construct_AB_order2(mesh, mg.FIBER_IOR, k)
    try:
        w, v, cnt = ws.solve_waveguide(mesh, 1.55, mg.FIBER_IOR, Nmax=3, target_neff=target)
    except ws._lib.WavesolveB200Error as e:
        assert "(2)" in str(e) or "(3)" in str(e)         # WS_ERR_NUMERIC / WS_ERR_NOCONV, with a message
        return
    for i in range(3):
        r = A @ v[i] - w[i] * (B @ v[i])
        assert np.linalg.norm(r) <= 1e-7 * np.linalg.norm(A @ v[i]), (i, w, wr)
