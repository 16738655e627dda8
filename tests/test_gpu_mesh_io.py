"""GPU: a mesh that went through the compact .npz container (wavesolve_b200.mesh_io) assembles to
bit-identical matrices and gives the same modes as the original object."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_vector_problem_from_container(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200 as ws
    mg = ws.meshgen
    mesh = mg.fiber_disc(2000, order=1, seed=8)
    ws.get_unique_edges(mesh)
    path = str(tmp_path / "fibre.npz")
    ws.save_mesh_npz(path, mesh)
    back = ws.load_mesh_npz(path)
    k = 2 * np.pi / 1.55
    A0, B0 = ws.construct_AB_vec(copy.deepcopy(mesh), mg.FIBER_IOR, k, sparse=True)
    A1, B1 = ws.construct_AB_vec(back, mg.FIBER_IOR, k, sparse=True)
    for M0, M1 in ((A0, A1), (B0, B1)):
        assert np.array_equal(M0.indptr, M1.indptr) and np.array_equal(M0.indices, M1.indices)
        assert np.array_equal(M0.data, M1.data)
    w0, v0, n0 = ws.solve_waveguide_vec(mesh, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False)
    w1, v1, n1 = ws.solve_waveguide_vec(back, 1.55, mg.FIBER_IOR, Nmax=4, verbose=False)
    assert n0 == n1 and np.allclose(np.sort(w0.real), np.sort(w1.real), rtol=1e-12, atol=0)


def test_scalar_problem_from_container(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200 as ws
    mg = ws.meshgen
    mesh = mg.lantern_cross_section(3000, seed=2)
    path = str(tmp_path / "lantern.npz")
    ws.save_mesh_npz(path, mesh, compressed=True)
    back = ws.load_mesh_npz(path)
    k = 2 * np.pi / 1.55
    A0, B0 = ws.construct_AB(mesh, mg.LANTERN_IOR, k, sparse=True)
    A1, B1 = ws.construct_AB(back, mg.LANTERN_IOR, k, sparse=True)
    assert np.array_equal(A0.indices, A1.indices) and np.array_equal(A0.data, A1.data) and np.array_equal(B0.data, B1.data)
