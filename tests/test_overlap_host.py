"""CPU: the host-side parts of wavesolve_b200.overlap -- the guided-mode count (fe_solver.py:330-346)
against the counts the literal reference returned (tests/golden), and the oracle's overlap expression
against B-orthonormality of the reference's own eigenvectors."""
import numpy as np
import pytest

from oracle import restated as R
from tests import util


@pytest.mark.parametrize("name", ["scalar_sq8", "scalar_disc600", "scalar_disc600_target", "scalar_c1_disc5k",
                                  "scalar_lantern2k"])
def test_count_modes_matches_reference_counts(name):
    import wavesolve_b200 as ws
    mesh, ior, z = util.load_golden(name)
    w = z["w"]                                  # descending, as solve_waveguide returns them
    assert ws.count_modes(w, util.WL, ior) == int(z["count"])
    assert ws.count_modes(w, util.WL, ior) == R.count_modes_scalar(w, 2 * np.pi / util.WL, ior)


def test_count_modes_edge_cases():
    import wavesolve_b200 as ws
    ior = {"core": 1.46, "cladding": 1.444}
    k = 2 * np.pi / util.WL
    inside = (k * 1.45) ** 2
    assert ws.count_modes([], util.WL, ior) == 0
    assert ws.count_modes([-1.0, inside, inside], util.WL, ior) == 2       # negative eigenvalues are skipped
    assert ws.count_modes([inside, (k * 1.40) ** 2, inside], util.WL, ior) == 1   # stops at the first outsider
    assert ws.count_modes([(k * 1.50) ** 2, inside], util.WL, ior) == 0


def test_oracle_overlap_of_reference_modes():
    mesh, ior, z = util.load_golden("scalar_disc600")
    B = util.golden_matrix(z, "B")
    v = z["v"]
    G = R.overlap_matrix(v, v, B)
    assert np.allclose(G, np.eye(len(v)), atol=1e-8)
    assert abs(R.overlap_matrix(v[0], v[0], B) - 1.0) < 1e-8
    U = v[:2] + 1j * v[2:4]
    assert np.allclose(R.overlap_matrix(U, U, B, conj=True), 2 * np.eye(2), atol=1e-8)
    assert np.allclose(R.normalize_modes(3.0 * v[:3], B), v[:3], atol=1e-12)
