"""wavesolve_b200 -- B200-native FEM hot path of jw-lin/wavesolve (assembly + generalized
sparse eigensolve), drop-in for ``wavesolve.fe_solver`` / ``wavesolve.waveguide`` hot-path
functions.  The CUDA library is required; there is no CPU fallback."""
from . import fe_solver, waveguide, meshgen, sweep, interp, mesh_io, overlap as _overlap_mod  # noqa: F401
from .fe_solver import (construct_AB, construct_AB_vec, construct_B, get_eff_index,  # noqa: F401
                        solve, solve_waveguide, solve_waveguide_vec, solve_sparse, compute_IOR_arr)
from .waveguide import get_unique_edges  # noqa: F401
from .interp import (construct_meshtree, find_triangle, find_triangle_KDtree, get_tri_idxs,  # noqa: F401
                     get_tri_idxs_KDtree, get_interp_weights, get_interp_weights_vec, interpolate,
                     interpolate_vec, unstructured_interpolate, mesh_interpolate)
from .overlap import (MassOperator, mass_operator, overlap, overlap_matrix, normalize_modes,  # noqa: F401
                      count_modes)
from .mesh_io import load_meshio_mesh, load_mesh_npz, save_mesh_npz  # noqa: F401

__version__ = "0.2.0"
