"""Drop-in for the hot path of reference ``fe_solver.py``: ``construct_AB``,
``construct_B``, ``construct_AB_vec``, ``solve_waveguide``, ``solve_waveguide_vec``,
``get_eff_index`` -- same signatures, argument meaning, return layout and exceptions, with
the numerics running on the B200 through the C ABI (``include/wavesolve_b200.h``).

Host work that remains here is what the reference also does on the host: reading the mesh
object, expanding ``cell_sets`` + ``IOR_dict`` into the material-major element order
(fe_solver.py:36, :48, :179) and wrapping results as scipy matrices / numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, device as _dev


# ---------------------------------------------------------------------------------------
# mesh marshalling (host)
# ---------------------------------------------------------------------------------------
def _material_major(mesh, IOR_dict, k, data=None):
    """Connectivity in the reference's assembly order and k^2 n^2 per element
    (fe_solver.py:36, :51 / :179-181)."""
    data = np.asarray(mesh.cells[1].data if data is None else data)
    conns, k2 = [], []
    for material in mesh.cell_sets.keys():
        idx = mesh.cell_sets[material][1]
        c = data[tuple(mesh.cell_sets[material])][0, :, 0, :] if idx is not None else data
        conns.append(c)
        k2.append(np.full(len(c), k ** 2 * IOR_dict[material] ** 2, dtype=np.float64))
    if conns:
        return np.vstack(conns), np.concatenate(k2)
    return data[:0], np.zeros(0)


def _as_i32(a, what):
    a = np.asarray(a)
    if a.size and (int(a.max()) >= 2 ** 31 or int(a.min()) < 0):
        raise ValueError("%s must be non-negative and fit in int32" % what)
    return np.ascontiguousarray(a.astype(np.int32))


class DevicePattern:
    """Owns a ``ws_pattern`` handle."""

    def __init__(self, dofs_d: torch.Tensor, N: int):
        self.dev = dofs_d.device
        self.dofs = dofs_d
        h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.ws_pattern_create(_dev.ptr(dofs_d), dofs_d.shape[0], int(N), C.byref(h),
                                       self.dev.index, _dev.stream_ptr(self.dev)))
        self.handle = h
        n, nnz, ne = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(L.ws_pattern_query(h, C.byref(n), C.byref(nnz), C.byref(ne)))
        self.N, self.nnz, self.Ne = n.value, nnz.value, ne.value

    def export(self):
        indptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.dev)
        indices = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
        _lib.check(_lib.lib().ws_pattern_export(self.handle, _dev.ptr(indptr), _dev.ptr(indices),
                                                self.dev.index, _dev.stream_ptr(self.dev)))
        return indptr, indices

    def prune(self, vals: torch.Tensor):
        """(indptr, indices, data) with exact zeros removed (lil_matrix semantics)."""
        indptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.dev)
        indices = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
        data = torch.empty(self.nnz, dtype=torch.float64, device=self.dev)
        n = C.c_int64()
        _lib.check(_lib.lib().ws_prune_zeros(self.handle, _dev.ptr(vals), _dev.ptr(indptr),
                                             _dev.ptr(indices), _dev.ptr(data), C.byref(n),
                                             self.dev.index, _dev.stream_ptr(self.dev)))
        return indptr, indices[: n.value], data[: n.value]

    def spmv(self, vals, x, out=None):
        if out is None:
            out = torch.empty(self.N, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_spmv(self.handle, _dev.ptr(vals), _dev.ptr(x), _dev.ptr(out),
                                      self.dev.index, _dev.stream_ptr(self.dev)))
        return out

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().ws_pattern_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ScalarProblem:
    """Device-resident scalar P2 problem: mesh arrays, pattern, A and B values."""

    def __init__(self, mesh, IOR_dict, k, device=None, all_cells_mass_only=False):
        self.dev = dev = _dev.require_cuda(device)
        if all_cells_mass_only:
            conn = np.asarray(mesh.cells[1].data)
            k2 = np.zeros(len(conn))
        else:
            conn, k2 = _material_major(mesh, IOR_dict, k)
        conn = _as_i32(conn, "node ids")
        # csr_matrix((data,(row,col))) infers the shape from the largest index (fe_solver.py:67)
        self.N = int(conn.max()) + 1 if conn.size else 0
        xy = np.ascontiguousarray(np.asarray(mesh.points, dtype=np.float64)[:, :2])
        self.h_dofs, self.h_xy, self.first_class_start = conn, xy, 0
        with torch.cuda.device(dev):
            self.xy = _dev.h2d(xy, dev)
            self.dofs = _dev.h2d(conn, dev)
            self.k2n2 = _dev.h2d(k2, dev)
            self.pattern = DevicePattern(self.dofs, self.N)
        self.h2d_bytes = xy.nbytes + conn.nbytes + k2.nbytes

    def assemble(self):
        p = self.pattern
        A = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        B = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_assemble_scalar_p2(p.handle, _dev.ptr(self.xy), _dev.ptr(self.dofs),
                                                    _dev.ptr(self.k2n2), _dev.ptr(A), _dev.ptr(B), 0,
                                                    self.dev.index, _dev.stream_ptr(self.dev)))
        return A, B

    def centroids(self):
        return _element_centroids(self.h_xy, self.h_dofs[:, :3])

    def assemble_mass(self):
        p = self.pattern
        B = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_assemble_mass_p2(p.handle, _dev.ptr(self.xy), _dev.ptr(self.dofs),
                                                  _dev.ptr(B), self.dev.index, _dev.stream_ptr(self.dev)))
        return B


def _element_centroids(h_xy, vert_ids):
    return np.ascontiguousarray(h_xy[vert_ids.astype(np.int64)].mean(axis=1))


class VectorProblem:
    """Device-resident vector (Whitney + P1) problem."""

    def __init__(self, mesh, IOR_dict, k, device=None):
        self.dev = dev = _dev.require_cuda(device)
        tris, k2 = _material_major(mesh, IOR_dict, k)
        eidx, _ = _material_major(mesh, IOR_dict, k, data=np.asarray(mesh.edge_indices))
        tris = _as_i32(tris, "vertex ids")
        eidx = _as_i32(eidx, "edge ids")
        self.Ntt = len(mesh.cells[0].data)                        # fe_solver.py:159-162
        self.Nzz = len(mesh.points)
        self.N = self.Ntt + self.Nzz
        dofs = np.hstack([eidx, tris + np.int32(self.Ntt)]).astype(np.int32)
        xy = np.ascontiguousarray(np.asarray(mesh.points, dtype=np.float64)[:, :2])
        self.h_dofs, self.h_xy, self.first_class_start = dofs, xy, self.Ntt
        self.h_tris = tris
        self.h_edges = _as_i32(np.asarray(mesh.cells[0].data), "edge table")
        with torch.cuda.device(dev):
            self.xy = _dev.h2d(xy, dev)
            self.tris = _dev.h2d(tris, dev)
            self.dofs = _dev.h2d(dofs, dev)
            self.k2n2 = _dev.h2d(k2, dev)
            self.edges = _dev.h2d(self.h_edges, dev)
            self.pattern = DevicePattern(self.dofs, self.N)
        self.h2d_bytes = xy.nbytes + tris.nbytes + dofs.nbytes + k2.nbytes + self.h_edges.nbytes

    def assemble(self, transpose=False):
        p = self.pattern
        A = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        B = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_assemble_vector_p1(p.handle, _dev.ptr(self.xy), _dev.ptr(self.tris),
                                                    _dev.ptr(self.k2n2), self.Ntt, _dev.ptr(A),
                                                    _dev.ptr(B), 1 if transpose else 0,
                                                    self.dev.index, _dev.stream_ptr(self.dev)))
        return A, B

    def assemble_qd(self, sigma):
        """Quasi-definite shift-invert matrix M' = T^T (A - sigma B) T on the pattern."""
        p = self.pattern
        M = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_assemble_vector_qd(p.handle, _dev.ptr(self.xy), _dev.ptr(self.tris),
                                                    _dev.ptr(self.k2n2), self.Ntt, float(sigma), _dev.ptr(M),
                                                    self.dev.index, _dev.stream_ptr(self.dev)))
        return M

    def apply_T(self, v, transpose=False):
        out = torch.empty_like(v)
        _lib.check(_lib.lib().ws_vector_T(self.pattern.handle, _dev.ptr(self.edges), _dev.ptr(self.xy), self.Ntt,
                                          1 if transpose else 0, _dev.ptr(v), _dev.ptr(out),
                                          self.dev.index, _dev.stream_ptr(self.dev)))
        return out

    def centroids(self):
        return _element_centroids(self.h_xy, self.h_tris)


def _to_scipy(cls, N, indptr, indices, data):
    return cls((_dev.d2h(data), _dev.d2h(indices), _dev.d2h(indptr)), shape=(N, N))


# ---------------------------------------------------------------------------------------
# reference API: assembly
# ---------------------------------------------------------------------------------------
def construct_AB_order2(mesh, IOR_dict, k, sparse=False, poke_index=None, device=None):
    """fe_solver.py:20-69.  ``sparse=True`` returns two CSR matrices sharing one pattern
    (explicit zeros kept, int32 indices, shape inferred from the largest node id).
    ``sparse=False`` raises UnboundLocalError in the reference (A_data undefined, fe_solver.py:53);
    here it returns the dense arrays the docstring promises."""
    if poke_index is not None:
        raise NotImplementedError("poke_index is broken in the reference (fe_solver.py:58-64) and not supported")
    prob = ScalarProblem(mesh, IOR_dict, k, device)
    with torch.cuda.device(prob.dev):
        A, B = prob.assemble()
        indptr, indices = prob.pattern.export()
        Am = _to_scipy(sp.csr_matrix, prob.N, indptr, indices, A)
        Bm = sp.csr_matrix((_dev.d2h(B), Am.indices.copy(), Am.indptr.copy()), shape=Am.shape)
    if not sparse:
        return Am.toarray(), Bm.toarray()
    return Am, Bm


def construct_B_order2(mesh, sparse=False, device=None):
    """fe_solver.py:71-99: mass matrix over all cells (materials ignored)."""
    prob = ScalarProblem(mesh, None, 0.0, device, all_cells_mass_only=True)
    with torch.cuda.device(prob.dev):
        B = prob.assemble_mass()
        indptr, indices = prob.pattern.export()
        Bm = _to_scipy(sp.csr_matrix, prob.N, indptr, indices, B)
    if not sparse:
        N = len(mesh.points)                                      # fe_solver.py:91-92
        out = np.zeros((N, N))
        out[: Bm.shape[0], : Bm.shape[1]] = Bm.toarray()
        return out
    return Bm


construct_B = construct_B_order2                                   # fe_solver.py:101


def construct_AB(mesh, IOR_dict, k, sparse=False, order=2, device=None):
    """fe_solver.py:131-146."""
    if order == 2:
        assert mesh.cells[1].data.shape[1] == 6, "must use order 2 mesh for order 2 solver."
        return construct_AB_order2(mesh, IOR_dict, k, sparse, device=device)
    elif order == 1:
        assert mesh.cells[1].data.shape[1] == 3, "must use order 1 mesh for order 1 solver"
        raise NotImplementedError("order=1 scalar assembly (fe_solver.py:103-129) is not on the accelerated path yet")
    else:
        raise NotImplementedError


def construct_AB_vec(mesh, IOR_dict, k, sparse=False, device=None):
    """fe_solver.py:148-214.  Unknowns: edges [0,Ntt) then nodes [Ntt,Ntt+Nzz).
    ``sparse=True`` returns CSC matrices without stored zeros (what lil_matrix -> tocsc gives);
    the values are assembled on the device in the reference's element order so that entries
    cancelling to exactly 0.0 are the same ones."""
    prob = VectorProblem(mesh, IOR_dict, k, device)
    with torch.cuda.device(prob.dev):
        # transpose=True: slot (i,j) of the symmetric pattern receives entry (j,i), i.e. the
        # CSR arrays of the transposed matrix == the CSC arrays of the matrix itself.
        A, B = prob.assemble(transpose=True)
        Ap = prob.pattern.prune(A)
        Bp = prob.pattern.prune(B)
        Am = _to_scipy(sp.csc_matrix, prob.N, *Ap)
        Bm = _to_scipy(sp.csc_matrix, prob.N, *Bp)
    if not sparse:
        return Am.toarray(), Bm.toarray()
    return Am, Bm


def get_eff_index(wl, w):
    """fe_solver.py:431-434."""
    k = 2 * np.pi / wl
    return np.sqrt(w / k ** 2)
