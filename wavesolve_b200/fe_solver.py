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


def _material_order(mesh, IOR_dict, k):
    """Element order of the reference's assembly loop (material by material in ``cell_sets``
    insertion order, fe_solver.py:36 / :179) as one index array into ``cells[1].data`` -- None when
    that is the identity -- and k^2 n^2 per element in that order."""
    idxs, k2 = [], []
    ne = len(mesh.cells[1].data)
    for material in mesh.cell_sets.keys():
        idx = mesh.cell_sets[material][1]
        idx = np.arange(ne, dtype=np.int64) if idx is None else np.asarray(idx).astype(np.int64, copy=False).ravel()
        idxs.append(idx)
        k2.append(np.full(len(idx), k ** 2 * IOR_dict[material] ** 2, dtype=np.float64))
    if not idxs:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    order = idxs[0] if len(idxs) == 1 else np.concatenate(idxs)
    if len(order) == ne and ne and order[0] == 0 and order[-1] == ne - 1 and np.array_equal(order, np.arange(ne)):
        return None, np.concatenate(k2)          # identity: nothing to gather, no bounds check needed
    lo, hi = (int(x) for x in torch.from_numpy(order).aminmax()) if order.size and order.flags.writeable else \
        ((int(order.min()), int(order.max())) if order.size else (0, 0))
    if order.size and (lo < 0 or hi >= ne):
        # numpy fancy indexing semantics of the reference: negative ids wrap, too large ids raise
        if hi >= ne or lo < -ne:
            raise IndexError("cell_sets index out of bounds for cells[1].data")
        order = np.where(order < 0, order + ne, order)
    return order, np.concatenate(k2)


def _k2n2_host(mesh, IOR_dict, k):
    """k^2 n^2 per element in material-major order (fe_solver.py:51 / :181) -- the only input
    that changes between the wavelengths / index sets of a sweep on one mesh."""
    out = []
    for material in mesh.cell_sets.keys():
        idx = mesh.cell_sets[material][1]
        n = len(idx) if idx is not None else len(mesh.cells[1].data)
        out.append(np.full(n, k ** 2 * IOR_dict[material] ** 2, dtype=np.float64))
    return np.concatenate(out) if out else np.zeros(0)


def _as_i32(a, what):
    """Integer ids -> contiguous int32 (validated: non-negative, below 2^31).  64-bit contiguous arrays
    (what pygmsh / get_unique_edges deliver) go through torch's multi-threaded CPU kernels when the
    process has more than one thread: 0.8 ms instead of 6 ms per 3M ids."""
    a = np.asarray(a)
    if a.size and a.dtype.kind in "iu" and a.dtype.itemsize == 8 and a.flags.c_contiguous and a.flags.writeable \
            and torch.get_num_threads() > 1:
        t = torch.from_numpy(a.view(np.int64))          # unsigned ids >= 2^63 show up as negative
        lo, hi = t.aminmax()
        if int(lo) < 0 or int(hi) >= 2 ** 31:
            raise ValueError("%s must be non-negative integers that fit in int32" % what)
        return t.to(torch.int32).numpy()
    if a.size:
        # one pass: seen as unsigned bit patterns of the same width, negative ids are huge
        bits = a.view(np.dtype("u%d" % a.dtype.itemsize)) if a.dtype.kind == "i" and a.flags.c_contiguous else a
        if a.dtype.kind not in "iu" or int(bits.max()) >= 2 ** 31 or (bits is a and a.dtype.kind == "i" and int(a.min()) < 0):
            raise ValueError("%s must be non-negative integers that fit in int32" % what)
    return np.ascontiguousarray(a.astype(np.int32))


def _ids_to_device(a, what, dev, bound=None):
    """Integer ids (any width, as pygmsh / get_unique_edges deliver them) -> contiguous int32 CUDA tensor,
    validated (non-negative, < 2^31, < ``bound``).  With more than one CPU thread the 64-bit ids are narrowed by
    torch's threaded CPU kernels before the copy (half the PCIe bytes); in a single-threaded process -- every
    rank under torchrun, which sets OMP_NUM_THREADS=1 -- the raw ids are copied and narrowed / validated on the
    device instead (numpy's single-threaded astype costs 6 ms per 3M ids, three times per problem)."""
    a = np.asarray(a)
    if a.size and a.dtype.kind in "iu" and a.dtype.itemsize == 8 and a.flags.c_contiguous and a.flags.writeable \
            and torch.get_num_threads() == 1:
        t = torch.from_numpy(a.view(np.int64)).to(dev)          # unsigned ids >= 2^63 show up as negative
        lo, hi = (int(x) for x in torch.stack(t.aminmax()).tolist())
        if lo < 0 or hi >= 2 ** 31:
            raise ValueError("%s must be non-negative integers that fit in int32" % what)
        if bound is not None and hi >= bound:
            raise IndexError("%s: id %d out of bounds for size %d" % (what, hi, bound))
        return t.to(torch.int32), a.size * 8
    h = _as_i32(a, what)
    if bound is not None and h.size and int(_imax(h)) >= bound:
        raise IndexError("%s: id %d out of bounds for size %d" % (what, int(_imax(h)), bound))
    return _dev.h2d(h, dev), h.nbytes


def _imax(a):
    """max of an int32 array (torch's threaded reduction for the large ones)."""
    if a.size > (1 << 16) and a.flags.c_contiguous and a.flags.writeable:
        return torch.from_numpy(a).max()
    return a.max() if a.size else 0


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

    def __init__(self, mesh, IOR_dict, k, device=None, all_cells_mass_only=False, build_pattern=True, order=2):
        self.dev = dev = _dev.require_cuda(device)
        self.order = order
        if all_cells_mass_only:
            conn = np.asarray(mesh.cells[1].data)
            k2 = np.zeros(len(conn))
        else:
            conn, k2 = _material_major(mesh, IOR_dict, k)
        conn = _as_i32(conn, "node ids")
        # the device kernels index xy[] with these ids unchecked: refuse what numpy would refuse (IndexError at
        # points[tris], shape_funcs.py:196) and the wrong element order (fe_solver.py:140, :143)
        width = 6 if order == 2 else 3
        if conn.ndim != 2 or conn.shape[1] < width:
            raise AssertionError("must use order %d mesh for order %d solver" % (order, order))
        if conn.size and int(conn.max()) >= len(mesh.points):
            raise IndexError("node id %d out of bounds for %d mesh points" % (int(conn.max()), len(mesh.points)))
        if order == 1:
            # linear triangles ride the six-unknown pattern machinery as (v0 v1 v2 v0 v1 v2)
            conn = np.ascontiguousarray(np.hstack([conn[:, :3], conn[:, :3]]))
            self.N = len(mesh.points)                             # fe_solver.py:107
        else:
            # csr_matrix((data,(row,col))) infers the shape from the largest index (fe_solver.py:67)
            self.N = int(conn.max()) + 1 if conn.size else 0
        xy = np.ascontiguousarray(np.asarray(mesh.points, dtype=np.float64)[:, :2])
        self.h_dofs, self.h_xy, self.first_class_start = conn, xy, 0
        with torch.cuda.device(dev):
            self.xy = _dev.h2d(xy, dev)
            self.dofs = _dev.h2d(conn, dev)
            self.k2n2 = _dev.h2d(k2, dev)
            self.pattern = DevicePattern(self.dofs, self.N) if build_pattern else None
        self.h2d_bytes = xy.nbytes + conn.nbytes + k2.nbytes

    def build_pattern(self):
        if self.pattern is None:
            with torch.cuda.device(self.dev):
                self.pattern = DevicePattern(self.dofs, self.N)
        return self.pattern

    def set_materials(self, mesh, IOR_dict, k):
        """New wavelength / index set on the same mesh (K11): only k^2 n^2 per element changes;
        pattern, incidence lists and symbolic analysis are kept."""
        k2 = _k2n2_host(mesh, IOR_dict, k)
        assert k2.shape[0] == self.k2n2.shape[0], "mesh does not match this problem"
        with torch.cuda.device(self.dev):
            self.k2n2.copy_(torch.from_numpy(k2), non_blocking=False)

    def assemble(self, fast=False, out=None):
        """A, B values on the pattern.  fast=True: reciprocal arithmetic (<= 2 ulp off the
        reference's), used inside the solve pipeline only.  ``out=(A, B)``: reuse value arrays
        (re-assembly per wavelength on a shared pattern, K11)."""
        p = self.build_pattern()
        A, B = out if out is not None else (torch.empty(p.nnz, dtype=torch.float64, device=self.dev),
                                            torch.empty(p.nnz, dtype=torch.float64, device=self.dev))
        assert A.numel() == p.nnz and B.numel() == p.nnz and A.is_contiguous() and B.is_contiguous()
        if self.order == 1:
            _lib.check(_lib.lib().ws_assemble_scalar_p1(p.handle, _dev.ptr(self.xy), _dev.ptr(self.dofs),
                                                        _dev.ptr(self.k2n2), _dev.ptr(A), _dev.ptr(B),
                                                        self.dev.index, _dev.stream_ptr(self.dev)))
            return A, B
        _lib.check(_lib.lib().ws_assemble_scalar_p2(p.handle, _dev.ptr(self.xy), _dev.ptr(self.dofs),
                                                    _dev.ptr(self.k2n2), _dev.ptr(A), _dev.ptr(B),
                                                    _lib.ASM_FAST if fast else 0,
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

    def fresh_pattern(self):
        """A new problem object that shares the resident input arrays (coordinates,
        connectivity, material data) but rebuilds the pattern -- bench.py's device-resident step."""
        o = object.__new__(VectorProblem)
        o.__dict__.update(self.__dict__)
        o.pattern = None
        return o

    def build_pattern(self):
        if self.pattern is None:
            with torch.cuda.device(self.dev):
                self.pattern = DevicePattern(self.dofs, self.N)
        return self.pattern

    set_materials = ScalarProblem.set_materials

    def __init__(self, mesh, IOR_dict, k, device=None, build_pattern=True):
        self.dev = dev = _dev.require_cuda(device)
        tris_all = np.asarray(mesh.cells[1].data)
        eidx_all = np.asarray(mesh.edge_indices)
        assert len(eidx_all) == len(tris_all), "edge_indices must have one row per triangle (run get_unique_edges)"
        order, k2 = _material_order(mesh, IOR_dict, k)
        self.Ntt = len(mesh.cells[0].data)                        # fe_solver.py:159-162
        self.Nzz = len(mesh.points)
        self.N = self.Ntt + self.Nzz
        self.first_class_start = self.Ntt
        tris_h = tris_all if tris_all.shape[1] == 3 else tris_all[:, :3]
        pts = np.asarray(mesh.points)
        whole = pts.dtype == np.float64 and pts.ndim == 2 and pts.shape[1] > 2 and pts.flags.c_contiguous
        # (Np, 3) float64 as pygmsh delivers it: the z column is dropped on the device (a strided host copy of
        # the first two columns costs more than moving the third one)
        xy = pts if whole else np.ascontiguousarray(np.asarray(pts, dtype=np.float64)[:, :2])
        with torch.cuda.device(dev):
            self.xy = _dev.h2d(xy, dev)[:, :2].contiguous() if whole else _dev.h2d(xy, dev)
            # ids are used unchecked by the kernels: validated here (numpy raises IndexError for these in the reference)
            self.edges, nb_e = _ids_to_device(np.asarray(mesh.cells[0].data), "edge table", dev, self.Nzz)
            self.k2n2 = _dev.h2d(k2, dev)
            t_d, nb_t = _ids_to_device(tris_h, "vertex ids", dev, self.Nzz)
            e_d, nb_i = _ids_to_device(eidx_all, "edge ids (stale edge_indices? run get_unique_edges)", dev, self.Ntt)
            if order is not None:
                # material-major element order (fe_solver.py:179) gathered on the device
                o_d = _dev.h2d(order, dev)
                t_d, e_d = t_d.index_select(0, o_d), e_d.index_select(0, o_d)
            self.tris = t_d.contiguous()
            self.dofs = torch.cat([e_d, self.tris + self.Ntt], dim=1).contiguous()   # edges first, then nodes
            self.pattern = DevicePattern(self.dofs, self.N) if build_pattern else None
        self.h2d_bytes = (xy.nbytes + nb_t + nb_i + k2.nbytes + nb_e + (order.nbytes if order is not None else 0))

    # host copies of the resident arrays (tests and diagnostics; the solve path does not need them)
    @property
    def h_dofs(self):
        return _dev.d2h(self.dofs)

    @property
    def h_xy(self):
        return _dev.d2h(self.xy)

    @property
    def h_tris(self):
        return _dev.d2h(self.tris)

    @property
    def h_edges(self):
        return _dev.d2h(self.edges)

    def assemble(self, transpose=False, fast=False, out=None):
        p = self.build_pattern()
        A, B = out if out is not None else (torch.empty(p.nnz, dtype=torch.float64, device=self.dev),
                                            torch.empty(p.nnz, dtype=torch.float64, device=self.dev))
        assert A.numel() == p.nnz and B.numel() == p.nnz and A.is_contiguous() and B.is_contiguous()
        _lib.check(_lib.lib().ws_assemble_vector_p1(p.handle, _dev.ptr(self.xy), _dev.ptr(self.tris),
                                                    _dev.ptr(self.k2n2), self.Ntt, _dev.ptr(A),
                                                    _dev.ptr(B),
                                                    (_lib.ASM_TRANSPOSE if transpose else 0) | (_lib.ASM_FAST if fast else 0),
                                                    self.dev.index, _dev.stream_ptr(self.dev)))
        return A, B

    def assemble_qd(self, sigma, with_B=False, fast=False):
        """Quasi-definite shift-invert matrix M' = T^T (A - sigma B) T on the pattern (and B in the
        same pass when with_B)."""
        p = self.build_pattern()
        M = torch.empty(p.nnz, dtype=torch.float64, device=self.dev)
        B = torch.empty(p.nnz, dtype=torch.float64, device=self.dev) if with_B else None
        _lib.check(_lib.lib().ws_assemble_vector_qd(p.handle, _dev.ptr(self.xy), _dev.ptr(self.tris),
                                                    _dev.ptr(self.k2n2), self.Ntt, float(sigma), _dev.ptr(M),
                                                    _dev.ptr(B), _lib.ASM_FAST if fast else 0,
                                                    self.dev.index, _dev.stream_ptr(self.dev)))
        return (M, B) if with_B else M

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


def construct_AB_order1(mesh, IOR_dict, k, sparse=False, device=None):
    """fe_solver.py:103-129: linear-triangle scalar assembly.  ``sparse=True`` returns two
    lil matrices of shape (len(points),)*2 (exact zeros not stored), ``sparse=False`` dense arrays."""
    prob = ScalarProblem(mesh, IOR_dict, k, device, order=1)
    with torch.cuda.device(prob.dev):
        A, B = prob.assemble()
        Am = _to_scipy(sp.csr_matrix, prob.N, *prob.pattern.prune(A))
        Bm = _to_scipy(sp.csr_matrix, prob.N, *prob.pattern.prune(B))
    if not sparse:
        return Am.toarray(), Bm.toarray()
    return Am.tolil(), Bm.tolil()


def construct_AB(mesh, IOR_dict, k, sparse=False, order=2, device=None):
    """fe_solver.py:131-146."""
    if order == 2:
        assert mesh.cells[1].data.shape[1] == 6, "must use order 2 mesh for order 2 solver."
        return construct_AB_order2(mesh, IOR_dict, k, sparse, device=device)
    elif order == 1:
        assert mesh.cells[1].data.shape[1] == 3, "must use order 1 mesh for order 1 solver"
        return construct_AB_order1(mesh, IOR_dict, k, sparse, device=device)
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


# ---------------------------------------------------------------------------------------
# reference API: eigen-solves
# ---------------------------------------------------------------------------------------
last_info = {}      # diagnostics of the most recent solve_waveguide / solve_waveguide_vec call


def _eig_device(pattern, solver, B_vals, sigma, kind, nev, N, dev, ncv=0, tol=0.0, maxit=0, seed=0,
                edges=None, xy=None, Ntt=0, out=None):
    """``out=(w, V)``: caller-provided device tensors (nev,) / (nev, N) -- in a sweep these are
    views of the rank's NCCL send buffer, so the Ritz-vector kernel (K10) writes the gather
    payload directly (K12)."""
    if out is None:
        w = torch.empty(nev, dtype=torch.float64, device=dev)
        V = torch.empty((nev, N), dtype=torch.float64, device=dev)
    else:
        w, V = out
        assert w.is_contiguous() and V.is_contiguous() and w.numel() == nev and V.numel() == nev * N
    info = np.zeros(5, dtype=np.int64)
    resid = np.zeros(nev, dtype=np.float64)
    rc = _lib.lib().ws_eig_solve(pattern.handle, solver.handle, _dev.ptr(B_vals), float(sigma), int(kind),
                                 int(nev), int(ncv), float(tol), int(maxit), int(seed),
                                 _dev.ptr(edges), _dev.ptr(xy), int(Ntt), _dev.ptr(w), _dev.ptr(V),
                                 info.ctypes.data, resid.ctypes.data, dev.index, _dev.stream_ptr(dev))
    out_info = dict(nconv=int(info[0]), n_op=int(info[1]), restarts=int(info[2]), ncomplex=int(info[3]), resid=resid)
    try:
        _lib.check(rc)
    except _lib.WavesolveB200Error as e:
        if e.rc == _lib.ERR_NOCONV:
            # like scipy's ArpackNoConvergence, the partial results travel with the exception
            e.eigenvalues, e.eigenvectors, e.info = w, V, out_info
        raise
    return w, V, out_info


def _count_modes(w_iter, k, IOR_dict, always=False):
    """fe_solver.py:330-346 / 409-424."""
    IORs = [ior[1] for ior in IOR_dict.items()]
    nmin, nmax = min(IORs), max(IORs)
    mode_count = 0
    for _w in w_iter:
        if _w < 0:
            continue
        ne = np.sqrt(_w / k ** 2)
        if (nmin <= ne <= nmax) or always:
            mode_count += 1
        else:
            break
    return mode_count


def solve_waveguide(mesh, wl, IOR_dict, plot=False, ignore_warning=False, sparse=True, Nmax=10, order=2,
                    target_neff=None, device=None, leaf_elems=None):
    """Scalar modes (fe_solver.py:292-348).  Returns ``(w, v, N)``: eigenvalues in descending
    order, eigenvectors as rows (B-orthonormal, like eigsh), number of guided modes.

    The whole pipeline runs on the device: assembly -> multifrontal LDL^T of A - sigma B ->
    B-orthogonal thick-restart Lanczos on (A - sigma B)^-1 B (the algorithm eigsh/ARPACK mode 3
    runs on the host at fe_solver.py:328) with scipy's tol = eps and ncv = max(2 Nmax + 4, 20) (scipy: 2 Nmax + 1; see csrc/ws_eig.cu).
    ``sparse=False`` (LAPACK eigh on dense matrices in the reference, which cannot even assemble
    them for order 2) is served by the same path with the shift at the top of the spectrum."""
    k = 2 * np.pi / wl
    if target_neff is None or not sparse:
        est_eigval = np.power(k * max(IOR_dict.values()), 2)
    else:
        est_eigval = np.power(k * target_neff, 2)
    if order == 2:
        assert mesh.cells[1].data.shape[1] == 6, "must use order 2 mesh for order 2 solver."
    elif order == 1:
        assert mesh.cells[1].data.shape[1] == 3, "must use order 1 mesh for order 1 solver"
    else:
        raise NotImplementedError
    prob = ScalarProblem(mesh, IOR_dict, k, device, build_pattern=False, order=order)
    N = prob.N
    if N > 2000 and not ignore_warning and not sparse:
        raise Exception("A and B matrices are larger than 2000 x 2000 - this may make your system unstable. consider setting sparse=True")
    with torch.cuda.device(prob.dev):
        wd, Vd, info = solve_scalar_resident(prob, est_eigval, Nmax, leaf_elems)
        w = _dev.d2h(wd)
        v = _dev.d2h_pinned(Vd)
    last_info.clear()
    last_info.update(info)
    mode_count = _count_modes(w, k, IOR_dict)
    if plot:
        _report_modes(wl, w, k, IOR_dict, warn_always=True)
    return w, v, mode_count


def _values_on_pattern(M, pattern, what):
    """Values of a scipy sparse matrix laid out on the slots of a device pattern.  A matrix with
    exactly the pattern's structure (what ``construct_AB`` returns, explicit zeros included) is
    uploaded as it is; any other matrix whose entries all lie inside the pattern (pruned zeros,
    lil -> csr of ``construct_AB_order1``) is scattered into place by a sorted-key lookup on the
    device; an entry outside the mesh's pattern is refused."""
    import scipy.sparse as sp
    if not sp.issparse(M):
        raise TypeError("%s: a scipy sparse matrix is expected (the dense `solve` of fe_solver.py:220 "
                        "is not part of the device path)" % what)
    M = M.tocsr()                                                  # fe_solver.py:266
    N, dev = pattern.N, pattern.dev
    if M.shape != (N, N):
        raise ValueError("%s has shape %s, the mesh gives %d unknowns" % (what, M.shape, N))
    if not M.has_canonical_format:
        M = M.copy()                                               # never touch the caller's matrix
        M.sum_duplicates()
    if np.iscomplexobj(M.data):
        raise TypeError("%s: real matrices only" % what)
    indptr = _dev.h2d(np.ascontiguousarray(M.indptr, dtype=np.int32), dev)
    indices = _dev.h2d(np.ascontiguousarray(M.indices, dtype=np.int32), dev)
    data = _dev.h2d(np.ascontiguousarray(M.data, dtype=np.float64), dev)
    p_indptr, p_indices = pattern.export()
    return _place_on_pattern(p_indptr, p_indices, indptr, indices, data, N, what)


def _place_on_pattern(p_indptr, p_indices, indptr, indices, data, N, what="matrix"):
    """CSR arrays (canonical: sorted, no duplicates) -> value array on the slots of the CSR pattern
    (p_indptr, p_indices).  Tensors of one device; index marshalling only (sorted (row, column) keys,
    one binary search per stored entry)."""
    nnz_p = p_indices.numel()
    if indices.numel() == nnz_p and torch.equal(indptr, p_indptr) and torch.equal(indices, p_indices):
        return data
    dev = p_indptr.device
    rows = torch.arange(N, dtype=torch.int64, device=dev)
    key_p = torch.repeat_interleave(rows, torch.diff(p_indptr).to(torch.int64)) * N + p_indices.to(torch.int64)
    key_m = torch.repeat_interleave(rows, torch.diff(indptr).to(torch.int64)) * N + indices.to(torch.int64)
    slot = torch.searchsorted(key_p, key_m)
    ok = slot < nnz_p
    if nnz_p:
        ok &= key_p[slot.clamp(max=nnz_p - 1)] == key_m
    if not bool(ok.all()):
        raise ValueError("%s has %d stored entries outside the sparsity pattern of the mesh"
                         % (what, int((~ok).sum())))
    vals = torch.zeros(nnz_p, dtype=torch.float64, device=dev)
    vals[slot] = data
    return vals


def solve_sparse(A, B, mesh, k, IOR_dict, plot=False, num_modes=6, device=None, leaf_elems=None):
    """The eigen-solve on matrices the caller already holds (fe_solver.py:262-290):
    ``eigsh(A.tocsr(), M=B.tocsr(), k=num_modes, which="SA", sigma=(k max n)^2)``, returned as
    ``(w descending, v rows, mode_count)``.  ``A`` / ``B``: scipy sparse matrices on the unknowns of
    ``mesh`` (``construct_AB(mesh, IOR_dict, k, sparse=True)`` of the reference or of this package, any
    format).  The values are placed on the mesh's device pattern and go through the same
    factorisation + Lanczos iteration as ``solve_waveguide``; the mesh supplies the coordinates the
    nested-dissection ordering works on (SuperLU orders the bare matrix with COLAMD)."""
    ncol = np.asarray(mesh.cells[1].data).shape[1]
    if ncol not in (3, 6):
        raise NotImplementedError("triangles with %d nodes" % ncol)
    order = 2 if ncol == 6 else 1
    est_eigval = np.power(k * max(IOR_dict.values()), 2)
    prob = ScalarProblem(mesh, IOR_dict, k, device, order=order)
    try:
        with torch.cuda.device(prob.dev):
            Av = _values_on_pattern(A, prob.pattern, "A")
            Bv = _values_on_pattern(B, prob.pattern, "B")
            wd, Vd, info = solve_scalar_resident(prob, est_eigval, num_modes, leaf_elems, values=(Av, Bv))
            w = _dev.d2h(wd)
            v = _dev.d2h_pinned(Vd)
    finally:
        prob.pattern.close()
    last_info.clear()
    last_info.update(info)
    mode_count = _count_modes(w, k, IOR_dict)
    if plot:
        _report_modes(2 * np.pi / k, w, k, IOR_dict, warn_always=True)
    return w, v, mode_count


def solve(A, B, mesh, k, IOR_dict, plot=False, device=None):
    """The legacy dense eigensolve (fe_solver.py:220-260): ALL eigenpairs of the dense pencil ``A v = w B v``
    (``scipy.linalg.eigh(A, B)`` in the reference), returned as ``(w descending, v rows, mode_count)``.

    Not part of the sparse hot path: a dense generalized symmetric eigenproblem for every eigenpair is a LAPACK job,
    and on the device it is a library job too -- Cholesky ``B = L L^T``, ``C = L^-1 A L^-T``, ``eigh(C)``,
    ``v = L^-T y`` through torch.linalg (cuSOLVER), so that the reference's full API is served from the GPU.  Like the
    reference it is only sensible for small N (the reference warns above 2000 x 2000)."""
    dev = _dev.require_cuda(device)
    if sp.issparse(A):
        A = A.toarray()
    if sp.issparse(B):
        B = B.toarray()
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1] or A.shape != B.shape:
        raise ValueError("A and B must be square matrices of the same shape")
    with torch.cuda.device(dev):
        Ad, Bd = _dev.h2d(A, dev), _dev.h2d(B, dev)
        Ld = torch.linalg.cholesky(Bd)                            # eigh(A, B) requires B positive definite as well
        Cd = torch.linalg.solve_triangular(Ld, Ad, upper=False)
        Cd = torch.linalg.solve_triangular(Ld, Cd.T.contiguous(), upper=False).T
        Cd = 0.5 * (Cd + Cd.T)
        wd, Yd = torch.linalg.eigh(Cd)
        Vd = torch.linalg.solve_triangular(Ld.T.contiguous(), Yd, upper=True)     # columns B-orthonormal, like scipy's eigh
        w = _dev.d2h(wd)[::-1]
        v = _dev.d2h(Vd).T[::-1]
    mode_count = _count_modes(w, k, IOR_dict)
    if plot:
        _report_modes(2 * np.pi / k, w, k, IOR_dict, warn_always=True)
    return w, v, mode_count


def compute_IOR_arr(mesh, IOR_dict):
    """Refractive index per element in ``cells[1].data`` order (fe_solver.py:446-452); elements
    in no material set stay 0, a later set overwrites an earlier one."""
    IORs = np.zeros(len(mesh.cells[1].data))
    for material in mesh.cell_sets.keys():
        IORs[tuple(mesh.cell_sets[material])] = IOR_dict[material]
    return IORs


def solve_vector_resident(prob, sigma, Nmax, leaf_elems=None, symbolic=None, solver=None, out=None):
    """Device part of the vector solve on a resident VectorProblem: assembly of B and of the
    quasi-definite M', symbolic analysis (host), factorisation, Krylov iteration.  Returns
    device tensors (w, V) and diagnostics.  ``solver``: a Solver already bound to this pattern
    (sweeps reuse the task plan and the front storage across wavelengths)."""
    from . import solver as _sv
    prob.build_pattern()
    Mq, B = prob.assemble_qd(sigma, with_B=True, fast=True)
    if solver is not None:
        S = solver
    else:
        if symbolic is None:
            symbolic = _sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt,
                                              leaf_elems or _sv.DEFAULT_LEAF_ELEMS)
        S = _sv.Solver(prob.pattern, symbolic)
    try:
        st = S.factor(Mq)
        wd, Vd, info = _eig_device(prob.pattern, S, B, sigma, 1, Nmax, prob.N, prob.dev,
                                   edges=prob.edges, xy=prob.xy, Ntt=prob.Ntt, out=out)
    finally:
        if solver is None:
            S.close()
    info.update(factor=st, N=prob.N, nnz=prob.pattern.nnz, sigma=float(sigma))
    return wd, Vd, info


def solve_scalar_resident(prob, sigma, Nmax, leaf_elems=None, symbolic=None, solver=None, out=None, values=None):
    """Device part of the scalar solve on a resident ScalarProblem (see solve_vector_resident).
    ``values=(A, B)``: device value arrays on the problem's pattern to use instead of assembling
    (``solve_sparse``: the caller brings the matrices)."""
    from . import solver as _sv
    prob.build_pattern()
    A, B = values if values is not None else prob.assemble(fast=True)
    if solver is not None:
        S = solver
    else:
        if symbolic is None:
            symbolic = _sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, 0, leaf_elems or _sv.DEFAULT_LEAF_ELEMS)
        S = _sv.Solver(prob.pattern, symbolic)
    try:
        st = S.factor(A, B, sigma)
        wd, Vd, info = _eig_device(prob.pattern, S, B, sigma, 0, Nmax, prob.N, prob.dev, out=out)
    finally:
        if solver is None:
            S.close()
    info.update(factor=st, N=prob.N, nnz=prob.pattern.nnz, sigma=float(sigma))
    return wd, Vd, info


def solve_waveguide_vec(mesh, wl, IOR_dict, plot=False, ignore_warning=False, sparse=True, Nmax=10,
                        target_neff=None, sparse_solve_mode='transform', verbose=True, device=None,
                        leaf_elems=None, real_output=False):
    """Vector modes with Whitney edge + P1 node elements (fe_solver.py:350-425).  Returns
    ``(w, v, N)``; ``v`` rows have unit 2-norm (edge unknowns first, then nodes).

    Every ``sparse_solve_mode`` is served by the same device path with the semantics of the
    reference's default 'transform' (eigenpairs of C = (A - sigma B)^-1 B with the smallest real
    part, w = sigma + 1/mu, fe_solver.py:398-400), applied matrix-free -- the reference builds the
    dense N x N matrix C.  'transform'/'pardiso' return complex128 like scipy's eigs; 'straight'
    (eigsh on an indefinite B, which the reference's author documents as unreliable) returns
    float64 with the *correct* eigenpairs.  Eigenvalues are returned sorted by descending real
    part (ARPACK's own order is not reproducible by any other implementation).

    ``real_output=True`` (extension): return float64 ``w`` / ``v`` in every mode.  The complex128 container scipy's
    ``eigs`` imposes carries a zero imaginary part (the iteration is real); at 1M triangles it doubles the result
    to 320 MB and costs ~7 ms of the device -> host copy."""
    from . import solver as _sv
    assert mesh.cells[1].data.shape[1] == 3, "must use order 1 mesh for vectorial solver"
    assert sparse_solve_mode in ["straight", "transform", "pardiso"], "sparse solve mode must be `straight` (plug into eigenproblem into eigsh) or `transform` (use spsolve to convert into an ordinary eigenproblen, then use eigs)"
    k = 2 * np.pi / wl
    if target_neff is None:
        est_eigval = np.power(k * max(IOR_dict.values()), 2)
    else:
        est_eigval = np.power(k * target_neff, 2)
    if verbose:
        print("building matrices ... ")
    prob = VectorProblem(mesh, IOR_dict, k, device, build_pattern=False)
    N = prob.N
    if N > 2000 and not ignore_warning and not sparse:
        raise Exception("A and B matrices are larger than 2000 x 2000 - this may make your system unstable. consider setting sparse=True")
    if verbose:
        print("solving...")
    with torch.cuda.device(prob.dev):
        wd, Vd, info = solve_vector_resident(prob, est_eigval, Nmax, leaf_elems)
        if sparse and sparse_solve_mode in ("transform", "pardiso") and not real_output:
            wd, Vd = _dev.as_complex(wd), _dev.as_complex(Vd)   # scipy.sparse.linalg.eigs returns complex arrays
        w = _dev.d2h(wd)
        v = _dev.d2h_pinned(Vd)
    if verbose:
        print("solving complete")
    last_info.clear()
    last_info.update(info)
    mode_count = _count_modes(w, k, IOR_dict, always=target_neff is not None)
    if plot:
        _report_modes(wl, w, k, IOR_dict, warn_always=target_neff is None)
    return w, v, mode_count


def _report_modes(wl, w, k, IOR_dict, warn_always=True):
    """The printing half of ``plot=True`` (fe_solver.py:338-342, 416-420); drawing needs
    matplotlib and stays on the host, outside this package."""
    nmin, nmax = min(IOR_dict.values()), max(IOR_dict.values())
    for _w in w:
        if _w < 0:
            continue
        ne = np.sqrt(_w / k ** 2)
        if not (nmin <= ne <= nmax) and warn_always:
            print("warning: spurious mode! stopping plotting ... ")
        print("effective index: ", get_eff_index(wl, _w))
