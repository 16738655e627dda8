"""Drop-ins for the mode-resampling functions of reference ``fe_solver.py`` / ``mesher.py``
(SURVEY.md section 8(f), rank 1): point location, interpolation weights and grid interpolation of
scalar (P2 / P1) and vector (Whitney edge) modes, on the device (``csrc/ws_interp.cu``).

Same names, argument meaning and return layout as the reference:

    construct_meshtree(mesh)                                   mesher.py:235
    find_triangle_KDtree(point, mesh, meshtree, ...)           fe_solver.py:848
    get_tri_idxs_KDtree(mesh, meshtree, xa, ya, maxr=0)        fe_solver.py:865
    find_triangle(gridpoint, mesh), get_tri_idxs(mesh, xa, ya) fe_solver.py:615, :662
    get_interp_weights(mesh, xa, ya, tri_idxs, order=2)        fe_solver.py:670
    get_interp_weights_vec(mesh, xa, ya, tri_idxs)             fe_solver.py:698
    interpolate(v, mesh, xa, ya, ...), interpolate_vec(...)    fe_solver.py:719, :746
    unstructured_interpolate(v, inmesh, point, ...)            fe_solver.py:776
    mesh_interpolate(v, inmesh, outmesh, ...)                  fe_solver.py:810

``meshtree`` is a :class:`MeshLocator` (a bin grid resident on the GPU) where the reference passes a
scipy ``KDTree`` of the triangle centroids; its search order (nearest centroid first) is
reproduced as the tie rule between the triangles that contain a point.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, device as _dev


class MeshLocator:
    """Device-resident point locator of one mesh: vertex coordinates, connectivity and the
    triangle bin grid (``ws_locator``)."""

    def __init__(self, mesh, device=None):
        self.dev = dev = _dev.require_cuda(device)
        conn = np.asarray(mesh.cells[1].data)
        if conn.size and (int(conn.max()) >= 2 ** 31 or int(conn.min()) < 0):
            raise ValueError("node ids must be non-negative and fit in int32")
        xy = np.ascontiguousarray(np.asarray(mesh.points, dtype=np.float64)[:, :2])
        self.Ne, self.stride = conn.shape
        self.Np = xy.shape[0]
        assert self.Ne > 0, "mesh has no triangles"
        with torch.cuda.device(dev):
            self.xy = _dev.h2d(xy, dev)
            self.conn = _dev.h2d(np.ascontiguousarray(conn.astype(np.int32)), dev)
            used = self.xy.index_select(0, self.conn[:, :3].reshape(-1).long())
            lo, hi = used.min(dim=0).values.tolist(), used.max(dim=0).values.tolist()
            # a degenerate box (all triangles on a line) still needs a positive extent
            ext = max(hi[0] - lo[0], hi[1] - lo[1], 1e-300)
            hi = [max(hi[0], lo[0] + 1e-12 * ext), max(hi[1], lo[1] + 1e-12 * ext)]
            h = C.c_void_p()
            _lib.check(_lib.lib().ws_locator_create(_dev.ptr(self.xy), _dev.ptr(self.conn), self.Ne, self.stride,
                                                    lo[0], lo[1], hi[0], hi[1], C.byref(h), dev.index,
                                                    _dev.stream_ptr(dev)))
        self.handle = h
        self._edge_dofs = None

    # -- device-side primitives ------------------------------------------------------------
    def _points(self, xa, ya, grid):
        xa = np.ascontiguousarray(np.asarray(xa, dtype=np.float64).ravel())
        ya = np.ascontiguousarray(np.asarray(ya, dtype=np.float64).ravel())
        if not grid:
            assert xa.shape == ya.shape
        return _dev.h2d(xa, self.dev), _dev.h2d(ya, self.dev), len(xa), (len(ya) if grid else 0)

    def locate(self, xa_d, ya_d, nx, ny, maxr=0.0, tie_rule=0):
        npts = nx * ny if ny > 0 else nx
        out = torch.empty(npts, dtype=torch.int64, device=self.dev)
        _lib.check(_lib.lib().ws_locate(self.handle, _dev.ptr(self.xy), _dev.ptr(self.conn), self.stride,
                                        _dev.ptr(xa_d), _dev.ptr(ya_d), nx, ny, float(maxr), int(tie_rule),
                                        _dev.ptr(out), self.dev.index, _dev.stream_ptr(self.dev)))
        return out

    def weights(self, xa_d, ya_d, nx, ny, tri_idx_d, kind):
        npts = nx * ny if ny > 0 else nx
        nw = {1: 3, 2: 6, 3: 6}[kind]
        w = torch.empty((npts, nw), dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_interp_weights(_dev.ptr(self.xy), _dev.ptr(self.conn), self.Ne, self.stride, kind,
                                                _dev.ptr(xa_d), _dev.ptr(ya_d), nx, ny, _dev.ptr(tri_idx_d),
                                                _dev.ptr(w), self.dev.index, _dev.stream_ptr(self.dev)))
        return w

    def apply(self, tri_idx_d, dof_d, nd, wdim, w_d, field):
        """field: numpy real or complex vector over the unknowns; returns a device tensor
        (npts[, wdim]) real or complex."""
        field = np.asarray(field)
        cplx = np.iscomplexobj(field)
        f = np.ascontiguousarray(field.astype(np.complex128 if cplx else np.float64))
        fd = _dev.h2d(f.view(np.float64) if cplx else f, self.dev)
        ncomp = 2 if cplx else 1
        npts = tri_idx_d.numel()
        out = torch.empty((npts, wdim, ncomp), dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_interp_apply(_dev.ptr(tri_idx_d), npts, self.Ne, _dev.ptr(dof_d), dof_d.shape[1], nd,
                                              wdim, _dev.ptr(w_d), _dev.ptr(fd), ncomp, _dev.ptr(out), self.dev.index,
                                              _dev.stream_ptr(self.dev)))
        out = torch.view_as_complex(out) if cplx else out[..., 0]
        return out if wdim > 1 else out[:, 0]

    def edge_dofs(self, mesh):
        if self._edge_dofs is None:
            e = np.asarray(mesh.edge_indices)
            if e.size and int(e.max()) >= 2 ** 31:
                raise ValueError("edge ids must fit in int32")
            self._edge_dofs = _dev.h2d(np.ascontiguousarray(e.astype(np.int32)), self.dev)
        return self._edge_dofs

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().ws_locator_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def construct_meshtree(mesh, device=None):
    """mesher.py:235 returns a KDTree of the triangle centroids; this returns the device locator
    that plays its role in every function below."""
    return MeshLocator(mesh, device)


def _locator(mesh, meshtree):
    if meshtree is None:
        return construct_meshtree(mesh)
    if not isinstance(meshtree, MeshLocator):
        raise TypeError("meshtree must come from wavesolve_b200.construct_meshtree(mesh) "
                        "(a scipy KDTree cannot be searched on the device)")
    return meshtree


def _tri_idxs(mesh, meshtree, xa, ya, maxr, tie_rule):
    loc = _locator(mesh, meshtree)
    with torch.cuda.device(loc.dev):
        xd, yd, nx, ny = loc._points(xa, ya, True)
        idx = loc.locate(xd, yd, nx, ny, maxr, tie_rule)
        return _dev.d2h(idx).reshape(nx, ny).astype(int)


def get_tri_idxs_KDtree(mesh, meshtree, xa, ya, maxr=0):
    """fe_solver.py:865-872: (len(xa), len(ya)) int array, -1 outside the mesh / beyond maxr."""
    return _tri_idxs(mesh, meshtree, xa, ya, maxr, 0)


def get_tri_idxs(mesh, xa, ya, meshtree=None):
    """fe_solver.py:662-668 (first containing triangle in mesh order)."""
    return _tri_idxs(mesh, meshtree, xa, ya, 0, 1)


def find_triangle_KDtree(point, mesh, meshtree, max_tries=-1, maxr=0):
    """fe_solver.py:848-863: index of the containing triangle, None if there is none, -1 beyond
    maxr.  (``max_tries`` bounds the reference's sequential search; the device search is exhaustive.)"""
    if maxr > 0 and np.sqrt(point[0] ** 2 + point[1] ** 2) > maxr:
        return -1
    idx = int(_tri_idxs(mesh, meshtree, [point[0]], [point[1]], 0, 0)[0, 0])
    return None if idx < 0 else idx


def find_triangle(gridpoint, mesh, meshtree=None):
    """fe_solver.py:615-633."""
    idx = int(_tri_idxs(mesh, meshtree, [gridpoint[0]], [gridpoint[1]], 0, 1)[0, 0])
    return None if idx < 0 else idx


def _weights(mesh, xa, ya, tri_idxs, kind, meshtree=None):
    loc = _locator(mesh, meshtree)
    tri_idxs = np.asarray(tri_idxs)
    with torch.cuda.device(loc.dev):
        xd, yd, nx, ny = loc._points(xa, ya, True)
        assert tri_idxs.shape == (nx, ny), "tri_idxs must have shape (len(xa), len(ya))"
        td = _dev.h2d(np.ascontiguousarray(tri_idxs.astype(np.int64)).ravel(), loc.dev)
        w = loc.weights(xd, yd, nx, ny, td, kind)
        return _dev.d2h(w), nx, ny


def get_interp_weights(mesh, xa, ya, tri_idxs, order=2, meshtree=None):
    """fe_solver.py:670-696: (len(xa), len(ya), 6 | 3) weights, NaN where tri_idxs == -1."""
    assert order in [1, 2], "order must be 1 or 2, corresponding to mesh element order"
    w, nx, ny = _weights(mesh, xa, ya, tri_idxs, 2 if order == 2 else 1, meshtree)
    return w.reshape(nx, ny, 6 if order == 2 else 3)


def get_interp_weights_vec(mesh, xa, ya, tri_idxs, meshtree=None):
    """fe_solver.py:698-717: (len(xa), len(ya), 3, 2) vector weights of the three edge functions."""
    w, nx, ny = _weights(mesh, xa, ya, tri_idxs, 3, meshtree)
    return w.reshape(nx, ny, 3, 2)


def _interpolate(v, mesh, xa, ya, tri_idxs, interp_weights, meshtree, maxr, kind):
    loc = _locator(mesh, meshtree)
    with torch.cuda.device(loc.dev):
        xd, yd, nx, ny = loc._points(xa, ya, True)
        if tri_idxs is None:
            td = loc.locate(xd, yd, nx, ny, maxr, 0)
        else:
            tri_idxs = np.asarray(tri_idxs)
            assert tri_idxs.shape == (nx, ny)
            td = _dev.h2d(np.ascontiguousarray(tri_idxs.astype(np.int64)).ravel(), loc.dev)
        if interp_weights is None:
            wd = loc.weights(xd, yd, nx, ny, td, kind)
        else:
            wd = _dev.h2d(np.ascontiguousarray(np.asarray(interp_weights, dtype=np.float64)).reshape(nx * ny, -1), loc.dev)
        if kind == 3:
            Ntt = len(mesh.cells[0].data)                       # fe_solver.py:760-762
            out = loc.apply(td, loc.edge_dofs(mesh), 3, 2, wd, np.asarray(v)[:Ntt])
            return _dev.d2h(out).reshape(nx, ny, 2)
        nd = 6 if kind == 2 else 3
        out = loc.apply(td, loc.conn, nd, 1, wd, v)
        return _dev.d2h(out).reshape(nx, ny)


def interpolate(v, mesh, xa, ya, tri_idxs=None, interp_weights=None, meshtree=None, order=2, maxr=0):
    """fe_solver.py:719-744: the SCALAR mode ``v`` on the grid xa x ya, NaN outside the mesh."""
    assert order in [1, 2]
    return _interpolate(v, mesh, xa, ya, tri_idxs, interp_weights, meshtree, maxr, 2 if order == 2 else 1)


def interpolate_vec(v, mesh, xa, ya, tri_idxs=None, interp_weights=None, meshtree=None, maxr=0):
    """fe_solver.py:746-774: transverse field (x, y components) of the VECTOR mode ``v`` on the
    grid; needs ``get_unique_edges(mesh)`` to have run."""
    return _interpolate(v, mesh, xa, ya, tri_idxs, interp_weights, meshtree, maxr, 3)


def _interpolate_points(v, inmesh, px, py, inmeshtree):
    loc = _locator(inmesh, inmeshtree)
    with torch.cuda.device(loc.dev):
        xd, yd, n, _ = loc._points(px, py, False)
        td = loc.locate(xd, yd, n, 0, 0.0, 0)
        wd = loc.weights(xd, yd, n, 0, td, 2)
        return _dev.d2h(loc.apply(td, loc.conn, 6, 1, wd, v))


def unstructured_interpolate(v, inmesh, point, inmeshtree=None, max_tries=10):
    """fe_solver.py:776-808: value of the P2 field at one point, NaN outside the mesh.
    (The reference gives up after ``max_tries`` nearest centroids; the device search is exhaustive.)"""
    return _interpolate_points(v, inmesh, [point[0]], [point[1]], inmeshtree)[0]


def mesh_interpolate(v, inmesh, outmesh, inmeshtree=None, max_tries=10):
    """fe_solver.py:810-828: the P2 field ``v`` of ``inmesh`` resampled on the nodes of ``outmesh``."""
    pts = np.asarray(outmesh.points, dtype=np.float64)
    return _interpolate_points(v, inmesh, pts[:, 0], pts[:, 1], inmeshtree)
