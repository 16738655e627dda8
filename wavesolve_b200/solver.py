"""Python handles for the sparse direct solver (symbolic analysis, numeric factorisation and
solves on the device; a host model of the symbolic analysis for the CPU tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

import os as _os

# elements per leaf of the nested-dissection tree (profiles/r02_leaf_elems_sweep.txt: solve / factor time against this)
DEFAULT_LEAF_ELEMS = int(_os.environ.get("WS_LEAF_ELEMS", "48"))


class Symbolic:
    """Host-side nested-dissection analysis (``ws_symbolic``)."""

    def __init__(self, dofs: np.ndarray, xy: np.ndarray, N: int, first_class_start: int = 0,
                 leaf_elems: int = DEFAULT_LEAF_ELEMS):
        """dofs (Ne,6) int32 unknowns per element, xy (Np,2) node coordinates.  Scalar P2
        (first_class_start == 0): vertex ids are dofs[:, :3]; vector (first_class_start == Ntt):
        vertex ids are dofs[:, 3:] - Ntt."""
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        assert dofs.ndim == 2 and dofs.shape[1] == 6 and xy.ndim == 2 and xy.shape[1] == 2
        vector = first_class_start > 0
        h = C.c_void_p()
        _lib.check(_lib.lib().ws_symbolic_create(dofs.ctypes.data, dofs.shape[0], int(N),
                                                 xy.ctypes.data, xy.shape[0], 3 if vector else 0,
                                                 int(first_class_start) if vector else 0,
                                                 int(first_class_start), int(leaf_elems), C.byref(h)))
        self.handle = h
        self._query()

    @classmethod
    def on_device(cls, dofs_d, xy_d, N: int, first_class_start: int = 0, leaf_elems: int = DEFAULT_LEAF_ELEMS):
        """The same analysis run on the device (``ws_symbolic_create_device``): ``dofs_d`` (Ne,6)
        int32 and ``xy_d`` (Np,2) float64 CUDA tensors.  This is what the solve pipeline uses; the
        host constructor above is the model the CPU tests check it against."""
        from . import device as _dev
        assert dofs_d.is_cuda and xy_d.is_cuda and dofs_d.is_contiguous() and xy_d.is_contiguous()
        assert dofs_d.dim() == 2 and dofs_d.shape[1] == 6 and xy_d.dim() == 2 and xy_d.shape[1] == 2
        self = object.__new__(cls)
        vector = first_class_start > 0
        h = C.c_void_p()
        dev = dofs_d.device
        _lib.check(_lib.lib().ws_symbolic_create_device(_dev.ptr(dofs_d), dofs_d.shape[0], int(N), _dev.ptr(xy_d),
                                                        xy_d.shape[0], 3 if vector else 0,
                                                        int(first_class_start) if vector else 0,
                                                        int(first_class_start), int(leaf_elems), C.byref(h),
                                                        dev.index, _dev.stream_ptr(dev)))
        self.handle = h
        self._query()
        return self

    def _query(self):
        info = np.zeros(8, dtype=np.int64)
        dinfo = np.zeros(4, dtype=np.float64)
        _lib.check(_lib.lib().ws_symbolic_query(self.handle, info.ctypes.data, dinfo.ctypes.data))
        (self.N, self.depth, self.nfronts, self.nnzL, self.Utotal, self.maxfront, self.maxp,
         self.sumq) = [int(x) for x in info]
        self.flops, self.t_kd, self.t_order, self.t_bnd = [float(x) for x in dinfo]

    def export(self):
        perm = np.zeros(self.N, dtype=np.int32)
        p = np.zeros(self.nfronts, dtype=np.int32)
        q = np.zeros(self.nfronts, dtype=np.int32)
        start = np.zeros(self.nfronts, dtype=np.int64)
        bndoff = np.zeros(self.nfronts + 1, dtype=np.int64)
        bnd = np.zeros(self.sumq, dtype=np.int32)
        rel = np.zeros(self.sumq, dtype=np.int32)
        _lib.check(_lib.lib().ws_symbolic_export(self.handle, perm.ctypes.data, p.ctypes.data,
                                                 q.ctypes.data, start.ctypes.data, bndoff.ctypes.data,
                                                 bnd.ctypes.data, rel.ctypes.data))
        return dict(perm=perm, p=p, q=q, start=start, bndoff=bndoff, bnd=bnd, rel=rel)

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().ws_symbolic_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    """Device multifrontal LDL^T bound to a pattern and a symbolic analysis (``ws_solver``)."""

    def __init__(self, pattern, symbolic: Symbolic):
        import torch
        from . import device as _dev
        self._torch, self._dev = torch, _dev
        self.pattern = pattern
        self.symbolic = symbolic
        self.dev = pattern.dev
        h = C.c_void_p()
        _lib.check(_lib.lib().ws_solver_create(pattern.handle, symbolic.handle, C.byref(h),
                                               self.dev.index, _dev.stream_ptr(self.dev)))
        self.handle = h
        self.N = pattern.N

    def factor(self, A_vals, B_vals=None, sigma=0.0):
        d = self._dev
        _lib.check(_lib.lib().ws_solver_factor(self.handle, d.ptr(A_vals), d.ptr(B_vals), float(sigma),
                                               self.dev.index, d.stream_ptr(self.dev)))
        return self.stats()

    def solve(self, b, out=None):
        torch, d = self._torch, self._dev
        nrhs = 1 if b.dim() == 1 else b.shape[0]
        if out is None:
            out = torch.empty_like(b)
        _lib.check(_lib.lib().ws_solver_solve(self.handle, d.ptr(b), d.ptr(out), nrhs,
                                              self.dev.index, d.stream_ptr(self.dev)))
        return out

    def check(self):
        """Synchronise and raise if a solve sweep reported an internal synchronisation error."""
        _lib.check(_lib.lib().ws_solver_check(self.handle, self.dev.index, self._dev.stream_ptr(self.dev)))

    def stats(self):
        info = np.zeros(8, dtype=np.int64)
        dinfo = np.zeros(3, dtype=np.float64)
        _lib.check(_lib.lib().ws_solver_stats(self.handle, info.ctypes.data, dinfo.ctypes.data))
        keys = ("nnzL", "fronts", "maxfront", "bytes", "neg_pivots", "pos_pivots", "bad_pivots", "depth")
        out = {k: int(v) for k, v in zip(keys, info)}
        out.update(flops=float(dinfo[0]), min_pivot=float(dinfo[1]), max_pivot=float(dinfo[2]))
        q = np.zeros(4, dtype=np.int64)
        dq = np.zeros(4, dtype=np.float64)
        _lib.check(_lib.lib().ws_solver_quality(self.handle, q.ctypes.data, dq.ctypes.data))
        out.update(perturbed_pivots=int(q[0]), refine_steps=int(q[1]), probe_refinements=int(q[2]),
                   backward_error_unrefined=float(dq[0]), backward_error=float(dq[1]), max_abs=float(dq[2]),
                   pivot_threshold=float(dq[3]))
        return out

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().ws_solver_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
