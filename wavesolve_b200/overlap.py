"""Overlaps of modes in the mass-matrix inner product, on the device.

The reference hands out ``construct_B`` (fe_solver.py:71-101) "for inner products" and leaves
``u.T @ B @ v`` (notes eq. 38) to scipy on the host -- per pair of modes, after copying B and the modes
around.  Here B stays on the device as (pattern, values) and a whole overlap matrix ``G = U B V^T`` of
mode stacks (modes are ROWS, the layout ``solve_waveguide`` / ``solve_waveguide_vec`` return) is one
C-ABI call (``ws_overlap``: streaming SpMV for two columns at a time + a deterministic Gram reduction).
Also here: B-normalisation (``ws_bnormalize``; what eigsh applies to its output) and the guided-mode
count of fe_solver.py:330-346 as a function of its own.

Inputs may be numpy arrays (results come back as numpy) or CUDA tensors (results stay on the device).
Complex modes are handled as real and imaginary planes (B is real): ``conj=False`` is the literal
``u.T @ B @ v``, ``conj=True`` the Hermitian ``u^H B v``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, device as _dev
from . import fe_solver as _fe


class MassOperator:
    """B as (pattern, values) resident on one GPU."""

    def __init__(self, problem, vals):
        self._prob = problem            # keeps pattern + mesh arrays alive
        self.pattern = problem.pattern
        self.vals = vals
        self.N = problem.N
        self.dev = problem.dev

    # -- constructors ----------------------------------------------------------------------
    @classmethod
    def scalar(cls, mesh, order=2, device=None):
        """Mass matrix over ALL cells of the mesh, materials ignored: ``construct_B(mesh)``
        (fe_solver.py:71-99) for order 2, the ``NN`` blocks of fe_solver.py:117-125 for order 1."""
        prob = _fe.ScalarProblem(mesh, None, 0.0, device, all_cells_mass_only=True, order=order)
        with torch.cuda.device(prob.dev):
            vals = prob.assemble_mass() if order == 2 else prob.assemble()[1]
        return cls(prob, vals)

    @classmethod
    def vector(cls, mesh, IOR_dict, k, device=None):
        """B of the vector formulation (fe_solver.py:193-210; depends on the materials and k).
        The mesh must carry the edge table (``get_unique_edges``)."""
        prob = _fe.VectorProblem(mesh, IOR_dict, k, device)
        with torch.cuda.device(prob.dev):
            vals = prob.assemble()[1]
        return cls(prob, vals)

    # -- marshalling -----------------------------------------------------------------------
    def _planes(self, X, what):
        """-> (float64 CUDA tensor [rows, N], n modes, is_complex, was_1d, was_numpy)"""
        was_numpy = not isinstance(X, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(X)) if was_numpy else X
        was_1d = t.dim() == 1
        if was_1d:
            t = t.unsqueeze(0)
        if t.dim() != 2 or t.shape[1] != self.N:
            raise ValueError("%s must have shape (N,) or (n_modes, N) with N = %d, got %s"
                             % (what, self.N, tuple(t.shape)))
        cplx = t.is_complex()
        if cplx:
            t = t.to(self.dev, torch.complex128)
            t = torch.cat([t.real, t.imag], dim=0)
        else:
            t = t.to(self.dev, torch.float64)
        return t.contiguous(), (t.shape[0] // 2 if cplx else t.shape[0]), cplx, was_1d, was_numpy

    def _gram(self, Ud, Vd):
        nu, nv = Ud.shape[0], Vd.shape[0]
        G = torch.empty((nu, nv), dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().ws_overlap(self.pattern.handle, _dev.ptr(self.vals), _dev.ptr(Ud), Ud.stride(0) if nu else self.N,
                                         nu, _dev.ptr(Vd), Vd.stride(0) if nv else self.N, nv, _dev.ptr(G),
                                         self.dev.index, _dev.stream_ptr(self.dev)))
        return G

    # -- operations ------------------------------------------------------------------------
    def overlap(self, U, V=None, conj=False):
        """``G[i, j] = U[i] . B . V[j]`` (``conj=True``: ``conj(U[i]) . B . V[j]``).  ``V=None`` means
        ``V = U``.  1-D inputs give a scalar / vector like numpy's matmul would."""
        with torch.cuda.device(self.dev):
            Ud, nu, cu, u1, unp = self._planes(U, "U")
            if V is None:
                Vd, nv, cv, v1, vnp = Ud, nu, cu, u1, unp
            else:
                Vd, nv, cv, v1, vnp = self._planes(V, "V")
            G = self._gram(Ud, Vd)
            if cu or cv:
                RR = G[:nu, :nv]
                zero = torch.zeros_like(RR)
                IR = G[nu:, :nv] if cu else zero
                RI = G[:nu, nv:] if cv else zero
                II = G[nu:, nv:] if (cu and cv) else zero
                s = -1.0 if conj else 1.0                          # (Ur + s i Ui) B (Vr + i Vi)^T
                G = torch.complex(RR - s * II, RI + s * IR)
            if u1 and v1:
                G = G[0, 0]
            elif u1:
                G = G[0]
            elif v1:
                G = G[:, 0]
            if unp and vnp:
                out = _dev.d2h(G)
                return out[()] if out.ndim == 0 else out
            return G

    def norms2(self, U):
        """``U[i] . B . U[i]`` per mode (complex modes: ``conj(U[i]) . B . U[i]``, real)."""
        with torch.cuda.device(self.dev):
            Ud, nu, cplx, u1, unp = self._planes(U, "U")
            d = torch.empty(Ud.shape[0], dtype=torch.float64, device=self.dev)
            _lib.check(_lib.lib().ws_bnormalize(self.pattern.handle, _dev.ptr(self.vals), _dev.ptr(Ud), self.N,
                                                Ud.shape[0], _dev.ptr(d), 0, self.dev.index,
                                                _dev.stream_ptr(self.dev)))
            if cplx:
                d = d[:nu] + d[nu:]
            if u1:
                d = d[0]
            if unp:
                out = _dev.d2h(d)
                return out[()] if out.ndim == 0 else out
            return d

    def normalize(self, U):
        """Modes scaled to unit B-norm (``u / sqrt(u.B.u)``), what eigsh does to the vectors it returns
        (fe_solver.py:328).  Real float64 CUDA tensors with contiguous rows are scaled IN PLACE and
        returned; anything else returns a new array of the input's kind.  Raises for a mode whose
        B-norm is not positive (B of the vector formulation is indefinite)."""
        with torch.cuda.device(self.dev):
            inplace = (isinstance(U, torch.Tensor) and U.is_cuda and U.dtype == torch.float64 and
                       U.dim() in (1, 2) and U.is_contiguous() and U.device == self.dev)
            if inplace:
                Ud = U.view(1, -1) if U.dim() == 1 else U
                if Ud.shape[1] != self.N:
                    raise ValueError("modes must have %d unknowns" % self.N)
                nu, cplx, u1, unp = Ud.shape[0], False, U.dim() == 1, False
            else:
                Ud, nu, cplx, u1, unp = self._planes(U, "U")
                if not unp and not cplx and Ud.data_ptr() == U.data_ptr():
                    Ud = Ud.clone()
            d = torch.empty(Ud.shape[0], dtype=torch.float64, device=self.dev)
            L = _lib.lib()
            if not cplx:
                _lib.check(L.ws_bnormalize(self.pattern.handle, _dev.ptr(self.vals), _dev.ptr(Ud), Ud.stride(0), nu,
                                           _dev.ptr(d), 1, self.dev.index, _dev.stream_ptr(self.dev)))
                dh = _dev.d2h(d)
                if not np.all(dh > 0):
                    raise ValueError("mode(s) %s have a non-positive B-norm" % np.nonzero(~(dh > 0))[0].tolist())
                out = Ud
            else:
                _lib.check(L.ws_bnormalize(self.pattern.handle, _dev.ptr(self.vals), _dev.ptr(Ud), self.N, 2 * nu,
                                           _dev.ptr(d), 0, self.dev.index, _dev.stream_ptr(self.dev)))
                dd = d[:nu] + d[nu:]
                dh = _dev.d2h(dd)
                if not np.all(dh > 0):
                    raise ValueError("mode(s) %s have a non-positive B-norm" % np.nonzero(~(dh > 0))[0].tolist())
                nrm = torch.sqrt(dd)[:, None]
                out = torch.complex(Ud[:nu] / nrm, Ud[nu:] / nrm)
            if inplace:
                return U
            if u1:
                out = out[0]
            return _dev.d2h(out) if unp else out

    def apply(self, V):
        """``B @ V[j]`` per mode (rows in, rows out); real modes only."""
        with torch.cuda.device(self.dev):
            Vd, nv, cplx, v1, vnp = self._planes(V, "V")
            if cplx:
                raise ValueError("apply() takes real modes")
            out = torch.empty_like(Vd)
            for j in range(nv):
                self.pattern.spmv(self.vals, Vd[j], out[j])
            if v1:
                out = out[0]
            return _dev.d2h(out) if vnp else out

    def close(self):
        if self._prob is not None and self._prob.pattern is not None:
            self._prob.pattern.close()
        self._prob = None


def mass_operator(mesh, IOR_dict=None, k=None, order=2, device=None):
    """``order=2`` / ``1``: scalar mass matrix of ``construct_B`` (fe_solver.py:71-101);
    ``order='vec'``: B of ``construct_AB_vec`` (needs ``IOR_dict`` and ``k``)."""
    if order in (1, 2):
        return MassOperator.scalar(mesh, order, device)
    if order == "vec":
        if IOR_dict is None or k is None:
            raise ValueError("the vector B depends on the materials and k (fe_solver.py:183-199)")
        return MassOperator.vector(mesh, IOR_dict, k, device)
    raise NotImplementedError(order)


def overlap(u, v, B, conj=False):
    """``u.T @ B @ v`` (notes eq. 38) with ``B`` a :class:`MassOperator` (or a mesh: scalar P2 mass matrix)."""
    op, owned = _as_operator(B)
    try:
        return op.overlap(u, v, conj=conj)
    finally:
        if owned:
            op.close()


def overlap_matrix(U, V=None, B=None, conj=False):
    """Overlap matrix of two stacks of modes (rows)."""
    op, owned = _as_operator(B)
    try:
        return op.overlap(U, V, conj=conj)
    finally:
        if owned:
            op.close()


def normalize_modes(U, B):
    op, owned = _as_operator(B)
    try:
        return op.normalize(U)
    finally:
        if owned:
            op.close()


def _as_operator(B):
    if isinstance(B, MassOperator):
        return B, False
    if B is None:
        raise ValueError("B: a MassOperator or a mesh")
    return MassOperator.scalar(B), True


def count_modes(w, wl, IOR_dict):
    """Number of guided modes among eigenvalues ``w`` in DESCENDING order (what ``solve_waveguide``
    returns): the loop of fe_solver.py:330-346 -- negative eigenvalues are skipped, counting stops at
    the first effective index outside [min(IOR), max(IOR)]."""
    k = 2 * np.pi / wl
    return _fe._count_modes(np.asarray(w), k, IOR_dict)
