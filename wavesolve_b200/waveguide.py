"""Drop-in for the hot-path function of reference ``waveguide.py``: ``get_unique_edges``."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, device as _dev
from .mesh_io import load_meshio_mesh  # noqa: F401  (waveguide.py:91 lives in this module in the reference)


def _torch_dtype(dt):
    """torch dtype with the same width as an integer numpy dtype (unsigned ids are non-negative
    and below 2^31 here, so the signed type of equal width carries the same bits)."""
    return {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[np.dtype(dt).itemsize]


def unique_edges_device(tris_d: torch.Tensor, num_points: int):
    """Device-side K1.  tris_d: (Ne,3) int32 CUDA tensor.  Returns (edges (Ntt,2) int32,
    edge_indices (Ne,3) int32) CUDA tensors."""
    dev = tris_d.device
    ne = tris_d.shape[0]
    edges = torch.empty((3 * ne, 2), dtype=torch.int32, device=dev)
    eidx = torch.empty((ne, 3), dtype=torch.int32, device=dev)
    ntt = C.c_int64(0)
    L = _lib.lib()
    _lib.check(L.ws_edges_first_appearance(_dev.ptr(tris_d), ne, int(num_points), _dev.ptr(edges),
                                           _dev.ptr(eidx), C.byref(ntt), dev.index, _dev.stream_ptr(dev)))
    return edges[: ntt.value], eidx


def get_unique_edges(mesh, mutate=True, device=None):
    """Reference ``waveguide.get_unique_edges`` (waveguide.py:10-43): unique edges numbered by
    first appearance, per-triangle edge ids; with ``mutate`` the mesh gains ``cells[0].data``,
    ``edge_indices`` (uint), ``edge_flips``, ``num_edges`` and ``num_points``.
    The O(E^2) list scan is replaced by a device sort/unique/rank (csrc/ws_pattern.cu)."""
    dev = _dev.require_cuda(device)
    from .fe_solver import _ids_to_device
    tris = np.asarray(mesh.cells[1].data)
    with torch.cuda.device(dev):
        td, _ = _ids_to_device(tris if tris.shape[1] == 3 else tris[:, :3], "vertex ids", dev)
        npts = int(td.max()) + 1 if td.numel() else 0
        edges_d, eidx_d = unique_edges_device(td, max(npts, mesh.points.shape[0]))
        # widen on the device, land in page-locked memory (the reference's arrays are 64-bit)
        wide = _torch_dtype(tris.dtype if tris.size else np.dtype(np.int64))
        edges = _dev.d2h_pinned(edges_d.to(wide)).view(tris.dtype if tris.size else np.int64)
        edge_indices = _dev.d2h_pinned(eidx_d.to(torch.int64)).view(np.uint)   # waveguide.py:16
        flips = _dev.d2h_pinned(torch.where(td <= td[:, [1, 2, 0]], 1.0, -1.0).to(torch.float64)) if mutate else None
    if mutate:
        mesh.cells[0].data = edges                                  # waveguide.py:38-42
        mesh.edge_indices = edge_indices
        mesh.edge_flips = flips                                     # waveguide.py:33-34
        mesh.num_edges = len(edges)
        mesh.num_points = mesh.points.shape[0]
    return edges, edge_indices
