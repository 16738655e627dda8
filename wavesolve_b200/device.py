"""Host <-> device plumbing (PyTorch is used for device memory and streams only)."""
from __future__ import annotations

import threading
import weakref

import numpy as np
import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.WavesolveB200Error(
            "wavesolve_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    if d.type != "cuda":
        raise _lib.WavesolveB200Error("device must be a CUDA device, got %r" % (device,))
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def h2d(arr: np.ndarray, dev: torch.device) -> torch.Tensor:
    """Host -> device copy of a numpy array on the current stream.  (Allocating a fresh pinned
    staging buffer per call costs more than the pageable copy itself for these sizes.)"""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.numel() == 0:
        return torch.empty(t.shape, dtype=t.dtype, device=dev)
    return t.to(dev)


def d2h(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


_pinned_lock = threading.Lock()
_pinned_free = []          # page-locked uint8 tensors that no result array uses any more (any size)
_PINNED_KEEP = 4           # free blocks kept in total: a loop that re-meshes per z-slice asks for a new size every
#                            call, so the cache is bounded across sizes and a block serves any request it can hold


def _pinned_release(buf):
    with _pinned_lock:
        _pinned_free.append(buf)
        if len(_pinned_free) > _PINNED_KEEP:
            _pinned_free.sort(key=lambda b: b.numel())
            del _pinned_free[0]                # drop the smallest: the large blocks are the expensive ones to pin


def _pinned_take(nbytes):
    """Smallest cached block with capacity in [nbytes, 2 nbytes], else a new one (sized with 1/8 headroom so that
    slightly larger meshes of a sweep fit the same block)."""
    with _pinned_lock:
        best = None
        for i, b in enumerate(_pinned_free):
            if nbytes <= b.numel() <= 2 * nbytes and (best is None or b.numel() < _pinned_free[best].numel()):
                best = i
        if best is not None:
            return _pinned_free.pop(best)
    return torch.empty(nbytes + (nbytes >> 3), dtype=torch.uint8, pin_memory=True)


def d2h_pinned(t: torch.Tensor) -> np.ndarray:
    """Device -> host copy of a large result through page-locked memory.  A pageable ``.cpu()`` of
    the 160-320 MB eigenvector block costs several times the DMA time in page faults, and so does
    pinning a fresh block per call, so the blocks are recycled: the returned numpy array owns its
    block, and when the caller drops the array (and every view of it) the block goes back to a small,
    bounded free list that serves any later request it is large enough for."""
    t = t.detach().contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20):
        return t.cpu().numpy()
    buf = _pinned_take(nbytes)
    host = buf[:nbytes].view(t.dtype).view(t.shape)
    host.copy_(t)
    arr = host.numpy()
    weakref.finalize(arr, _pinned_release, buf)
    return arr


def d2h_staged(t: torch.Tensor, out: np.ndarray = None) -> np.ndarray:
    """Device -> host copy of a VERY large result (a gathered sweep: GBs) into ordinary pageable memory through the
    library's two reusable page-locked staging blocks (``ws_copy_to_host``): the DMA of chunk i+1 overlaps a few host
    threads copying chunk i out.  Page-locking the destination itself costs ~1 s per GB (measured: 2.0 s for the
    2.06 GB of a 256-problem sweep), more than the whole transfer at PCIe rate; a Python-level staged copy is bound by
    numpy's single-threaded memcpy into freshly mapped pages (0.6 s)."""
    t = t.detach().contiguous()
    nbytes = t.numel() * t.element_size()
    if out is None and nbytes < (128 << 20):
        return d2h_pinned(t)
    dt = np.dtype(str(t.dtype).replace("torch.", ""))
    ret = out
    if out is None:
        out = np.empty(tuple(t.shape), dtype=dt)
    else:
        assert out.flags.c_contiguous and out.nbytes == nbytes
    if nbytes:
        _lib.check(_lib.lib().ws_copy_to_host(t.data_ptr(), out.ctypes.data, nbytes, 0, t.device.index,
                                              torch.cuda.current_stream(t.device).cuda_stream))
    return ret if ret is not None else out


def as_complex(t: torch.Tensor) -> torch.Tensor:
    """float64 -> complex128 with zero imaginary part, on the device (scipy's eigs returns complex
    arrays; converting 160 MB on the host costs more than moving the extra bytes)."""
    c = torch.zeros(t.shape + (2,), dtype=t.dtype, device=t.device)
    c[..., 0] = t
    return torch.view_as_complex(c)


def ptr(t) -> int:
    if t is None:
        return 0
    return t.data_ptr() if t.numel() else 0
