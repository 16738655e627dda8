"""Host <-> device plumbing (PyTorch is used for device memory and streams only)."""
from __future__ import annotations

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


def d2h_pinned(t: torch.Tensor) -> np.ndarray:
    """Device -> host copy of a large result through page-locked memory (torch's caching host
    allocator keeps the block for the next call once the caller drops the array; a pageable
    ``.cpu()`` of the 160 MB eigenvector block costs several times the DMA time in page faults).
    The returned numpy array owns the pinned block."""
    t = t.detach()
    if t.numel() < (1 << 16):
        return t.cpu().numpy()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out.numpy()


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
