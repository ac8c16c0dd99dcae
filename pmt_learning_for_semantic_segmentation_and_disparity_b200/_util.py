"""Host-side argument checks shared by the operator shims (no compute happens here)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def pair(v, name="argument"):
    if isinstance(v, (tuple, list)):
        if len(v) != 2:
            raise ValueError(f"{name} must be an int or a pair, got {v!r}")
        return int(v[0]), int(v[1])
    return int(v), int(v)


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """Mirror of the backend's input contract: a CUDA float32 tensor; made contiguous like the reference
    wrappers do (`.contiguous()`), never moved or cast silently."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200 ops have no CPU path (no fallback by design); move it to a CUDA device")
    if t.dtype != torch.float32:
        raise NotImplementedError(f"{name} has dtype {t.dtype}; only torch.float32 is implemented")
    return t.contiguous()


def same_device(*tensors):
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise RuntimeError(f"tensors are on different devices: {dev} vs {t.device}")
    return dev


_EMPTY = 256  # non-null, 16-byte aligned, never dereferenced: every entry point returns before touching a zero-sized buffer


def ptr(t):
    """Device pointer of a tensor for the C ABI (NULL for None).  PyTorch gives empty tensors a NULL data pointer,
    which the library would reject as a missing argument, so they are passed as a sentinel address instead."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr() if t.numel() else _EMPTY)


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(fn_name: str, device, *args) -> None:
    """Run one C-ABI entry point on `device`'s current stream."""
    lib = _lib.load()
    with torch.cuda.device(device):
        status = getattr(lib, fn_name)(*args, stream_ptr(device))
    _lib.check(status, fn_name)
