"""Small helpers shared by the operator wrappers: numpy <-> device plumbing through torch."""
import ctypes as C

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

from . import _lib
from ._lib import F32, F64, CHAN_CUBE, CHAN_CONTINUUM


def is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def chan_mode(grid_parms):
    mode = grid_parms["chan_mode"]
    if mode == "cube":
        return CHAN_CUBE
    if mode == "continuum":
        return CHAN_CONTINUUM
    raise ValueError("chan_mode must be 'cube' or 'continuum', got %r" % (mode,))


def precision_of(x):
    if is_torch(x):
        return F32 if x.dtype in (torch.float32, torch.complex64) else F64
    return F32 if np.asarray(x).dtype in (np.float32, np.complex64) else F64


def torch_dtypes(precision):
    return (torch.float32, torch.complex64) if precision == F32 else (torch.float64, torch.complex128)


def device_of(*xs):
    """Device of the first CUDA tensor among xs, else the current CUDA device (raises without a GPU)."""
    _lib.require_device()
    for x in xs:
        if is_torch(x) and x.is_cuda:
            return x.device
    return torch.device("cuda", torch.cuda.current_device())


class Uploader:
    """Moves inputs to one device with the wanted dtype and keeps them alive until the launch is queued."""

    def __init__(self, device):
        self.device = device
        self.keep = []

    def __call__(self, x, dtype):
        if x is None:
            return None
        t = x if is_torch(x) else torch.as_tensor(np.ascontiguousarray(x))
        t = t.to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def back(t, like_torch):
    """Return type follows the input type: torch in -> torch out, numpy in -> numpy out."""
    return t if like_torch else t.cpu().numpy()
