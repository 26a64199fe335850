"""Small helpers shared by the operator wrappers: numpy <-> device plumbing through torch."""
import ctypes as C
import threading

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


_CONST_MAX_BYTES = 256 * 1024
_const_cache = {}            # (device, dtype, shape, content hash) -> device tensor; insertion-ordered, oldest evicted
_const_lock = threading.Lock()


def small_constant(x, dtype, device):
    """Device copy of a SMALL host array (channel frequencies, tap tables, correcting functions, index maps), cached by
    content.  Uploading a pageable numpy array is a synchronous cudaMemcpy that first drains the stream, so doing it on
    every call would serialise a caller that queues several datasets back to back; the cached tensor is read-only."""
    a = np.ascontiguousarray(x)
    key = (str(device), str(dtype), a.shape, a.dtype.str, hash(a.tobytes()))
    with _const_lock:
        t = _const_cache.get(key)
        if t is None:
            t = torch.as_tensor(a).to(device=device, dtype=dtype).contiguous()
            while len(_const_cache) >= 256:
                _const_cache.pop(next(iter(_const_cache)))
            _const_cache[key] = t
    return t


def to_device(x, dtype, device):
    """numpy / torch -> contiguous tensor of `dtype` on `device` (small host arrays through the constant cache)."""
    if is_torch(x):
        return x.to(device=device, dtype=dtype).contiguous()
    a = np.asarray(x)
    if a.nbytes <= _CONST_MAX_BYTES:
        return small_constant(a, dtype, device)
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dtype).contiguous()


class Uploader:
    """Moves inputs to one device with the wanted dtype and keeps them alive until the launch is queued."""

    def __init__(self, device):
        self.device = device
        self.keep = []

    def __call__(self, x, dtype):
        if x is None:
            return None
        t = to_device(x, dtype, self.device)
        self.keep.append(t)
        return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def back(t, like_torch):
    """Return type follows the input type: torch in -> torch out, numpy in -> numpy out."""
    return t if like_torch else t.cpu().numpy()
