"""apply_flags on the device.

Mirrors cngi/vis/apply_flags.py:53 -- `apply_flags(mxds, vis, flags='FLAG')`: every data variable of partition `vis`
whose dims equal a flag variable's dims becomes `dv.where(flag == 0).astype(dv.dtype)`, i.e. NaN (complex: NaN + NaN j,
xarray's fill value for complex dtypes) where flagged; the flag variables themselves and variables of other dims are
passed through; a copy of the mxds with the new partition is returned (mxds_copier, cngi/_utils/_io.py:28).  The masking runs in cngi_b200_apply_flags
(csrc/apply_flags.cu); outputs are CUDA tensors.  Integer / bool variables with FLAG's dims (the reference casts NaN
back to int, undefined behaviour) raise instead of being guessed.
"""
import numpy as np

from . import _lib
from ._devutil import torch, is_torch, device_of

_KIND = None


def _kind(dtype):
    global _KIND
    if _KIND is None:
        _KIND = {torch.float32: _lib.ELEM_F32, torch.float64: _lib.ELEM_F64,
                 torch.complex64: _lib.ELEM_C64, torch.complex128: _lib.ELEM_C128}
    if dtype not in _KIND:
        raise TypeError("apply_flags: variables of dtype %s cannot hold NaN (float / complex only)" % (dtype,))
    return _KIND[dtype]


def _shape(x):
    return tuple(x.shape)


def apply_flags_chunk(data, flag, out=None, inplace=False, n_flagged=None):
    """out = where(flag == 0, data, NaN) for one variable (any shape; flag has the same shape, bool or uint8).

    data / flag: numpy arrays, lazy zarr arrays or CUDA tensors.  inplace=True overwrites `data` (a CUDA tensor) reading
    only the flag bytes -- the vis_data[flag] = nan of synthesis_imaging_cube.py:180.  n_flagged: optional CUDA uint64/int64
    scalar tensor that the number of flagged elements is added to.  Returns the CUDA result tensor.
    """
    dev = device_of(data, flag)
    d = data if is_torch(data) else torch.as_tensor(np.ascontiguousarray(data))
    f = flag if is_torch(flag) else torch.as_tensor(np.ascontiguousarray(flag))
    if _shape(d) != _shape(f):
        raise ValueError("apply_flags: data %s and flag %s differ in shape" % (_shape(d), _shape(f)))
    kind = _kind(d.dtype)
    if inplace and not (is_torch(data) and data.is_cuda and data.is_contiguous()):
        raise ValueError("apply_flags: inplace needs a contiguous CUDA tensor")
    d = d.to(dev).contiguous()
    f = f.to(dev)
    f = (f.view(torch.uint8) if f.dtype == torch.bool else (f != 0).view(torch.uint8) if f.dtype != torch.uint8 else f)
    f = f.contiguous()
    if inplace:
        out = d
    elif out is None:
        out = torch.empty_like(d)
    elif not (is_torch(out) and out.is_cuda and out.is_contiguous() and out.dtype == d.dtype and _shape(out) == _shape(d)):
        raise ValueError("apply_flags: out must be a contiguous CUDA tensor like data")
    if n_flagged is not None and not (is_torch(n_flagged) and n_flagged.is_cuda and n_flagged.element_size() == 8):
        raise ValueError("apply_flags: n_flagged must be an 8-byte integer CUDA tensor")
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cngi_b200_apply_flags(d.data_ptr() if d.numel() else None, out.data_ptr() if d.numel() else None,
                                                    f.data_ptr() if d.numel() else None, d.numel(), kind,
                                                    None if n_flagged is None else n_flagged.data_ptr(), st),
                   "cngi_b200_apply_flags")
    return out


def apply_flags(mxds, vis, flags="FLAG"):
    """cngi/vis/apply_flags.py:53 on a dataset of datasets (read_vis.Mxds, or any object with an `.attrs` mapping of
    partitions that are mappings of arrays carrying `.dims`; plain mappings of arrays are matched by shape instead)."""
    xds = mxds.attrs[vis]
    flags = [str(f) for f in np.atleast_1d(flags)]
    flagged = {k: v for k, v in xds.items() if k != "chunks"}
    # dims never change (where keeps them), so they are taken once from the input variables
    dims = {k: (tuple(getattr(v, "dims", ()) or ()) or ("shape",) + _shape(v)) for k, v in flagged.items()}
    data_vars = getattr(xds, "data_vars", None) or list(flagged)
    for fv in flags:
        for dv in data_vars:
            if dv == fv:
                continue                      # dont flag the flags (:34)
            if dims[dv] == dims[fv]:
                v = flagged[dv]
                if (v.dtype == torch.bool) if is_torch(v) else (np.dtype(v.dtype) == np.bool_):
                    # another flag variable of the same dims: where() -> NaN -> astype(bool) is True, i.e. a logical or
                    dev = device_of(v, flagged[fv])
                    a = v if is_torch(v) else torch.as_tensor(np.ascontiguousarray(v))
                    b = flagged[fv] if is_torch(flagged[fv]) else torch.as_tensor(np.ascontiguousarray(flagged[fv]))
                    flagged[dv] = a.to(dev) | (b.to(dev) != 0)
                else:
                    flagged[dv] = apply_flags_chunk(v, flagged[fv])
    out = mxds.copy() if hasattr(mxds, "copy") else mxds
    out.attrs = dict(out.attrs)
    out.attrs[vis] = flagged
    return out
