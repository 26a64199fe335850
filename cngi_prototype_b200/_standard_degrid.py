"""Degridding predict (model grid -> model visibilities), the adjoint of the standard gridder.

The reference only has a stub (ngcasa/imaging/predict_modelvis_image.py:20-40) and a "still needs to be
implemented" branch (_imaging_utils/_standard_grid.py:418-430); the operator here follows the same per-chunk
wrapper conventions as the gridders so that it can feed MODEL_DATA to self_cal (calibration/self_cal.py:107-109).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import (torch, is_torch, chan_mode, precision_of, torch_dtypes, device_of, Uploader, ptr, stream,
                       back)


def _standard_degrid_numpy_wrap(model_grid, uvw, freq_chan, cgk_1D, grid_parms, n_pol=None, normalize=False,
                                algorithm=0):
    """vis (n_time, n_baseline, n_chan, n_pol) from a kernel-side model grid (n_imag_chan, n_imag_pol, n_u, n_v).

    normalize=False is the exact adjoint of the gridder; normalize=True divides every sample by its tap sum, which
    is what a predict needs (the imaging side divides by sum_weight = sum w * tap sum, make_image.py:123-130)."""
    L = _lib.lib()
    like_torch = is_torch(model_grid)
    dev = device_of(model_grid, uvw)
    up = Uploader(dev)
    precision = precision_of(model_grid)
    _, cdt = torch_dtypes(precision)
    g = up(model_grid, cdt)
    uvw_t = up(uvw, torch.float64)
    freq_t = up(freq_chan, torch.float64)
    n_time, n_baseline = int(uvw_t.shape[0]), int(uvw_t.shape[1])
    n_chan = int(freq_t.numel())
    n_ic, n_ip, n_u, n_v = (int(s) for s in g.shape)
    n_pol = n_ip if n_pol is None else int(n_pol)
    vis = torch.empty((n_time, n_baseline, n_chan, n_pol), dtype=cdt, device=dev)
    a = _lib.StdDegridArgs()
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = n_time, n_baseline, n_chan, n_pol
    a.n_imag_chan, a.n_imag_pol, a.n_u, a.n_v = n_ic, n_ip, n_u, n_v
    a.model_grid, a.uvw, a.freq_chan = ptr(g), ptr(uvw_t), ptr(freq_t)
    a.cgk_1D, a.vis = ptr(up(cgk_1D, torch.float64)), ptr(vis)
    cell = grid_parms["cell_size"]
    a.delta_lm[0], a.delta_lm[1] = float(cell[0]), float(cell[1])
    a.support, a.oversampling = int(grid_parms["support"]), int(grid_parms["oversampling"])
    a.precision, a.chan_mode, a.normalize = precision, chan_mode(grid_parms), int(bool(normalize))
    a.algorithm = int(algorithm)   # 0 auto, 1 gather kernel, 2 register-window kernel
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_standard_degrid(C.byref(a), stream()), "cngi_b200_standard_degrid")
    return back(vis, like_torch)
