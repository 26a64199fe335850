"""Imaging-weight operators (density grid, Briggs factors, weight degrid).

Mirrors /root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:180 (psf wrapper with
do_imaging_weight), :443 (_standard_imaging_weight_degrid_numpy_wrap) and
ngcasa/imaging/make_imaging_weight.py:198-213 (calculate_briggs_parms).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import (torch, is_torch, chan_mode, precision_of, torch_dtypes, device_of, Uploader, ptr, stream,
                       back)


def imaging_weight_grid(uvw, weight, freq_chan, grid_parms, grid=None, sum_weight=None, first_pol_only=False):
    """Density grid rho (n_imag_chan, n_pol, n_u, n_v) float64 and sum_weight (n_imag_chan, n_pol).

    first_pol_only (n_pol >= 2): all pol planes are identical by construction (pol-averaged weights,
    _standard_grid.py:328-330); update plane 0 only and let the caller replicate it (`replicate_pol_planes`)."""
    L = _lib.lib()
    like_torch = is_torch(weight)
    dev = device_of(weight, uvw)
    up = Uploader(dev)
    precision = precision_of(weight)
    rdt, _ = torch_dtypes(precision)
    w = up(weight, rdt)
    n_time, n_baseline, n_chan, n_pol = (int(s) for s in w.shape)
    n_ic = n_chan if grid_parms["chan_mode"] == "cube" else 1
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    if grid is None:
        grid = torch.zeros((n_ic, n_pol, int(n_uv[0]), int(n_uv[1])), dtype=torch.float64, device=dev)
    if sum_weight is None:
        sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev)
    a = _lib.IwGridArgs()
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = n_time, n_baseline, n_chan, n_pol
    a.n_imag_chan, a.n_imag_pol, a.n_u, a.n_v = n_ic, n_pol, int(n_uv[0]), int(n_uv[1])
    a.weight, a.uvw, a.freq_chan = ptr(w), ptr(up(uvw, torch.float64)), ptr(up(freq_chan, torch.float64))
    a.density, a.sum_weight = ptr(grid), ptr(sum_weight)
    cell = grid_parms["cell_size"]
    a.delta_lm[0], a.delta_lm[1] = float(cell[0]), float(cell[1])
    a.precision, a.chan_mode, a.first_pol_only = precision, chan_mode(grid_parms), int(bool(first_pol_only))
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_imaging_weight_grid(C.byref(a), stream()), "cngi_b200_imaging_weight_grid")
    return back(grid, like_torch), back(sum_weight, like_torch)


def replicate_pol_planes(density, sum_weight):
    """Copies pol plane 0 (and its sum_weight) into the other pol planes -- the second half of first_pol_only."""
    if density.shape[1] > 1:
        density[:, 1:] = density[:, :1]
        sum_weight[:, 1:] = sum_weight[:, :1]
    return density, sum_weight


def calculate_briggs_parms(grid_of_imaging_weights, sum_weight, imaging_weights_parms):
    """(2, n_chan, n_pol) Briggs factors from the kernel-side density grid (n_chan, n_pol, n_u, n_v).

    weighting 'briggs': f0 = (5*10^-robust)^2 / (sum(rho^2)/sum_weight), f1 = 1; anything else
    ('uniform'): f0 = 1, f1 = 0.  ('briggs_abs' raises NameError in the reference -- dead code.)
    """
    L = _lib.lib()
    like_torch = is_torch(grid_of_imaging_weights)
    dev = device_of(grid_of_imaging_weights, sum_weight)
    up = Uploader(dev)
    rho = up(grid_of_imaging_weights, torch.float64)
    sw = up(sum_weight, torch.float64)
    n_planes = int(sw.numel())
    n_cells = int(rho.numel() // n_planes)
    bf = torch.empty((2,) + tuple(sw.shape), dtype=torch.float64, device=dev)
    weighting = imaging_weights_parms["weighting"]
    if weighting == "briggs_abs":
        raise NotImplementedError("briggs_abs is dead code in the reference (NameError at make_imaging_weight.py:207)")
    robust = float(imaging_weights_parms.get("robust", 0.5))
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_briggs_factors(ptr(rho), ptr(sw), ptr(bf), n_planes, n_cells, robust,
                                              0 if weighting == "briggs" else 1, stream()),
                   "cngi_b200_briggs_factors")
    return back(bf, like_torch)


def _standard_imaging_weight_degrid_numpy_wrap(grid_imaging_weight, uvw, natural_imaging_weight, briggs_factors,
                                               freq_chan, grid_parms, kernel_side_layout=False):
    """imaging_weight (n_time, n_baseline, n_chan, n_pol).

    grid_imaging_weight is API-side (n_u, n_v, n_chan, n_pol) as in the reference (:514); pass
    kernel_side_layout=True to hand in the (n_chan, n_pol, n_u, n_v) array the gridder produced -- no
    transpose is made either way (the kernel takes strides).
    """
    L = _lib.lib()
    like_torch = is_torch(natural_imaging_weight)
    dev = device_of(natural_imaging_weight, grid_imaging_weight)
    up = Uploader(dev)
    precision = precision_of(natural_imaging_weight)
    rdt, _ = torch_dtypes(precision)
    nat = up(natural_imaging_weight, rdt)
    n_time, n_baseline, n_chan, n_pol = (int(s) for s in nat.shape)
    n_ic = n_chan if grid_parms["chan_mode"] == "cube" else 1
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    g = grid_imaging_weight
    if not is_torch(g):
        g = torch.as_tensor(np.asarray(g))
    g = g.to(device=dev, dtype=torch.float64)   # keeps strides; may be a moveaxis view
    if kernel_side_layout:
        assert tuple(g.shape) == (n_ic, n_pol, int(n_uv[0]), int(n_uv[1])), tuple(g.shape)
        st = g.stride()
        strides = (st[2], st[3], st[0], st[1])
    else:
        assert tuple(g.shape) == (int(n_uv[0]), int(n_uv[1]), n_ic, n_pol), tuple(g.shape)
        strides = tuple(g.stride())
    out = torch.empty(nat.shape, dtype=rdt, device=dev)
    a = _lib.IwDegridArgs()
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = n_time, n_baseline, n_chan, n_pol
    a.n_imag_chan, a.n_imag_pol, a.n_u, a.n_v = n_ic, n_pol, int(n_uv[0]), int(n_uv[1])
    a.natural_weight, a.uvw, a.freq_chan = ptr(nat), ptr(up(uvw, torch.float64)), ptr(up(freq_chan, torch.float64))
    a.density = ptr(g)
    for i in range(4):
        a.density_stride[i] = int(strides[i])
    # pol planes / factor columns that are stride-0 views of one plane are identical by construction (first_pol_only)
    a.pol_shared = int(n_pol == 2 and strides[3] == 0 and is_torch(briggs_factors) and briggs_factors.dim() == 3
                       and briggs_factors.stride(2) == 0)
    a.briggs_factors = ptr(up(briggs_factors, torch.float64))
    a.imaging_weight = ptr(out)
    cell = grid_parms["cell_size"]
    a.delta_lm[0], a.delta_lm[1] = float(cell[0]), float(cell[1])
    a.precision, a.chan_mode = precision, chan_mode(grid_parms)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_imaging_weight_degrid(C.byref(a), stream()), "cngi_b200_imaging_weight_degrid")
    return back(out, like_torch)
