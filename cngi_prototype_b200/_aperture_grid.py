"""Per-chunk aperture (A-projection / mosaic) gridding operators with the reference's names and arguments.

Mirrors /root/reference/ngcasa/imaging/_imaging_utils/_aperture_grid.py:
  _aperture_weight_grid_numpy_wrap :146, _aperture_grid_numpy_wrap :294, _aperture_psf_grid_numpy_wrap :333.
grid_parms keys read: chan_mode, image_size_padded, cell_size, oversampling (int[2]), field_id, do_psf.
The grid is always complex (_aperture_grid.py:62).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import (torch, is_torch, chan_mode, precision_of, torch_dtypes, device_of, Uploader, ptr, stream,
                       back)


def _aperture(entry, vis_data, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map, conv_kernel,
              weight_support, phase_gradient, freq_chan, grid_parms, do_psf, flag=None, grid=None, sum_weight=None):
    L = _lib.lib()
    like_torch = is_torch(imaging_weight)
    dev = device_of(imaging_weight, vis_data, uvw)
    up = Uploader(dev)
    precision = precision_of(imaging_weight)
    rdt, cdt = torch_dtypes(precision)
    w = up(imaging_weight, rdt)
    n_time, n_baseline, n_chan, n_pol = (int(s) for s in w.shape)
    n_ic = n_chan if grid_parms["chan_mode"] == "cube" else 1
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    ck = up(conv_kernel, torch.float64)
    ws = weight_support.cpu().numpy() if is_torch(weight_support) else np.asarray(weight_support)
    if grid is None:
        grid = torch.zeros((n_ic, n_pol, int(n_uv[0]), int(n_uv[1])), dtype=cdt, device=dev)
    if sum_weight is None:
        sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev)
    fid = grid_parms["field_id"]
    a = _lib.ApertureGridArgs()
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = n_time, n_baseline, n_chan, n_pol
    a.n_imag_chan, a.n_imag_pol, a.n_u, a.n_v = n_ic, n_pol, int(n_uv[0]), int(n_uv[1])
    a.vis = ptr(None if vis_data is None else up(vis_data, cdt))
    a.weight = ptr(w)
    a.flag = ptr(up(flag, torch.uint8))
    a.uvw, a.freq_chan = ptr(up(uvw, torch.float64)), ptr(up(freq_chan, torch.float64))
    a.field, a.field_id = ptr(up(field, torch.int64)), ptr(up(fid, torch.int64))
    a.cf_baseline_map, a.cf_chan_map = ptr(up(cf_baseline_map, torch.int64)), ptr(up(cf_chan_map, torch.int64))
    a.cf_pol_map = ptr(up(cf_pol_map, torch.int64))
    a.conv_kernel, a.weight_support = ptr(ck), ptr(up(ws, torch.int64))
    a.phase_gradient = ptr(up(phase_gradient, torch.complex128))
    a.grid, a.sum_weight = ptr(grid), ptr(sum_weight)
    cell = grid_parms["cell_size"]
    a.delta_lm[0], a.delta_lm[1] = float(cell[0]), float(cell[1])
    a.n_field = int(len(fid))
    a.n_cfb, a.n_cfc, a.n_cfp, a.n_cu, a.n_cv = (int(s) for s in ck.shape)
    os_ = np.asarray(grid_parms["oversampling"]).astype(np.int64)
    a.oversampling[0], a.oversampling[1] = int(os_[0]), int(os_[1])
    a.max_support = int(ws.max())          # np.max(weight_support), _aperture_grid.py:397
    a.precision, a.do_psf, a.chan_mode = precision, int(bool(do_psf)), chan_mode(grid_parms)
    with torch.cuda.device(dev):
        _lib.check(getattr(L, entry)(C.byref(a), stream()), entry)
    return back(grid, like_torch), back(sum_weight, like_torch)


def _aperture_grid_numpy_wrap(vis_data, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map,
                              conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms, **kw):
    return _aperture("cngi_b200_aperture_grid", vis_data, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map,
                     cf_pol_map, conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms,
                     grid_parms["do_psf"], **kw)


def _aperture_psf_grid_numpy_wrap(uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map, conv_kernel,
                                  weight_support, phase_gradient, freq_chan, grid_parms, **kw):
    return _aperture("cngi_b200_aperture_grid", None, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map,
                     cf_pol_map, conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms, True, **kw)


def _aperture_weight_grid_numpy_wrap(uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map,
                                     weight_conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms, **kw):
    return _aperture("cngi_b200_aperture_weight_grid", None, uvw, imaging_weight, field, cf_baseline_map,
                     cf_chan_map, cf_pol_map, weight_conv_kernel, weight_support, phase_gradient, freq_chan,
                     grid_parms, True, **kw)
