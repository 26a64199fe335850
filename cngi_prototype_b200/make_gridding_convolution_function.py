"""A-term gridding convolution functions, computed on the GPU.

Mirrors the a_term branch of /root/reference/ngcasa/imaging/make_gridding_convolution_function.py (the only branch the
reference enables, :131-132) on plain parameters -- the reference pulls ANTENNA1/2, chan, pol and FIELD.PHASE_DIR out of
an xarray mxds (:107-116); here they are entries of gcf_parms:

  create_cf_baseline_map :512-528, create_cf_chan_map :536-560      host integers (kept on the host)
  make_baseline_patterns :394 + fft :246-247 + resize_and_calc_support :361     -> cngi_b200_make_gcf
  make_phase_gradient :331-359                                                  -> cngi_b200_phase_gradient
      (astropy.wcs's RA---SIN / DEC--SIN world2pix is two numbers per field: written out analytically here)

The returned mapping uses the reference's gcf_dataset variable names, so it plugs into
_aperture_grid._aperture_grid_numpy_wrap exactly like the reference's dataset does (_aperture_grid.py:59,71-77).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import torch, device_of, ptr, stream


def create_cf_baseline_map(unique_ant_indx, basline_ant, n_unique_ant):
    """Antenna-type pairs (i <= j) and, per baseline, the index of its pair.  As in the reference only the ordered
    pair (type[ant1], type[ant2]) is matched; a baseline whose types come as (j, i), j > i, keeps index 0."""
    pairs = np.array([(i, j) for i in range(n_unique_ant) for j in range(i, n_unique_ant)], dtype=int).reshape(-1, 2)
    types = np.asarray(unique_ant_indx)[np.asarray(basline_ant)]
    cf_baseline_map = np.zeros(types.shape[0], dtype=int)
    for k in range(len(pairs)):
        cf_baseline_map[(types[:, 0] == pairs[k, 0]) & (types[:, 1] == pairs[k, 1])] = k
    return cf_baseline_map, pairs


def create_cf_chan_map(freq_chan, chan_tolerance_factor):
    """Channels -> PB frequencies: one PB per `chan_tolerance_factor` of fractional bandwidth."""
    f = np.asarray(freq_chan, dtype=np.float64)
    span = np.max(f) - np.min(f)
    n_pb_chan = int(np.floor(span / (np.max(f) * chan_tolerance_factor)) + 0.5) or 1
    if n_pb_chan >= len(f):
        return np.arange(len(f)), f
    step = span / n_pb_chan
    pb_freq = np.arange(n_pb_chan) * step + np.min(f) + step / 2
    cf_chan_map = np.abs(f[:, None] - pb_freq[None, :]).argmin(axis=1).astype(int)
    return cf_chan_map, pb_freq


def _sin_offset_in_pixels(field_phase_dir, phase_center, cell_size):
    """(all_world2pix(dir, 1) - crpix) for ctype RA---SIN / DEC--SIN, crval = phase_center, cdelt = cell_size."""
    d = np.asarray(field_phase_dir, dtype=np.float64).reshape(-1, 2)
    dra = d[:, 0] - phase_center[0]
    x = np.cos(d[:, 1]) * np.sin(dra)
    y = np.sin(d[:, 1]) * np.cos(phase_center[1]) - np.cos(d[:, 1]) * np.sin(phase_center[1]) * np.cos(dra)
    return np.stack([x / cell_size[0], y / cell_size[1]], axis=1)


def make_phase_gradient(field_phase_dir, gcf_parms, grid_parms, device=None):
    """(n_field, cu, cv) complex128 CUDA tensor: exp(i (x pix_x + y pix_y)) about the CF centre."""
    L = _lib.lib()
    dev = device if device is not None else device_of()
    n_pad = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    osamp = np.asarray(gcf_parms["oversampling"]).astype(np.int64)
    size = np.asarray(gcf_parms["resize_conv_size"]).astype(np.int64)
    pix_dist = _sin_offset_in_pixels(field_phase_dir, gcf_parms["phase_center"], grid_parms["cell_size"])
    pix = -(pix_dist) * 2 * np.pi / (n_pad * osamp)
    pix_t = torch.as_tensor(np.ascontiguousarray(pix), device=dev)
    out = torch.empty((pix.shape[0], int(size[0]), int(size[1])), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_phase_gradient(ptr(pix_t), pix.shape[0], int(size[0]), int(size[1]), ptr(out), stream()),
                   "cngi_b200_phase_gradient")
    return out


_STATUS_TEXT = "######### ERROR: support_cut_level too small or imsize too small."


def make_gridding_convolution_function(gcf_parms, grid_parms, device=None):
    """gcf_parms: function ('casa_airy' | 'airy'), list_dish_diameters, list_blockage_diameters, unique_ant_indx,
    basline_ant (n_baseline, 2), freq_chan, pol, field_phase_dir (n_field, 2), phase_center, field_id (optional),
    oversampling [10, 10], max_support [15, 15], support_cut_level 0.025, chan_tolerance_factor 0.005.
    grid_parms: image_size, image_size_padded, cell_size (radians, x negative).
    Returns a dict of CUDA tensors / host integer arrays with the reference's gcf_dataset names."""
    L = _lib.lib()
    dev = device if device is not None else device_of()
    g = dict(gcf_parms)
    function = g.get("function", "casa_airy")
    assert function in ("casa_airy", "airy"), "######### ERROR: Only airy and casa_airy function has been implemented"
    osamp = np.asarray(g.get("oversampling", [10, 10])).astype(np.int64)
    max_support = np.asarray(g.get("max_support", [15, 15])).astype(np.int64)
    dish = np.ascontiguousarray(g["list_dish_diameters"], dtype=np.float64)
    block = np.ascontiguousarray(g["list_blockage_diameters"], dtype=np.float64)
    assert len(dish) == len(block), "######### ERROR: gcf_parms checking failed"
    g["oversampling"], g["max_support"] = osamp, max_support
    g["resize_conv_size"] = (max_support + 1) * osamp
    n_pad = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    cell = np.asarray(grid_parms["cell_size"], dtype=np.float64)

    cf_baseline_map, pairs = create_cf_baseline_map(np.asarray(g["unique_ant_indx"]), np.asarray(g["basline_ant"]),
                                                    len(dish))
    cf_chan_map, pb_freq = create_cf_chan_map(g["freq_chan"], g.get("chan_tolerance_factor", 0.005))
    pairs64 = np.ascontiguousarray(pairs, dtype=np.int64)
    pb_freq64 = np.ascontiguousarray(pb_freq, dtype=np.float64)
    n_pair, n_freq = len(pairs64), len(pb_freq64)
    cu, cv = int(g["resize_conv_size"][0]), int(g["resize_conv_size"][1])

    conv_kernel = torch.empty((n_pair, n_freq, 1, cu, cv), dtype=torch.float64, device=dev)
    weight_conv_kernel = torch.empty_like(conv_kernel)
    support = torch.zeros((n_pair, n_freq, 1, 2), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    a = _lib.GcfArgs()
    a.n_pad[0], a.n_pad[1] = int(n_pad[0]), int(n_pad[1])
    a.conv_size[0], a.conv_size[1] = cu, cv
    pb_cell = cell * osamp                       # make_baseline_patterns :398
    a.pb_cell[0], a.pb_cell[1] = float(pb_cell[0]), float(pb_cell[1])
    a.oversampling[0], a.oversampling[1] = int(osamp[0]), int(osamp[1])
    a.max_support[0], a.max_support[1] = int(max_support[0]), int(max_support[1])
    a.function = 1 if function == "casa_airy" else 0
    a.n_dish = len(dish)
    a.dish_diameter_host, a.blockage_diameter_host = dish.ctypes.data, block.ctypes.data
    a.n_pair, a.ant_pairs_host = n_pair, pairs64.ctypes.data
    a.n_freq, a.pb_freq_host = n_freq, pb_freq64.ctypes.data
    a.support_cut_level = float(g.get("support_cut_level", 2.5e-2))
    a.conv_kernel, a.weight_conv_kernel = ptr(conv_kernel), ptr(weight_conv_kernel)
    a.support, a.status = ptr(support), ptr(status)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_make_gcf(C.byref(a), stream()), "cngi_b200_make_gcf")
    assert int(status.item()) == 0, _STATUS_TEXT + " (status %d)" % int(status.item())

    phase_gradient = make_phase_gradient(g["field_phase_dir"], g, dict(grid_parms, image_size_padded=n_pad), device=dev)
    n_field = phase_gradient.shape[0]
    return dict(
        SUPPORT=support, CONV_KERNEL=conv_kernel, WEIGHT_CONV_KERNEL=weight_conv_kernel, PHASE_GRADIENT=phase_gradient,
        CF_BASELINE_MAP=cf_baseline_map, CF_CHAN_MAP=cf_chan_map, CF_POL_MAP=np.zeros(len(g["pol"]), dtype=int),
        PS_CORR_IMAGE=torch.ones(tuple(int(s) for s in grid_parms["image_size"]), dtype=torch.float64, device=dev),
        field_id=np.asarray(g.get("field_id", np.arange(n_field))), pb_freq=pb_freq, pb_ant_pairs=pairs,
        oversampling=osamp, cell_uv=1 / (n_pad * cell * osamp))
