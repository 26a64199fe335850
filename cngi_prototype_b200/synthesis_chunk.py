"""Fused per-chunk imaging helpers, device resident.

Mirrors the chunk-level functions of /root/reference/ngcasa/imaging/synthesis_imaging_cube.py:
  _make_imaging_weight_chunk :288-308, _calculate_briggs_parms :310-325, _make_image :230-243,
  _make_psf :245-259, correct_image :223-228.
Inputs are torch CUDA tensors (or numpy arrays, which are uploaded); everything between the upload and the
returned tensors stays on the GPU.  Unlike the reference (which grids with data_weight after computing
imaging_weights, synthesis_imaging_cube.py:186,206,211 -- a known quirk), callers choose which weights to grid.
"""
import numpy as np

from . import _lib
from ._devutil import torch, is_torch
from ._standard_grid import standard_grid, standard_grid_image_psf
from ._imaging_weight import imaging_weight_grid, calculate_briggs_parms, _standard_imaging_weight_degrid_numpy_wrap
from ._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D, correcting_function_1D
from ._fft import grid_to_image


def _make_imaging_weight_chunk(uvw, data_weight, freq_chan, grid_parms, imaging_weights_parms, density=None,
                               sum_weight=None, reduce_fn=None):
    """natural -> data_weight; otherwise density grid (A2) -> Briggs factors (A3) -> degrid (A4).

    density / sum_weight: optional pre-zeroed device buffers to accumulate into.
    reduce_fn(density, sum_weight): optional hook called between the grid and the degrid -- the multi-GPU
    path all-reduces the partial density grids there (every rank needs the full density).
    """
    if imaging_weights_parms["weighting"] == "natural":
        return data_weight
    gp = dict(grid_parms)
    gp["image_size_padded"] = grid_parms["image_size"]   # no padding: no FFT follows (make_imaging_weight.py:153)
    gp["oversampling"], gp["support"] = 0, 1
    gp["do_psf"], gp["complex_grid"], gp["do_imaging_weight"] = True, False, True
    density, sum_weight = imaging_weight_grid(uvw, data_weight, freq_chan, gp, grid=density, sum_weight=sum_weight)
    if reduce_fn is not None:
        reduce_fn(density, sum_weight)
    briggs = calculate_briggs_parms(density, sum_weight, imaging_weights_parms)
    return _standard_imaging_weight_degrid_numpy_wrap(density, uvw, data_weight, briggs, freq_chan, gp,
                                                      kernel_side_layout=True)


def _make_image(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, flag=None, correct=False):
    """Grid (complex) -> ifft -> crop -> real * N  [-> / sum_weight / PS image when correct]."""
    grid, sum_weight = standard_grid(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, False, True, flag=flag)
    return _finish(grid, sum_weight, grid_parms, correct), sum_weight


def _make_psf(uvw, weight, freq_chan, cgk_1D, grid_parms, correct=False):
    grid, sum_weight = standard_grid(None, uvw, weight, freq_chan, cgk_1D, grid_parms, True, False)
    return _finish(grid, sum_weight, grid_parms, correct), sum_weight


def _finish(grid, sum_weight, grid_parms, correct):
    if correct:
        cu, cv = correcting_function_1D(grid_parms["image_size_padded"], grid_parms["image_size"])
        return grid_to_image(grid, grid_parms["image_size"], sum_weight=sum_weight, corr_u=cu, corr_v=cv)
    return grid_to_image(grid, grid_parms["image_size"])


def _make_pb(n_pol, freq_chan, pb_parms, grid_parms):
    """synthesis_imaging_cube.py:262-284: PB of every dish type for the chunk's channels, (l, m, chan, pol, dish)."""
    from . import make_pb as _mp
    func = _mp._casa_airy_disk if pb_parms.get("function", "casa_airy") == "casa_airy" else _mp._airy_disk
    f = freq_chan.cpu().numpy() if is_torch(freq_chan) else np.asarray(freq_chan)
    return func(f, np.zeros(n_pol), dict(pb_parms, ipower=2), grid_parms)


def synthesis_imaging_chunk(vis_data, uvw, data_weight, flag, freq_chan, grid_parms, imaging_weights_parms, fused=True,
                            pb_parms=None, grid_with="imaging_weight"):
    """Weights -> [PB] -> PSF -> image for one channel chunk (the shape of _synthesis_imaging_cube_std_chunk :171-220
    without the beam fit, which is image analysis).  Returns image, image_sum_weight, psf, psf_sum_weight
    (images API-side (l, m, chan, pol)), plus pb (l, m, chan, pol, dish) when pb_parms is given."""
    assert grid_with in ("imaging_weight", "data_weight"), grid_with
    if pb_parms is not None:
        out = synthesis_imaging_chunk(vis_data, uvw, data_weight, flag, freq_chan, grid_parms, imaging_weights_parms, fused,
                                      grid_with=grid_with)
        return out + (_make_pb(int(data_weight.shape[3]), freq_chan, pb_parms, grid_parms),)
    gp = dict(grid_parms)
    gp["oversampling"], gp["support"] = 100, 7
    cgk_1D = _create_prolate_spheroidal_kernel_1D(100, 7)
    w = _make_imaging_weight_chunk(uvw, data_weight, freq_chan, gp, imaging_weights_parms)
    if grid_with == "data_weight":
        # the reference's chunk function computes the imaging weights and then grids psf and image with data_weight
        # (synthesis_imaging_cube.py:183,206,210): this switch reproduces its output exactly; the default uses the weights
        # it computed (what make_imaging_weight + make_psf / make_image do) -- see INTEGRATION.md
        w = data_weight
    if fused:   # one pass over uvw / weights / vis for both grids (cngi_b200_standard_grid_image_psf)
        grid, img_sw, psf_grid, psf_sw = standard_grid_image_psf(vis_data, uvw, w, freq_chan, cgk_1D, gp, flag=flag)
        return _finish(grid, img_sw, gp, True), img_sw, _finish(psf_grid, psf_sw, gp, True), psf_sw
    psf, psf_sw = _make_psf(uvw, w, freq_chan, cgk_1D, gp, correct=True)
    img, img_sw = _make_image(vis_data, uvw, w, freq_chan, cgk_1D, gp, flag=flag, correct=True)
    return img, img_sw, psf, psf_sw
