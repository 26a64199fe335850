"""Mosaic / A-projection imaging API on mapping datasets: make_mosaic_pb, make_image_with_gcf, make_psf_with_gcf.

Mirrors /root/reference/ngcasa/imaging/make_mosaic_pb.py:100-160, make_image_with_gcf.py:135-150,
make_psf_with_gcf.py:126-140 and _imaging_utils/_normalize.py:20-89 (the 'forward' direction): the aperture gridders
(A5/A6) -> inverse FFT -> crop (A9) -> normalisation by the oversampling sinc, PS_CORR_IMAGE and PB / WEIGHT_PB (A10),
all on one device.  BASELINE config 3 ("mosaic make_pb + aperture gridding with
make_gridding_convolution_function") is the chain

    direction_rotate.direction_rotate -> make_gridding_convolution_function.make_gridding_convolution_function
    -> make_mosaic_pb -> make_image_with_gcf / make_psf_with_gcf                                   (mosaic_imaging)

Datasets are mappings (xarray is not in this image): vis_dataset {DATA, UVW, WEIGHT | IMAGING_WEIGHT, FIELD_ID, chan,
optional FLAG}; gcf_dataset = the dict make_gridding_convolution_function returns; img_dataset {PB, WEIGHT_PB}
(l, m, chan, pol).  Where the reference builds a dask graph over chunks plus a tree sum (_aperture_grid.py:25-142),
the time chunks are accumulated into one device-resident grid.
"""
import copy

import numpy as np

from ._devutil import torch, is_torch, device_of
from ._aperture_grid import _aperture
from ._fft import grid_to_image
from .imaging import _check_grid_parms, _time_chunks, _dev, _out


def _check_norm_parms(norm_parms):
    """_check_imaging_parms.py:110-120: norm_type flat_sky (default) | flat_noise | none, single_precision True,
    pb_limit 0.2."""
    norm_parms.setdefault("norm_type", "flat_sky")
    norm_parms.setdefault("single_precision", True)
    norm_parms.setdefault("pb_limit", 0.2)
    if norm_parms["norm_type"] not in ("flat_noise", "flat_sky", "none"):
        print("######### Parameter checking error: norm_type must be flat_noise, flat_sky or none")
        return False
    return isinstance(norm_parms["single_precision"], bool)


def _gcf_host(gcf_dataset, key):
    x = gcf_dataset[key]
    return x.cpu().numpy() if is_torch(x) else np.asarray(x)


def _aperture_graph(vis_dataset, gcf_dataset, grid_parms, mode, time_chunk, weight_key, apply_flags=False):
    """The role of _graph_aperture_grid (_aperture_grid.py:25-142): mode 'weight' -> A6 with WEIGHT_CONV_KERNEL,
    'psf' / 'image' -> A5 with CONV_KERNEL.  Returns kernel-side grid (n_chan, n_pol, n_u, n_v), sum_weight, parms."""
    _gp = copy.deepcopy(grid_parms)
    assert _check_grid_parms(_gp), "######### ERROR: grid_parms checking failed"
    _gp["oversampling"] = np.asarray(_gcf_host(gcf_dataset, "oversampling")).astype(np.int64)
    _gp["field_id"] = np.asarray(_gcf_host(gcf_dataset, "field_id")).astype(np.int64)
    _gp["do_psf"] = mode == "psf"
    wkey = weight_key if weight_key in vis_dataset else "WEIGHT"
    dev = device_of(vis_dataset[wkey], vis_dataset["UVW"])
    w, uvw = _dev(vis_dataset, wkey, dev), _dev(vis_dataset, "UVW", dev, torch.float64)
    freq = _dev(vis_dataset, "chan", dev, torch.float64)
    field = _dev(vis_dataset, "FIELD_ID", dev, torch.int64).reshape(w.shape[0], w.shape[1])
    vis = _dev(vis_dataset, "DATA", dev) if mode == "image" else None
    # like the reference, FLAG is not read unless asked (apply_flags is a separate step there, cngi/vis/apply_flags.py:53);
    # when fused it drops the sample from every product (image, psf, weight) alike -- see imaging.make_image
    flag = _dev(vis_dataset, "FLAG", dev, torch.uint8) if (apply_flags and "FLAG" in vis_dataset) else None
    if flag is not None and mode != "image":
        w = torch.where(flag != 0, torch.full((), float("nan"), dtype=w.dtype, device=dev), w)
        flag = None
    entry = "cngi_b200_aperture_weight_grid" if mode == "weight" else "cngi_b200_aperture_grid"
    kernel = gcf_dataset["WEIGHT_CONV_KERNEL" if mode == "weight" else "CONV_KERNEL"]
    maps = [_gcf_host(gcf_dataset, k) for k in ("CF_BASELINE_MAP", "CF_CHAN_MAP", "CF_POL_MAP")]
    support = _gcf_host(gcf_dataset, "SUPPORT")
    grid = sw = None
    for sl in _time_chunks(w.shape[0], time_chunk):
        grid, sw = _aperture(entry, None if vis is None else vis[sl], uvw[sl], w[sl], field[sl], *maps, kernel, support,
                             gcf_dataset["PHASE_GRADIENT"], freq, _gp, _gp["do_psf"],
                             flag=None if flag is None else flag[sl], grid=grid, sum_weight=sw)
    return grid, sw, _gp


def make_mosaic_pb(vis_dataset, gcf_dataset, grid_parms, time_chunk=0, weight_key="IMAGING_WEIGHT", apply_flags=False):
    """WEIGHT_PB = ifft of the gridded weight CFs / sum_weight, PB = sqrt(|WEIGHT_PB|)  (make_mosaic_pb.py:115-129).
    Returns {'PB', 'WEIGHT_PB', 'WEIGHT_PB_SUM_WEIGHT'}; images are (l, m, chan, pol)."""
    like_torch = is_torch(vis_dataset["UVW"])
    grid, sw, gp = _aperture_graph(vis_dataset, gcf_dataset, grid_parms, "weight", time_chunk, weight_key, apply_flags)
    weight_image = grid_to_image(grid, gp["image_size"], sum_weight=sw)
    return {"PB": _out(torch.sqrt(torch.abs(weight_image)), like_torch), "WEIGHT_PB": _out(weight_image, like_torch),
            "WEIGHT_PB_SUM_WEIGHT": _out(sw, like_torch)}


def _sinc_1d(n, oversampling):
    c = n // 2
    return np.sinc(np.arange(-c, n - c) / (n * oversampling))


def _normalized(grid, sw, gp, gcf_dataset, img_dataset, norm_parms, divide_by_centre):
    """_normalize.py 'forward' (:59-89) folded into the post-FFT pass: / sum_weight / (sinc_x sinc_y PS_CORR_IMAGE N),
    N = PB (flat_noise) | WEIGHT_PB (flat_sky) | 1 (none); pixels with PB < pb_limit -> 0; optional f32 round trip;
    make_psf_with_gcf additionally divides by the centre pixel (:140)."""
    _np = copy.deepcopy(norm_parms)
    assert _check_norm_parms(_np), "######### ERROR: norm_parms checking failed"
    dev = grid.device
    n_l, n_m = int(gp["image_size"][0]), int(gp["image_size"][1])
    osamp = gp["oversampling"]

    def kernel_side(x):       # (l, m, chan, pol) -> (chan, pol, l, m)
        t = x if is_torch(x) else torch.as_tensor(np.ascontiguousarray(x))
        return t.to(dev).permute(2, 3, 0, 1).contiguous()

    ps = gcf_dataset["PS_CORR_IMAGE"]
    ps = (ps if is_torch(ps) else torch.as_tensor(np.ascontiguousarray(ps))).to(dev)
    pb = kernel_side(img_dataset["PB"])
    if _np["norm_type"] == "flat_noise":
        norm = ps[None, None] * pb
    elif _np["norm_type"] == "flat_sky":
        norm = ps[None, None] * kernel_side(img_dataset["WEIGHT_PB"])
    else:
        norm = ps
    use_limit = _np["pb_limit"] > 0
    return grid_to_image(grid, gp["image_size"], sum_weight=sw, corr_u=_sinc_1d(n_l, int(osamp[0])),
                         corr_v=_sinc_1d(n_m, int(osamp[1])), norm_image=norm, pb_image=pb if use_limit else None,
                         pb_limit=float(_np["pb_limit"]) if use_limit else 0.0, divide_by_centre=divide_by_centre,
                         centre_pixel=gp.get("image_center"),
                         single_precision_roundtrip=_np["single_precision"])


def make_image_with_gcf(vis_dataset, gcf_dataset, img_dataset, grid_parms, norm_parms, time_chunk=0,
                        weight_key="IMAGING_WEIGHT", apply_flags=False):
    """IMAGE (l, m, chan, pol), SUM_WEIGHT (chan, pol): A5 image mode -> ifft -> crop -> _normalize."""
    like_torch = is_torch(vis_dataset["DATA"])
    grid, sw, gp = _aperture_graph(vis_dataset, gcf_dataset, grid_parms, "image", time_chunk, weight_key, apply_flags)
    img = _normalized(grid, sw, gp, gcf_dataset, img_dataset, norm_parms, False)
    return {"IMAGE": _out(img, like_torch), "SUM_WEIGHT": _out(sw, like_torch)}


def make_psf_with_gcf(vis_dataset, gcf_dataset, img_dataset, grid_parms, norm_parms, time_chunk=0,
                      weight_key="IMAGING_WEIGHT", apply_flags=False):
    """PSF (l, m, chan, pol) normalised to its centre pixel, PSF_SUM_WEIGHT: A5 psf mode (the Gaussian beam fit,
    make_psf_with_gcf.py:142-150, is image analysis and out of scope)."""
    like_torch = is_torch(vis_dataset["UVW"])
    grid, sw, gp = _aperture_graph(vis_dataset, gcf_dataset, grid_parms, "psf", time_chunk, weight_key, apply_flags)
    img = _normalized(grid, sw, gp, gcf_dataset, img_dataset, norm_parms, True)
    return {"PSF": _out(img, like_torch), "PSF_SUM_WEIGHT": _out(sw, like_torch)}


def mosaic_imaging(vis_dataset, field_dataset, rotation_parms, gcf_parms, grid_parms, norm_parms, time_chunk=0,
                   apply_flags=False):
    """BASELINE config 3 end to end on one device: rotate to the mosaic phase centre, build the A-term CFs, the
    mosaic PB, then the image and PSF.  Returns (img_dataset dict, gcf_dataset dict, rotated vis dataset)."""
    from .direction_rotate import direction_rotate
    from .make_gridding_convolution_function import make_gridding_convolution_function
    rot = direction_rotate(vis_dataset, field_dataset, rotation_parms)
    vis_rot = dict(rot, DATA=rot["DATA_ROT"], UVW=rot["UVW_ROT"])
    gp = copy.deepcopy(grid_parms)
    assert _check_grid_parms(gp), "######### ERROR: grid_parms checking failed"
    g = dict(gcf_parms)
    g.setdefault("freq_chan", vis_dataset["chan"].cpu().numpy() if is_torch(vis_dataset["chan"]) else vis_dataset["chan"])
    g.setdefault("field_phase_dir", np.asarray(field_dataset["PHASE_DIR"]))
    g.setdefault("field_id", np.asarray(field_dataset["field_id"]))
    g.setdefault("phase_center", np.asarray(rotation_parms["new_phase_center"], dtype=np.float64))
    gcf = make_gridding_convolution_function(g, gp)
    img = make_mosaic_pb(vis_rot, gcf, grid_parms, time_chunk, apply_flags=apply_flags)
    img.update(make_image_with_gcf(vis_rot, gcf, img, grid_parms, norm_parms, time_chunk, apply_flags=apply_flags))
    img.update(make_psf_with_gcf(vis_rot, gcf, img, grid_parms, norm_parms, time_chunk, apply_flags=apply_flags))
    return img, gcf, vis_rot
