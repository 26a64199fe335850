"""uv-grid -> image: cuFFT inverse transform, crop, normalise (device resident).

Mirrors make_image.py:116-130 (ifft2 + _remove_padding + correct_image), make_psf.py:117-130 and
_imaging_utils/_normalize.py:39-89 of the reference.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from ._devutil import torch, is_torch, precision_of, torch_dtypes, device_of, Uploader, ptr, stream, back


class FFTPlan:
    """cuFFT plan + work buffer for batches of n_u x n_v planes (reused across calls)."""

    def __init__(self, n_u, n_v, max_planes, precision):
        _lib.require_device()
        self._h = C.c_void_p()
        self.key = (int(n_u), int(n_v), int(max_planes), int(precision))
        _lib.check(_lib.lib().cngi_b200_fft_plan_create(C.byref(self._h), int(n_u), int(n_v), int(max_planes),
                                                        int(precision)), "cngi_b200_fft_plan_create")

    def close(self):
        if self._h:
            _lib.lib().cngi_b200_fft_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plans = {}   # insertion-ordered: least recently used first
_plans_lock = threading.Lock()


def _plan_for(n_u, n_v, n_planes, precision, device):
    """Cached plan of a geometry ON `device` (the plan, its work buffer and phase tables are allocated on the current
    device, so creation happens inside `torch.cuda.device(device)`); least-recently-used eviction beyond 8 plans."""
    # bound the work buffer: at most ~2 GiB of complex planes per batch
    cb = 8 if precision == _lib.F32 else 16
    max_planes = max(1, min(int(n_planes), (2 << 30) // (int(n_u) * int(n_v) * cb)))
    # a plan owns its work buffer, so two host threads (dask runs the chunk functions from a thread pool) must not share one
    key = (int(n_u), int(n_v), max_planes, int(precision), str(device), threading.get_ident())
    with _plans_lock:
        plan = _plans.pop(key, None)
        if plan is None:
            while len(_plans) >= 8:
                _plans.pop(next(iter(_plans))).close()
            with torch.cuda.device(device):
                plan = FFTPlan(n_u, n_v, max_planes, precision)
        _plans[key] = plan   # (re)insert as the most recently used
    return plan


def grid_to_image(grid, image_size, sum_weight=None, corr_u=None, corr_v=None, norm_image=None, pb_image=None,
                  pb_limit=0.0, divide_by_centre=False, single_precision_roundtrip=False, centre_pixel=None):
    """Kernel-side grid (n_chan, n_pol, n_u, n_v), complex or real -> API-side image (l, m, n_chan, n_pol), real.

    image = Re(fftshift(ifft2(ifftshift(grid)))) cropped * (n_u*n_v), / sum_weight (0 -> 1),
            / (corr_u[l]*corr_v[m] * norm_image), zeroed where pb_image < pb_limit.
    norm_image / pb_image are kernel-side (n_chan, n_pol, l, m) or (l, m) (broadcast).
    divide_by_centre: every plane is divided by its pixel `centre_pixel` (default (l // 2, m // 2); make_psf_with_gcf.py:140
    uses grid_parms['image_center']); a plane whose centre value is 0 or not finite is left undivided.
    """
    L = _lib.lib()
    like_torch = is_torch(grid)
    dev = device_of(grid)
    up = Uploader(dev)
    precision = precision_of(grid)
    rdt, cdt = torch_dtypes(precision)
    is_complex = grid.is_complex() if is_torch(grid) else np.iscomplexobj(grid)
    g = up(grid, cdt if is_complex else rdt)
    n_c, n_p, n_u, n_v = (int(s) for s in g.shape)
    n_l, n_m = int(image_size[0]), int(image_size[1])
    image = torch.empty((n_c, n_p, n_l, n_m), dtype=rdt, device=dev)
    a = _lib.GridToImageArgs()
    a.n_planes, a.n_u, a.n_v = n_c * n_p, n_u, n_v
    a.image_size[0], a.image_size[1] = n_l, n_m
    a.grid, a.grid_is_complex, a.precision = ptr(g), int(is_complex), precision
    a.sum_weight = ptr(up(sum_weight, torch.float64))
    a.corr_u, a.corr_v = ptr(up(corr_u, torch.float64)), ptr(up(corr_v, torch.float64))
    for name, img in (("norm_image", norm_image), ("pb_image", pb_image)):
        t = up(img, rdt)
        setattr(a, name, ptr(t))
        if t is not None:
            assert tuple(t.shape) in ((n_l, n_m), (n_c, n_p, n_l, n_m)), tuple(t.shape)
            setattr(a, name + "_planes", 1 if t.dim() == 2 else n_c * n_p)
    a.pb_limit = float(pb_limit)
    a.divide_by_centre, a.single_precision_roundtrip = int(bool(divide_by_centre)), int(bool(single_precision_roundtrip))
    if divide_by_centre and centre_pixel is not None:
        a.divide_by_centre = 2
        a.centre_pixel[0], a.centre_pixel[1] = int(centre_pixel[0]), int(centre_pixel[1])
    a.image = ptr(image)
    plan = _plan_for(n_u, n_v, n_c * n_p, precision, dev)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_grid_to_image(plan._h, C.byref(a), stream()), "cngi_b200_grid_to_image")
    out = image.permute(2, 3, 0, 1)   # (l, m, chan, pol) view, like the reference's moveaxis
    return out if like_torch else out.cpu().numpy()


def image_to_grid(image, image_size_padded, corr_u=None, corr_v=None):
    """API-side real image (l, m, n_chan, n_pol) -> kernel-side complex grid (n_chan, n_pol, n_u, n_v):
    fftshift(fft2(ifftshift(pad(image / (corr_u x corr_v))))), the inverse of grid_to_image's transform (unnormalised
    forward DFT).  What a degridding predict feeds to _standard_degrid (predict_modelvis_image.py:37-40 lists the
    steps; the reference never implemented them)."""
    L = _lib.lib()
    like_torch = is_torch(image)
    dev = device_of(image)
    up = Uploader(dev)
    precision = precision_of(image)
    rdt, cdt = torch_dtypes(precision)
    img = up(image, rdt).permute(2, 3, 0, 1).contiguous()          # kernel-side planes
    n_c, n_p, n_l, n_m = (int(s) for s in img.shape)
    n_u, n_v = int(image_size_padded[0]), int(image_size_padded[1])
    grid = torch.empty((n_c, n_p, n_u, n_v), dtype=cdt, device=dev)
    a = _lib.ImageToGridArgs()
    a.n_planes, a.n_u, a.n_v = n_c * n_p, n_u, n_v
    a.image_size[0], a.image_size[1] = n_l, n_m
    a.image, a.grid, a.precision = ptr(img), ptr(grid), precision
    a.corr_u, a.corr_v = ptr(up(corr_u, torch.float64)), ptr(up(corr_v, torch.float64))
    plan = _plan_for(n_u, n_v, n_c * n_p, precision, dev)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_image_to_grid(plan._h, C.byref(a), stream()), "cngi_b200_image_to_grid")
    return grid if like_torch else grid.cpu().numpy()
