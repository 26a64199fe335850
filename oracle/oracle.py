"""CPU parity oracle for the ngcasa convolutional-gridding hot path.

TEST INFRASTRUCTURE ONLY -- only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.
The product package ``cngi_prototype_b200`` never does.

The loops live in ``cngi_oracle.c`` (plain C, fp64, same operation order as the
reference's numba code) and are called through ctypes; the table / FFT / normalise
steps are restated in numpy.  Function names and argument order mirror the
reference's per-chunk operator boundary (paths relative to /root/reference):

  ngcasa/imaging/_imaging_utils/_standard_grid.py:123,180,443
  ngcasa/imaging/_imaging_utils/_aperture_grid.py:146,294,333
  ngcasa/imaging/_imaging_utils/_gridding_convolutional_kernels.py:35,101,151
  ngcasa/imaging/make_imaging_weight.py:198-213, make_image.py:116-130,
  ngcasa/imaging/_imaging_utils/_normalize.py:39-89, _remove_padding.py:20-31

Parity of this oracle is PINNED: tests/golden/*.npz were produced by the reference's
own numba functions (tests/golden/make_golden.py) and tests/test_oracle_golden.py
requires bit-exact agreement.  Exception: ``standard_degrid`` (A7) has no reference
implementation -- parity unpinned, see DESIGN.md.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64 = ctypes.c_int64
_dp = ctypes.c_void_p


def build(force=False):
    """Compile cngi_oracle.c -> liboracle.so with gcc (see Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cngi_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_standard_grid_mt.restype = ctypes.c_int
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _int64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _maps(n_chan, n_pol, chan_mode):
    # _standard_grid.py:150-159
    if chan_mode == "cube":
        n_imag_chan = n_chan
        chan_map = np.arange(n_chan, dtype=np.int64)
    else:
        n_imag_chan = 1
        chan_map = np.zeros(n_chan, dtype=np.int64)
    pol_map = np.arange(n_pol, dtype=np.int64)
    return n_imag_chan, chan_map, pol_map


# --------------------------------------------------------------------------- A8
def _prolate_spheroidal_function(u):
    """grdsf(nu) and (1-nu^2)*grdsf(nu), Schwab's rational approximation (M=6, alpha=1).

    Restates _gridding_convolutional_kernels.py:101-148: two pieces split at 0.75, numerator
    degree 4 / denominator degree 2 in (nu^2 - nu_end^2), evaluated as sum_k coef_k * x**k in
    ascending k (np.power, not Horner, so the rounding matches).
    """
    P = np.array([[8.203343e-2, -3.644705e-1, 6.278660e-1, -5.335581e-1, 2.312756e-1],
                  [4.028559e-3, -3.697768e-2, 1.021332e-1, -1.201436e-1, 6.412774e-2]])
    Q = np.array([[1.0000000e0, 8.212018e-1, 2.078043e-1],
                  [1.0000000e0, 9.599102e-1, 2.918724e-1]])
    nu = np.abs(np.asarray(u, dtype=np.float64))
    piece = np.zeros(nu.shape, dtype=np.int64)
    nu_end = np.zeros(nu.shape, dtype=np.float64)
    lo = (nu >= 0.0) & (nu < 0.75)
    hi = (nu >= 0.75) & (nu <= 1.0)
    piece[hi] = 1
    nu_end[lo] = 0.75
    nu_end[hi] = 1.0
    x = nu ** 2 - nu_end ** 2
    num = P[piece, 0]
    for k in range(1, P.shape[1]):
        num = num + P[piece, k] * np.power(x, k)
    den = Q[piece, 0]
    for k in range(1, Q.shape[1]):
        den = den + Q[piece, k] * np.power(x, k)
    grdsf = np.zeros(nu.shape, dtype=np.float64)
    good = den > 0.0
    grdsf[good] = num[good] / den[good]
    grdsf[nu > 1.0] = 0.0
    return grdsf, (1 - nu ** 2) * grdsf


def _create_prolate_spheroidal_kernel_1D(oversampling, support):
    """Half kernel tap table, length oversampling*(support//2+1) (:151-158)."""
    half = support // 2
    nu = np.arange(oversampling * half) / (half * oversampling)
    table = np.zeros(oversampling * (half + 1))
    table[: oversampling * half] = _prolate_spheroidal_function(nu)[1]
    return table


def _coordinates(npixel):
    return (np.arange(npixel) - npixel // 2) / npixel


def _create_prolate_spheroidal_image_2D(n_xy):
    """Image-plane correcting function, outer product of grdsf(|2 x|) (:35-98 tail, :182-196)."""
    gx = _prolate_spheroidal_function(np.abs(2.0 * _coordinates(int(n_xy[0]))))[0]
    gy = _prolate_spheroidal_function(np.abs(2.0 * _coordinates(int(n_xy[1]))))[0]
    return np.outer(gx, gy)


def _create_prolate_spheroidal_kernel(oversampling, support, n_uv):
    """Returns (None, kernel_image): only the correcting image is used downstream (make_image.py:109)."""
    return None, _create_prolate_spheroidal_image_2D(n_uv)


# ---------------------------------------------------------------------- A1 / A2
def _standard_grid_numpy_wrap(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, n_threads=1):
    """_standard_grid.py:123-177.  Returns (grid (n_ic,n_pol,n_u,n_v), sum_weight (n_ic,n_pol))."""
    weight = _f64(weight)
    n_time, n_baseline, n_chan, n_pol = weight.shape
    n_ic, chan_map, pol_map = _maps(n_chan, n_pol, grid_parms["chan_mode"])
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    delta_lm = _f64(grid_parms["cell_size"])
    complex_grid = bool(grid_parms["complex_grid"])
    do_psf = bool(grid_parms["do_psf"])
    grid = np.zeros((n_ic, n_pol, n_uv[0], n_uv[1]), dtype=np.complex128 if complex_grid else np.double)
    sum_weight = np.zeros((n_ic, n_pol), dtype=np.double)
    vis = None if do_psf else _c128(vis_data)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    cgk = _f64(cgk_1D)
    _lib().oracle_standard_grid_mt(
        _p(grid), _p(sum_weight), ctypes.c_int(do_psf), ctypes.c_int(0), ctypes.c_int(complex_grid),
        _p(vis), _p(uvw), _p(freq), _p(chan_map), _p(pol_map), _p(weight), _p(cgk),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_ic), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), _i64(int(grid_parms["support"])),
        _i64(int(grid_parms["oversampling"])), ctypes.c_int(grid_parms["chan_mode"] == "cube"),
        ctypes.c_int(n_threads))
    return grid, sum_weight


def _standard_grid_psf_numpy_wrap(uvw, weight, freq_chan, cgk_1D, grid_parms, n_threads=1):
    """_standard_grid.py:180-236 (real grid; also the imaging-weight density grid, A2)."""
    weight = _f64(weight)
    n_time, n_baseline, n_chan, n_pol = weight.shape
    n_ic, chan_map, pol_map = _maps(n_chan, n_pol, grid_parms["chan_mode"])
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    delta_lm = _f64(grid_parms["cell_size"])
    grid = np.zeros((n_ic, n_pol, n_uv[0], n_uv[1]), dtype=np.double)
    sum_weight = np.zeros((n_ic, n_pol), dtype=np.double)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    cgk = _f64(cgk_1D)
    _lib().oracle_standard_grid_mt(
        _p(grid), _p(sum_weight), ctypes.c_int(bool(grid_parms["do_psf"])),
        ctypes.c_int(bool(grid_parms["do_imaging_weight"])), ctypes.c_int(0),
        _p(None), _p(uvw), _p(freq), _p(chan_map), _p(pol_map), _p(weight), _p(cgk),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_ic), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), _i64(int(grid_parms["support"])),
        _i64(int(grid_parms["oversampling"])), ctypes.c_int(grid_parms["chan_mode"] == "cube"),
        ctypes.c_int(n_threads))
    return grid, sum_weight


# --------------------------------------------------------------------------- A3
def _calculate_briggs_parms(grid_of_imaging_weights, sum_weight, imaging_weights_parms):
    """make_imaging_weight.py:198-213 / synthesis_imaging_cube.py:310-325.

    grid_of_imaging_weights is kernel-side (n_chan, n_pol, n_u, n_v); returns (2, n_chan, n_pol).
    """
    if imaging_weights_parms["weighting"] == "briggs":
        robust = imaging_weights_parms["robust"]
        f = np.ones((2,) + sum_weight.shape)
        sq = np.sum(grid_of_imaging_weights ** 2, axis=(2, 3))
        f[0] = np.square(5.0 * 10.0 ** (-robust)) / (sq / sum_weight)
    else:  # uniform
        f = np.zeros((2,) + sum_weight.shape)
        f[0] = 1.0
    return f


# --------------------------------------------------------------------------- A4
def _standard_imaging_weight_degrid_numpy_wrap(grid_imaging_weight, uvw, natural_imaging_weight,
                                               briggs_factors, freq_chan, grid_parms):
    """_standard_grid.py:443-518.  grid_imaging_weight is API-side (n_u, n_v, n_chan, n_pol)."""
    nat = _f64(natural_imaging_weight)
    n_time, n_baseline, n_chan, n_pol = nat.shape
    n_ic, chan_map, pol_map = _maps(n_chan, n_pol, grid_parms["chan_mode"])
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    delta_lm = _f64(grid_parms["cell_size"])
    g = _f64(grid_imaging_weight)
    bf = _f64(briggs_factors)
    assert g.shape == (n_uv[0], n_uv[1], n_ic, n_pol), g.shape
    assert bf.shape == (2, n_ic, n_pol), bf.shape
    out = np.zeros(nat.shape, dtype=np.double)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    _lib().oracle_imaging_weight_degrid(
        _p(out), _p(g), _p(bf), _p(uvw), _p(freq), _p(chan_map), _p(pol_map), _p(nat),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_ic), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm))
    return out


# ---------------------------------------------------------------------- A5 / A6
def _aperture_common(imaging_weight, grid_parms):
    w = _f64(imaging_weight)
    n_time, n_baseline, n_chan, n_pol = w.shape
    n_ic, chan_map, pol_map = _maps(n_chan, n_pol, grid_parms["chan_mode"])
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    grid = np.zeros((n_ic, n_pol, n_uv[0], n_uv[1]), dtype=np.complex128)
    sum_weight = np.zeros((n_ic, n_pol), dtype=np.double)
    return w, n_ic, chan_map, pol_map, n_uv, grid, sum_weight


def _aperture_grid_numpy_wrap(vis_data, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map,
                              conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms):
    """_aperture_grid.py:294-331 (image) and :333-371 (psf when grid_parms['do_psf'])."""
    w, n_ic, chan_map, pol_map, n_uv, grid, sum_weight = _aperture_common(imaging_weight, grid_parms)
    n_time, n_baseline, n_chan, n_pol = w.shape
    do_psf = bool(grid_parms["do_psf"])
    vis = None if do_psf else _c128(vis_data)
    ck = _f64(conv_kernel)
    ws = _int64(weight_support)
    pg = _c128(phase_gradient)
    fld = _int64(field)
    fid = _int64(grid_parms["field_id"])
    os_ = _int64(grid_parms["oversampling"])
    cfb, cfc, cfp = _int64(cf_baseline_map), _int64(cf_chan_map), _int64(cf_pol_map)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    delta_lm = _f64(grid_parms["cell_size"])
    _lib().oracle_aperture_grid(
        _p(grid), _p(sum_weight), ctypes.c_int(do_psf), _p(vis), _p(uvw), _p(freq), _p(chan_map), _p(pol_map),
        _p(cfb), _p(cfc), _p(cfp), _p(w), _p(ck), _p(ws), _p(pg), _p(fld), _p(fid), _i64(len(fid)),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), _p(os_),
        _i64(ck.shape[0]), _i64(ck.shape[1]), _i64(ck.shape[2]), _i64(ck.shape[3]), _i64(ck.shape[4]))
    return grid, sum_weight


def _aperture_psf_grid_numpy_wrap(uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map,
                                  conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms):
    gp = dict(grid_parms)
    gp["do_psf"] = True
    return _aperture_grid_numpy_wrap(None, uvw, imaging_weight, field, cf_baseline_map, cf_chan_map,
                                     cf_pol_map, conv_kernel, weight_support, phase_gradient, freq_chan, gp)


def _aperture_weight_grid_numpy_wrap(uvw, imaging_weight, field, cf_baseline_map, cf_chan_map, cf_pol_map,
                                     weight_conv_kernel, weight_support, phase_gradient, freq_chan, grid_parms):
    """_aperture_grid.py:146-178."""
    w, n_ic, chan_map, pol_map, n_uv, grid, sum_weight = _aperture_common(imaging_weight, grid_parms)
    n_time, n_baseline, n_chan, n_pol = w.shape
    ck = _f64(weight_conv_kernel)
    ws = _int64(weight_support)
    pg = _c128(phase_gradient)
    fld = _int64(field)
    fid = _int64(grid_parms["field_id"])
    os_ = _int64(grid_parms["oversampling"])
    cfb, cfc, cfp = _int64(cf_baseline_map), _int64(cf_chan_map), _int64(cf_pol_map)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    delta_lm = _f64(grid_parms["cell_size"])
    _lib().oracle_aperture_weight_grid(
        _p(grid), _p(sum_weight), _p(uvw), _p(freq), _p(chan_map), _p(pol_map),
        _p(cfb), _p(cfc), _p(cfp), _p(w), _p(ck), _p(ws), _p(pg), _p(fld), _p(fid), _i64(len(fid)),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), _p(os_),
        _i64(ck.shape[0]), _i64(ck.shape[1]), _i64(ck.shape[2]), _i64(ck.shape[3]), _i64(ck.shape[4]))
    return grid, sum_weight


# --------------------------------------------------------------------------- A7
def _standard_degrid_numpy_wrap(model_grid, uvw, freq_chan, cgk_1D, grid_parms, n_pol=None, normalize=False):
    """Degrid predict (adjoint of A1).  NO reference implementation: parity unpinned.

    model_grid: kernel-side (n_imag_chan, n_pol, n_u, n_v) complex.  Returns vis (n_t,n_b,n_c,n_pol).
    """
    g = _c128(model_grid)
    uvw = _f64(uvw)
    freq = _f64(freq_chan)
    n_time, n_baseline = uvw.shape[:2]
    n_chan = len(freq)
    n_pol = g.shape[1] if n_pol is None else n_pol
    n_ic, chan_map, pol_map = _maps(n_chan, n_pol, grid_parms["chan_mode"])
    assert g.shape[0] == n_ic
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    delta_lm = _f64(grid_parms["cell_size"])
    cgk = _f64(cgk_1D)
    vis = np.zeros((n_time, n_baseline, n_chan, n_pol), dtype=np.complex128)
    _lib().oracle_standard_degrid(
        _p(vis), _p(g), _p(uvw), _p(freq), _p(chan_map), _p(pol_map), _p(cgk),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(g.shape[1]),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), _i64(int(grid_parms["support"])),
        _i64(int(grid_parms["oversampling"])), ctypes.c_int(int(bool(normalize))))
    return vis


# ---------------------------------------------------------------------- A9 / A10
def _remove_padding(image, image_size):
    """Centre crop of the two leading axes (_remove_padding.py:20-31)."""
    image_size = np.asarray(image_size)
    padded = np.array(image.shape[0:2])
    start = padded // 2 - image_size // 2
    end = start + image_size
    return image[start[0]:end[0], start[1]:end[1]]


def grid_to_uncorrected_image(grid, image_size):
    """fftshift(ifft2(ifftshift(G))) -> crop -> .real * (n_u*n_v)   (make_image.py:116-120).

    grid is kernel-side (n_chan, n_pol, n_u, n_v); the result is API-side (l, m, chan, pol),
    exactly as synthesis_imaging_cube.py:230-243 does with numpy.fft.
    """
    g = np.moveaxis(grid, (0, 1), (2, 3))
    n_u, n_v = g.shape[0], g.shape[1]
    img = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(g, axes=(0, 1)), axes=(0, 1)), axes=(0, 1))
    return _remove_padding(img, image_size).real * (n_u * n_v)


def correct_image(uncorrected, sum_weight, correcting_cgk):
    """(img / sum_weight[0->1]) / correcting image   (make_image.py:123-130).

    uncorrected (l,m,chan,pol); sum_weight (chan,pol); correcting_cgk (l,m) already cropped.
    """
    sw = np.array(sum_weight, dtype=np.float64, copy=True)
    sw[sw == 0] = 1
    return (uncorrected / sw[None, None, :, :]) / correcting_cgk[:, :, None, None]


def normalize_image(image, sum_weight, normalizing_image, oversampling, correct_oversampling=True):
    """_normalize.py:39-57.  image (l,m,chan,pol); normalizing_image broadcastable to image."""
    sw = np.array(sum_weight, dtype=np.float64, copy=True)
    sw[sw == 0] = 1
    if correct_oversampling:
        size = np.array(image.shape[0:2])
        center = size // 2
        sincx = np.sinc(np.arange(-center[0], size[0] - center[0]) / (size[0] * oversampling[0]))
        sincy = np.sinc(np.arange(-center[1], size[1] - center[1]) / (size[1] * oversampling[1]))
        osc = np.dot(sincx[:, None], sincy[None, :])
        return (image / sw[None, None]) / (osc[:, :, None, None] * normalizing_image)
    return (image / sw[None, None]) / normalizing_image


def max_threads():
    return int(_lib().oracle_max_threads())
