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
                                               briggs_factors, freq_chan, grid_parms, n_threads=1):
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
    _lib().oracle_imaging_weight_degrid_mt(
        _p(out), _p(g), _p(bf), _p(uvw), _p(freq), _p(chan_map), _p(pol_map), _p(nat),
        _i64(n_time), _i64(n_baseline), _i64(n_chan), _i64(n_pol), _i64(n_ic), _i64(n_pol),
        _i64(int(n_uv[0])), _i64(int(n_uv[1])), _p(delta_lm), ctypes.c_int(n_threads))
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


# --------------------------------------------------------------------------------------------------
# N3  direction_rotate: uvw rotation + visibility phasor.
#     Restates /root/reference/ngcasa/imaging/direction_rotate.py:127-248 with explicit loops
#     (no scipy Rotation, no BLAS matmul, no xarray).  PINNED by tests/golden/direction_rotate_*.npz,
#     which were produced by the reference's own calc_rotation_mats / apply_rotation_matrix / apply_phasor.
# --------------------------------------------------------------------------------------------------
INT_NAN = -2147483648          # cngi/_utils/_constants.py:19
_C0 = 299792458.0              # scipy.constants.c (direction_rotate.py:239)


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def directional_cosine(radec):
    """direction_rotate.py:177-188."""
    return np.array([np.cos(radec[0]) * np.cos(radec[1]), np.sin(radec[0]) * np.cos(radec[1]), np.sin(radec[1])])


def calc_rotation_mats(field_id_of_samples, field_ids, field_phase_dir, new_phase_center,
                       common_tangent_reprojection=True):
    """direction_rotate.py:127-175.  field_id_of_samples: the FIELD_ID array (n_time, n_baseline);
    field_ids/field_phase_dir: the FIELD table (ids, (n, 2) ra/dec radians).  Intrinsic Euler 'XZ' = Rx*Rz,
    'ZX' = Rz*Rx (scipy.spatial.transform convention for upper-case axes).
    Returns uvw_rotmat (n_field, 3, 3), phase_rotation (n_field, 3), rot_field_id (sorted unique ids > -1)."""
    ra, dec = new_phase_center[0], new_phase_center[1]
    rot_new = _rot_x(np.pi / 2 - dec) @ _rot_z(-ra + np.pi / 2)
    cos_new = directional_cosine(np.array([ra, dec]))
    ids = np.unique(np.asarray(field_id_of_samples))
    ids = ids[ids > -1]
    table = {int(f): np.asarray(field_phase_dir)[i] for i, f in enumerate(np.asarray(field_ids))}
    rotmat = np.zeros((len(ids), 3, 3))
    phase_rot = np.zeros((len(ids), 3))
    for i, f in enumerate(ids):
        pc = table[int(f)]
        rot_field = _rot_z(-np.pi / 2 + pc[0]) @ _rot_x(pc[1] - np.pi / 2)
        rotmat[i] = (rot_new @ rot_field).T
        if common_tangent_reprojection:
            rotmat[i, 2, 0:2] = 0.0
        phase_rot[i] = rot_new @ (cos_new - directional_cosine(pc))
    return rotmat, phase_rot, ids


def field_index_per_time(field_id_of_samples, rot_field_id, nan_value=-1, greater=True):
    """The per-integration field lookup of direction_rotate.py:196-201 (ids > -1) / :224-228 (ids != INT_NAN):
    the field must be constant over baseline (assert in the reference)."""
    f = np.asarray(field_id_of_samples)
    out = np.zeros(f.shape[0], dtype=np.int64)
    for t in range(f.shape[0]):
        row = f[t]
        u = np.unique(row[row > -1] if greater else row[row != INT_NAN])
        assert len(u) == 1, "direction_rotate only supports xds where field_id remains constant over baseline."
        out[t] = np.where(np.asarray(rot_field_id) == u[0])[0][0]
    return out


def apply_rotation_matrix(uvw, field_id_of_samples, uvw_rotmat, rot_field_id):
    """direction_rotate.py:190-213: uvw_rot[t, b, :] = uvw[t, b, :] @ R[field(t)], written out term by term."""
    idx = field_index_per_time(field_id_of_samples, rot_field_id, greater=True)
    R = uvw_rotmat[idx]                                   # (n_time, 3, 3)
    out = np.empty_like(uvw)
    for k in range(3):
        out[:, :, k] = (uvw[:, :, 0] * R[:, None, 0, k] + uvw[:, :, 1] * R[:, None, 1, k]) + uvw[:, :, 2] * R[:, None, 2, k]
    return out


def apply_phasor(vis_data, uvw_rot, field_id_of_samples, freq_chan, phase_rotation, rot_field_id,
                 common_tangent_reprojection, single_precision):
    """direction_rotate.py:217-248.  phase = ((2*pi*d) * f) * (1/c) -- numpy evaluates
    `2.0*1j*np.pi*d*f/c` left to right and its complex/real division multiplies by the reciprocal
    (verified bit for bit against numpy in tests/test_oracle_golden.py)."""
    idx = field_index_per_time(field_id_of_samples, rot_field_id, greater=False)
    P = phase_rotation[idx]                               # (n_time, 3)
    d = uvw_rot[:, :, 0] * P[:, None, 0] + uvw_rot[:, :, 1] * P[:, None, 1]
    if not common_tangent_reprojection:
        d = d + uvw_rot[:, :, 2] * P[:, None, 2]
    y = (((2.0 * np.pi) * d)[:, :, None] * np.asarray(freq_chan)[None, None, :]) * (1.0 / _C0)
    phasor = np.cos(y) + 1j * np.sin(y)
    out = vis_data * phasor[:, :, :, None]
    if single_precision:
        out = out.astype(np.complex64).astype(np.complex128)
    return out


# --------------------------------------------------------------------------------------------------
# N2  make_gridding_convolution_function, a_term only (the reference forces
#     a_term=True, ps_term=False: make_gridding_convolution_function.py:131-132).
#     Restates make_gridding_convolution_function.py:161-311 (pipeline), :331-359 (phase gradient),
#     :361-392 (resize + support), :394-412 (baseline patterns), :414-457 (support search), :512-560 (maps) and
#     _imaging_utils/_make_pb_symmetric.py:135-235 (Airy patterns) with numpy.fft and scipy.special.jn.
#     PINNED by tests/golden/gcf_*.npz (reference functions run by file path), except world2pix: astropy is absent
#     here, so the FITS SIN projection is written out and pinned to the CASA test vector quoted in the reference's
#     own comments (:565-577) -- "parity unpinned" beyond that vector.
# --------------------------------------------------------------------------------------------------
def airy_disk_rorder(freq_chan, n_pol, pb_parms, grid_parms, casa=True):
    """_make_pb_symmetric.py:187-235 (casa=True) / :135-183.  Returns (n_dish, n_chan, n_pol, n0, n1)."""
    from scipy.special import jn
    cell, size, centre = grid_parms["cell_size"], grid_parms["image_size"], grid_parms["image_center"]
    k = (2 * np.pi * np.asarray(freq_chan, dtype=np.float64)) / _C0
    x = np.arange(-centre[0], size[0] - centre[0]) * cell[0]
    y = np.arange(-centre[1], size[1] - centre[1]) * cell[1]
    xg, yg = np.meshgrid(x, y, indexing="ij")
    rad = np.sqrt(xg ** 2 + yg ** 2)
    out = np.zeros((len(pb_parms["list_dish_diameters"]), len(k), 1, size[0], size[1]))
    for i, (dish, block) in enumerate(zip(pb_parms["list_dish_diameters"], pb_parms["list_blockage_diameters"])):
        r = np.moveaxis(rad[:, :, None] * k * (dish / 2), 2, 0)
        r[:, centre[0], centre[1]] = 1.0
        if block == 0.0:
            v = 2.0 * jn(1, r) / r
        elif casa:
            area_ratio, length_ratio = (dish / block) ** 2, dish / block
            v = (area_ratio * 2.0 * jn(1, r) / r - 2.0 * jn(1, r * length_ratio) / (r * length_ratio)) / (area_ratio - 1.0)
        else:
            e = block / dish
            v = (2.0 * jn(1, r) / r - 2.0 * e * jn(1, r * e) / r) / (1.0 - e ** 2)
        out[i, :, 0] = v ** pb_parms["ipower"]
    out[:, :, 0, centre[0], centre[1]] = 1.0
    return np.tile(out, (1, 1, n_pol, 1, 1))


def create_cf_baseline_map(unique_ant_indx, baseline_ant, n_unique_ant):
    """:512-528.  Pairs (i <= j) of antenna types; baselines whose types come as (j, i) with j > i keep map 0
    (the reference only matches the ordered pair)."""
    pairs = np.array([[i, j] for i in range(n_unique_ant) for j in range(i, n_unique_ant)], dtype=int).reshape(-1, 2)
    types = np.asarray(unique_ant_indx)[np.asarray(baseline_ant)]
    cf_map = np.zeros(len(types), dtype=int)
    for k, (i, j) in enumerate(pairs):
        cf_map[(types[:, 0] == i) & (types[:, 1] == j)] = k
    return cf_map, pairs


def create_cf_chan_map(freq_chan, chan_tolerance_factor):
    """:536-560."""
    f = np.asarray(freq_chan, dtype=np.float64)
    n_chan = len(f)
    tol = np.max(f) * chan_tolerance_factor
    n_pb = int(np.floor((np.max(f) - np.min(f)) / tol) + 0.5)
    if n_pb == 0:
        n_pb = 1
    if n_pb >= n_chan:
        return np.arange(n_chan), f
    width = (np.max(f) - np.min(f)) / n_pb
    pb_freq = np.arange(n_pb) * width + np.min(f) + width / 2
    return np.array([np.abs(pb_freq - v).argmin() for v in f], dtype=int), pb_freq


def calc_conv_size(sub, n_pad, cut_level, oversampling, max_support):
    """:414-457.  Walks +x then +y from the centre of the (real) weight kernel until <= cut * max|.|."""
    a = np.abs(sub)
    cut = cut_level * a.max()
    assert a.min() < cut, "######### ERROR: support_cut_level too small or imsize too small."
    sup = []
    for axis in (0, 1):
        i = [n_pad[0] // 2, n_pad[1] // 2]
        while sub[i[0], i[1]] > cut:
            i[axis] += 1
            assert i[axis] < n_pad[axis], "######### ERROR: support_cut_level too small or imsize too small."
        approx = i[axis] - n_pad[axis] // 2
        sup.append((int(0.5 + approx / oversampling[axis]) + 1) * 2 + 1)
    assert sup[0] < max_support[0] and sup[1] < max_support[1], \
        "######### ERROR: support_cut_level too small or imsize too small."
    s = max(sup)
    return [s, s]


def resize_and_calc_support(conv_kernel, conv_weight_kernel, gcf_parms, grid_parms):
    """:361-392.  Inputs (n_pair, n_chan, n_pol, n0, n1) real; crop to resize_conv_size about the centre, divide
    each by the sum over its (support + 1) * oversampling window / (os_u os_v)."""
    n_pad, rs, osamp = grid_parms["image_size_padded"], gcf_parms["resize_conv_size"], gcf_parms["oversampling"]
    shape = conv_kernel.shape[:3]
    support = np.zeros(shape + (2,), dtype=int)
    start = n_pad // 2 - rs // 2
    ck = np.zeros(shape + tuple(rs))
    wk = np.zeros(shape + tuple(rs))
    for idx in np.ndindex(*shape):
        support[idx] = calc_conv_size(conv_weight_kernel[idx], n_pad, gcf_parms["support_cut_level"], osamp,
                                      gcf_parms["max_support"])
        emb = (support[idx] + 1) * osamp
        e0 = rs // 2 - emb // 2
        for src, dst in ((conv_kernel, ck), (conv_weight_kernel, wk)):
            c = src[idx][start[0]:start[0] + rs[0], start[1]:start[1] + rs[1]]
            norm = np.real(np.sum(c[e0[0]:e0[0] + emb[0], e0[1]:e0[1] + emb[1]]) / (osamp[0] * osamp[1]))
            dst[idx] = c / norm
    return ck, wk, support


def sin_world2pix_offset(field_phase_dir, phase_center, cell_size):
    """Pixel offset of each field centre from the reference pixel under the FITS RA---SIN / DEC--SIN projection
    (what `w.all_world2pix(dir_deg, 1) - n_pad//2` returns at :348 with crpix = n_pad//2, cdelt = cell in degrees):
    x = cos(dec) sin(ra - ra0), y = sin(dec) cos(dec0) - cos(dec) sin(dec0) cos(ra - ra0), offset = (x, y) / cell."""
    d = np.asarray(field_phase_dir, dtype=np.float64).reshape(-1, 2)
    ra0, dec0 = phase_center[0], phase_center[1]
    dra = d[:, 0] - ra0
    x = np.cos(d[:, 1]) * np.sin(dra)
    y = np.sin(d[:, 1]) * np.cos(dec0) - np.cos(d[:, 1]) * np.sin(dec0) * np.cos(dra)
    return np.stack([x / cell_size[0], y / cell_size[1]], axis=1)


def make_phase_gradient(field_phase_dir, gcf_parms, grid_parms):
    """:331-359 with the analytic SIN projection in place of astropy.wcs."""
    n_pad = grid_parms["image_size_padded"]
    pix_dist = sin_world2pix_offset(field_phase_dir, gcf_parms["phase_center"], grid_parms["cell_size"])
    pix = -(pix_dist) * 2 * np.pi / (n_pad * gcf_parms["oversampling"])
    size = gcf_parms["resize_conv_size"]
    c = size // 2
    xg, yg = np.meshgrid(np.arange(-c[0], size[0] - c[0]), np.arange(-c[1], size[1] - c[1]), indexing="ij")
    return np.moveaxis(np.exp(1j * (xg[:, :, None] * pix[:, 0] + yg[:, :, None] * pix[:, 1])), 2, 0)


def make_gridding_convolution_function(gcf_parms, grid_parms):
    """The a_term branch of :161-311 on plain arrays.  gcf_parms: function, list_dish_diameters,
    list_blockage_diameters, unique_ant_indx, basline_ant (n_baseline, 2), freq_chan, pol, field_phase_dir (n_field, 2),
    phase_center, oversampling, max_support, support_cut_level, chan_tolerance_factor.  Returns a dict with the
    reference's gcf_dataset variable names."""
    gp = dict(gcf_parms)
    gp["oversampling"] = np.asarray(gp.get("oversampling", [10, 10])).astype(int)
    gp["max_support"] = np.asarray(gp.get("max_support", [15, 15])).astype(int)
    gp.setdefault("support_cut_level", 2.5e-2)
    gp.setdefault("chan_tolerance_factor", 0.005)
    gp["resize_conv_size"] = (gp["max_support"] + 1) * gp["oversampling"]
    n_pad = np.asarray(grid_parms["image_size_padded"]).astype(int)
    cf_bl_map, pairs = create_cf_baseline_map(np.asarray(gp["unique_ant_indx"]), np.asarray(gp["basline_ant"]),
                                              len(gp["list_dish_diameters"]))
    cf_chan_map, pb_freq = create_cf_chan_map(gp["freq_chan"], gp["chan_tolerance_factor"])
    pb_grid = dict(cell_size=np.asarray(grid_parms["cell_size"]) * gp["oversampling"], image_size=n_pad,
                   image_center=n_pad // 2)
    casa = gp.get("function", "casa_airy") == "casa_airy"
    planes = {}
    for ipower in (1, 2):
        pat = airy_disk_rorder(pb_freq, 1, dict(gp, ipower=ipower), pb_grid, casa=casa)
        bp = np.zeros((len(pairs), len(pb_freq), 1, n_pad[0], n_pad[1]))
        for k, (i, j) in enumerate(pairs):
            bp[k, :, 0] = pat[i, :, 0] * pat[j, :, 0]
        planes[ipower] = np.real(np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(bp, axes=(3, 4)), axes=(3, 4)), axes=(3, 4)))
    ck, wk, support = resize_and_calc_support(planes[1], planes[2], gp, dict(grid_parms, image_size_padded=n_pad))
    pg = make_phase_gradient(gp["field_phase_dir"], gp, dict(grid_parms, image_size_padded=n_pad))
    return dict(SUPPORT=support, CONV_KERNEL=ck, WEIGHT_CONV_KERNEL=wk, PHASE_GRADIENT=pg,
                CF_BASELINE_MAP=cf_bl_map, CF_CHAN_MAP=cf_chan_map, CF_POL_MAP=np.zeros(len(gp["pol"]), dtype=int),
                PS_CORR_IMAGE=np.ones(tuple(grid_parms["image_size"])), pb_freq=pb_freq, pb_ant_pairs=pairs,
                oversampling=gp["oversampling"])


def airy_disk(freq_chan, n_pol, pb_parms, grid_parms, casa=True):
    """_make_pb_symmetric.py:79-132 (casa=True) / :26-76: the image-ordered variant, (l, m, chan, pol, dish).  Same
    arithmetic as airy_disk_rorder (the two reference functions differ only in the axis order of the result)."""
    return np.moveaxis(airy_disk_rorder(freq_chan, n_pol, pb_parms, grid_parms, casa=casa), (0, 1, 2, 3, 4), (4, 2, 3, 0, 1))


# ---- N4: apply_flags ------------------------------------------------------------------------------------------------
def apply_flags_variable(data, flag):
    """cngi/vis/apply_flags.py:53 for one variable: ``dv.where(flag == 0).astype(dv.dtype)``.

    xarray (absent from this image; requirements.txt xarray>=0.16) implements DataArray.where(cond) as
    numpy.where(cond, data, fill) with fill = dtypes.get_fill_value(dtype): NaN for floats, NaN + NaN j for complex
    (xarray core/dtypes.py, maybe_promote).  PARITY UNPINNED against xarray itself; synthesis_imaging_cube.py:180
    (``vis_data[flag] = np.nan``) differs only in the imaginary part of flagged complex samples (0 instead of NaN),
    which no gridder can see (both are masked by the isnan test at _standard_grid.py:340).
    """
    data = np.asarray(data)
    if not (np.issubdtype(data.dtype, np.floating) or np.issubdtype(data.dtype, np.complexfloating)):
        raise TypeError("apply_flags_variable: float / complex variables only")
    fill = (np.nan + np.nan * 1j) if np.issubdtype(data.dtype, np.complexfloating) else np.nan
    return np.where(np.asarray(flag) == 0, data, fill).astype(data.dtype)
