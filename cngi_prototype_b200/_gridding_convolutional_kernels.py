"""Prolate-spheroidal gridding tables (host side, numpy): the tap table the gridder consumes and the
image-plane correcting function.

Mirrors /root/reference/ngcasa/imaging/_imaging_utils/_gridding_convolutional_kernels.py
(_prolate_spheroidal_function :101, _create_prolate_spheroidal_kernel_1D :151,
_create_prolate_spheroidal_kernel :35, _create_prolate_spheroidal_image_2D :182).  Negligible cost, computed
once per image and uploaded; the powers are evaluated term by term so the values equal the reference's
to the last bit (the tap table feeds bit-exact index tests).
"""
import numpy as np

# Schwab's rational approximation to the 0-order spheroidal function, m = 6, alpha = 1, in two pieces.
_NUM = np.array([[8.203343e-2, -3.644705e-1, 6.278660e-1, -5.335581e-1, 2.312756e-1],
                 [4.028559e-3, -3.697768e-2, 1.021332e-1, -1.201436e-1, 6.412774e-2]])
_DEN = np.array([[1.0000000e0, 8.212018e-1, 2.078043e-1],
                 [1.0000000e0, 9.599102e-1, 2.918724e-1]])
_SPLIT = 0.75


def _prolate_spheroidal_function(u):
    """Returns (grdsf(nu), (1 - nu^2) * grdsf(nu)) for nu = |u|; 0 outside [0, 1]."""
    nu = np.abs(np.asarray(u, dtype=np.float64))
    upper = (nu >= _SPLIT) & (nu <= 1.0)
    inside = (nu >= 0.0) & (nu <= 1.0)
    piece = upper.astype(np.int64)
    edge = np.where(upper, 1.0, np.where(inside, _SPLIT, 0.0))
    x = nu ** 2 - edge ** 2
    num = _NUM[piece, 0]
    for k in range(1, _NUM.shape[1]):
        num = num + _NUM[piece, k] * np.power(x, k)
    den = _DEN[piece, 0]
    for k in range(1, _DEN.shape[1]):
        den = den + _DEN[piece, k] * np.power(x, k)
    grdsf = np.zeros(nu.shape, dtype=np.float64)
    ok = den > 0.0
    grdsf[ok] = num[ok] / den[ok]
    grdsf[nu > 1.0] = 0.0
    return grdsf, (1 - nu ** 2) * grdsf


def _create_prolate_spheroidal_kernel_1D(oversampling, support):
    """Half tap table: taps at nu = k / (oversampling * (support//2)), zero-padded by one more cell."""
    half = support // 2
    table = np.zeros(oversampling * (half + 1))
    nu = np.arange(oversampling * half) / (half * oversampling)
    table[: oversampling * half] = _prolate_spheroidal_function(nu)[1]
    return table


def _correcting_1D(n):
    x = (np.arange(int(n)) - int(n) // 2) / int(n)
    return _prolate_spheroidal_function(np.abs(2.0 * x))[0]


def _create_prolate_spheroidal_image_2D(n_xy):
    return np.outer(_correcting_1D(n_xy[0]), _correcting_1D(n_xy[1]))


def _create_prolate_spheroidal_kernel(oversampling, support, n_uv):
    """(kernel, kernel_image) like the reference; the 4-D oversampled kernel is never used downstream
    (make_image.py:109 keeps only the image), so None is returned in its place."""
    return None, _create_prolate_spheroidal_image_2D(n_uv)


def correcting_function_1D(n_uv_padded, image_size):
    """Separable, cropped correcting function: (corr_u[l], corr_v[m]) with
    corr_image[l, m] == corr_u[l] * corr_v[m] == _remove_padding(kernel_image, image_size)[l, m]."""
    out = []
    for n_pad, n in zip(n_uv_padded, image_size):
        n_pad, n = int(n_pad), int(n)
        start = n_pad // 2 - n // 2
        out.append(_correcting_1D(n_pad)[start:start + n].copy())
    return out[0], out[1]
