"""GPU parity of grid -> image (A9 cuFFT + crop, A10 correct / normalise) against the numpy restatement."""
import numpy as np
import pytest

from _util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fft():
    import torch
    assert torch.cuda.is_available()
    from cngi_prototype_b200 import _fft
    return _fft


@pytest.mark.parametrize("n_pad,n_img", [((64, 64), (48, 48)), ((61, 75), (50, 63)), ((75, 61), (75, 61)),
                                         ((128, 96), (107, 80)), ((45, 45), (38, 37))])
@pytest.mark.parametrize("kind", ["complex", "real"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_grid_to_image(fft, oracle, n_pad, n_img, kind, prec):
    """Even / odd padded sizes (fftshift != ifftshift for odd), non-square, crop windows of either parity."""
    from cngi_prototype_b200._gridding_convolutional_kernels import correcting_function_1D
    rng = np.random.default_rng(sum(n_pad) + sum(n_img))
    shape = (3, 2) + tuple(n_pad)
    g = rng.standard_normal(shape)
    if kind == "complex":
        g = g + 1j * rng.standard_normal(shape)
    if prec == "f32":
        g = g.astype(np.complex64 if kind == "complex" else np.float32)
    sw = rng.uniform(1, 3, size=(3, 2))
    sw[1, 0] = 0.0   # sum_weight == 0 -> 1 (make_image.py:125)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(n_pad), np.array(n_img))
    ref = oracle.correct_image(oracle.grid_to_uncorrected_image(g.astype(np.complex128 if kind == "complex" else np.float64),
                                                                np.array(n_img)), sw, corr)
    cu, cv = correcting_function_1D(n_pad, n_img)
    img = fft.grid_to_image(g, n_img, sum_weight=sw, corr_u=cu, corr_v=cv)
    assert img.shape == ref.shape
    assert rel_err(img, ref) <= (1e-12 if prec == "f64" else 1e-5)
    raw = fft.grid_to_image(g, n_img)
    ref_raw = oracle.grid_to_uncorrected_image(g.astype(np.complex128 if kind == "complex" else np.float64), np.array(n_img))
    assert rel_err(raw, ref_raw) <= (1e-12 if prec == "f64" else 1e-5)


def test_normalize_with_pb(fft, oracle):
    """_normalize.py:39-89: sinc oversampling correction * PS_CORR_IMAGE * PB, pb_limit mask, f32 round trip."""
    rng = np.random.default_rng(3)
    n_pad, n_img, osamp = (80, 72), (64, 60), (10, 10)
    g = rng.standard_normal((2, 2) + n_pad) + 1j * rng.standard_normal((2, 2) + n_pad)
    sw = rng.uniform(1, 3, size=(2, 2))
    x = np.linspace(-1, 1, n_img[0])[:, None]
    y = np.linspace(-1, 1, n_img[1])[None, :]
    pb = np.exp(-2.5 * (x ** 2 + y ** 2))[None, None] * np.ones((2, 2, 1, 1))   # kernel-side (chan, pol, l, m)
    ps = oracle._create_prolate_spheroidal_image_2D(n_img)
    raw = oracle.grid_to_uncorrected_image(g, np.array(n_img))
    pb_api = np.moveaxis(pb, (0, 1), (2, 3))
    ref = oracle.normalize_image(raw, sw, ps[:, :, None, None] * pb_api, osamp)
    ref[pb_api < 0.2] = 0.0
    ref = ref.astype(np.float32).astype(np.float64)
    c = np.array(n_img) // 2
    sincx = np.sinc(np.arange(-c[0], n_img[0] - c[0]) / (n_img[0] * osamp[0]))
    sincy = np.sinc(np.arange(-c[1], n_img[1] - c[1]) / (n_img[1] * osamp[1]))
    img = fft.grid_to_image(g, n_img, sum_weight=sw, corr_u=sincx, corr_v=sincy, norm_image=ps[None, None] * pb,
                            pb_image=pb, pb_limit=0.2, single_precision_roundtrip=True)
    assert np.array_equal(img == 0, ref == 0)
    assert rel_err(img, ref) <= 2e-7   # one f32 ulp: the round trip can flip on 1e-16 differences


def test_end_to_end_dirty_image_vs_oracle(fft, oracle):
    """make_image chain: grid -> ifft -> crop -> correct, padded 1.2x (odd padded size), vs oracle chain."""
    from cngi_prototype_b200 import synth, _standard_grid
    from cngi_prototype_b200._gridding_convolutional_kernels import (_create_prolate_spheroidal_kernel_1D,
                                                                   correcting_function_1D)
    d = synth.config_c1(n_time=40, n_chan=4)
    n_img = np.array([200, 200])
    n_pad = (n_img * 1.2).astype(int) + np.array([1, 0])   # (241, 240)
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(int(n_pad[0]), d["cell"], chan_mode="cube")
    gp["image_size_padded"] = n_pad
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(n_pad), n_img)
    ref = oracle.correct_image(oracle.grid_to_uncorrected_image(g_ref, n_img), s_ref, corr)
    import torch
    T = [torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")]
    g, s = _standard_grid._standard_grid_numpy_wrap(*T, cgk, gp)
    cu, cv = correcting_function_1D(n_pad, n_img)
    img = fft.grid_to_image(g, n_img, sum_weight=s, corr_u=cu, corr_v=cv)
    assert rel_err(img.cpu().numpy(), ref) <= 1e-12
