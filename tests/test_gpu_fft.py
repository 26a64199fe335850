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


@pytest.mark.parametrize("n_pad,n_img,dtype", [((64, 64), (52, 52), np.float64), ((75, 61), (61, 51), np.float64),
                                               ((61, 75), (61, 75), np.float64), ((96, 80), (80, 66), np.float32)])
def test_image_to_grid_vs_numpy(fft, oracle, n_pad, n_img, dtype):
    """fftshift(fft2(ifftshift(pad(img / corr)))): the inverse of make_image.py:116-130 (even, odd, non-square, no pad)."""
    rng = np.random.default_rng(11)
    img = rng.standard_normal(n_img + (2, 2)).astype(dtype)            # API-side (l, m, chan, pol)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(list(n_pad)), n_img)
    from cngi_prototype_b200._gridding_convolutional_kernels import correcting_function_1D
    cu, cv = correcting_function_1D(n_pad, n_img)
    np.testing.assert_allclose(cu[:, None] * cv[None, :], corr, rtol=1e-12)
    padded = np.zeros(n_pad + (2, 2))
    s0, s1 = n_pad[0] // 2 - n_img[0] // 2, n_pad[1] // 2 - n_img[1] // 2
    padded[s0:s0 + n_img[0], s1:s1 + n_img[1]] = img.astype(np.float64) / corr[:, :, None, None]
    ref = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(padded, axes=(0, 1)), axes=(0, 1)), axes=(0, 1))
    ref = np.moveaxis(ref, (2, 3), (0, 1))                               # kernel-side (chan, pol, u, v)
    got = fft.image_to_grid(img, n_pad, corr_u=cu, corr_v=cv)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < (2e-6 if dtype == np.float32 else 1e-13)
    # round trip through grid_to_image (unnormalised forward and inverse: factor n_u*n_v is already in grid_to_image's "* N")
    back = fft.grid_to_image(got, n_img, corr_u=cu, corr_v=cv)
    scale = n_pad[0] * n_pad[1]      # grid_to_image applies the reference's "* (n_u * n_v)" (make_image.py:120)
    assert rel_err(back / scale * (cu[:, None] * cv[None, :])[:, :, None, None] ** 2, img.astype(np.float64)) < (1e-5 if dtype == np.float32 else 1e-12)


def test_predict_modelvis_image(oracle):
    """BASELINE config 4 chain: model image -> / PS image -> pad -> FFT -> degrid (normalised) == the oracle chain, and
    close to the analytic visibilities of the point sources."""
    from cngi_prototype_b200 import synth, imaging
    d = synth.config_c4(n_time=12, n_chan=3)
    n = 300
    cell_as = float(abs(d["cell"]) * 3600 * 180 / np.pi)
    grid_parms = dict(image_size=[n, n], cell_size=[cell_as, cell_as], fft_padding=1.2, chan_mode="continuum")
    model = np.zeros((n, n, 1, 2))
    src = [(n // 2 + 20, n // 2 - 31, 1.0), (n // 2 - 50, n // 2 + 12, 0.6), (n // 2, n // 2, 0.25)]
    for (i, j, amp) in src:
        model[i, j, 0, :] = amp
    out = imaging.predict_modelvis_image({"MODEL": model}, {"UVW": d["uvw"], "chan": d["freq_chan"]}, grid_parms,
                                         time_chunk=5)
    v = out["MODEL_DATA"]
    assert v.shape == d["uvw"].shape[:2] + (3, 2)
    # oracle chain
    n_pad = int(1.2 * n)
    cell = np.array([-cell_as, cell_as]) * np.pi / (3600 * 180)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D([n_pad, n_pad]), [n, n])
    padded = np.zeros((n_pad, n_pad, 1, 2))
    s = n_pad // 2 - n // 2
    padded[s:s + n, s:s + n] = model / corr[:, :, None, None]
    G = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(padded, axes=(0, 1)), axes=(0, 1)), axes=(0, 1))
    gp = dict(chan_mode="continuum", image_size_padded=np.array([n_pad, n_pad]), cell_size=cell, oversampling=100, support=7)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    ref = oracle._standard_degrid_numpy_wrap(np.moveaxis(G, (2, 3), (0, 1)), d["uvw"], d["freq_chan"], cgk, gp, normalize=True)
    assert np.array_equal(v == 0, ref == 0)
    assert rel_err(v, ref) < 1e-11
    us = -(d["freq_chan"] * cell[0] * n_pad) / 299792458.0
    vs = -(d["freq_chan"] * cell[1] * n_pad) / 299792458.0
    up = d["uvw"][:, :, 0, None] * us[None, None, :]
    vp = d["uvw"][:, :, 1, None] * vs[None, None, :]
    analytic = np.zeros(up.shape, dtype=np.complex128)
    for (i, j, amp) in src:
        analytic += amp * np.exp(-2j * np.pi * (up * (i - n // 2) + vp * (j - n // 2)) / n_pad)
    ok = v[..., 0] != 0
    assert ok.mean() > 0.9 and np.abs(v[..., 0][ok] - analytic[ok]).max() < 1e-2


@pytest.mark.parametrize("n_pad,n_img", [((1774, 1774), (1500, 1400)),     # 2 * 887
                                         ((1563, 2084), (1300, 1700)),     # 3 * 521 (generic n1) x 4 * 521
                                         ((983, 2615), (800, 2179)),       # n1 = 1 x 5 * 523
                                         ((4915, 1046), (4096, 871)),      # 5 * 983: 4096 padded by 1.2
                                         ((1966, 10460), (1638, 8716))])   # 2 * 983 x 20 * 523
def test_shared_memory_bluestein_sizes_match_numpy_and_cufft(fft, oracle, n_pad, n_img, monkeypatch):
    """Sides n1 * prime (512 < prime <= 1021, n1 <= 32) of complex64 grids take the shared-memory Bluestein passes
    (csrc/fft_bluestein.cu) instead of cuFFT: same image as numpy.fft and as cuFFT (CNGI_FFT_BLUESTEIN=0) to fp32
    rounding, both directions, complex and real (psf) grids."""
    import torch
    rng = np.random.default_rng(n_pad[0] + n_pad[1])
    g = (rng.standard_normal((2,) + tuple(n_pad)) + 1j * rng.standard_normal((2,) + tuple(n_pad))).astype(np.complex64)
    g = g.reshape((1, 2) + tuple(n_pad))
    g[0, 1, n_pad[0] // 2, n_pad[1] // 2] += 50.0                 # a bright centre cell: a smooth pedestal under the noise
    ref = oracle.grid_to_uncorrected_image(g.astype(np.complex128), np.array(n_img))
    gt = torch.as_tensor(g).cuda()
    fft._plans.clear()
    img = fft.grid_to_image(gt, n_img).cpu().numpy()
    assert rel_err(img, ref) <= 2e-6
    gr = torch.as_tensor(np.ascontiguousarray(g.real)).cuda()
    img_r = fft.grid_to_image(gr, n_img).cpu().numpy()
    ref_r = oracle.grid_to_uncorrected_image(g.real.astype(np.float64), np.array(n_img))
    assert rel_err(img_r, ref_r) <= 2e-6
    # forward: image -> grid (config 4's first step) and back
    model = torch.as_tensor(rng.standard_normal(tuple(n_img) + (1, 2)).astype(np.float32)).cuda()
    grid_b = fft.image_to_grid(model, n_pad).cpu().numpy()
    monkeypatch.setenv("CNGI_FFT_BLUESTEIN", "0")
    fft._plans.clear()
    img_c = fft.grid_to_image(gt, n_img).cpu().numpy()
    grid_c = fft.image_to_grid(model, n_pad).cpu().numpy()
    fft._plans.clear()
    assert rel_err(img, img_c) <= 2e-6 and rel_err(img_c, ref) <= 2e-6
    assert rel_err(grid_b, grid_c) <= 2e-6
    padded = np.zeros(tuple(n_pad))
    s0, s1 = n_pad[0] // 2 - n_img[0] // 2, n_pad[1] // 2 - n_img[1] // 2
    padded[s0:s0 + n_img[0], s1:s1 + n_img[1]] = model[:, :, 0, 1].cpu().numpy()
    ref_g = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(padded)))
    assert rel_err(grid_b[0, 1], ref_g) <= 2e-6
