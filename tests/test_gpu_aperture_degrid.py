"""GPU parity of the aperture gridders (A5/A6) and the degrid predict (A7)."""
import numpy as np
import pytest

from _util import load_golden, rel_err, same_support

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ap():
    import torch
    assert torch.cuda.is_available()
    from cngi_prototype_b200 import _aperture_grid
    return _aperture_grid


def _common(d):
    return (d["uvw"], d["weight"], d["field"], d["gcf_cf_baseline_map"], d["gcf_cf_chan_map"], d["gcf_cf_pol_map"])


@pytest.mark.parametrize("name", ["aperture_cube", "aperture_continuum"])
def test_aperture_golden(ap, name):
    d, gp = load_golden(name)
    g, s = ap._aperture_grid_numpy_wrap(d["vis"], *_common(d), d["gcf_conv_kernel"], d["gcf_weight_support"],
                                        d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert same_support(g, d["grid"]) and rel_err(g, d["grid"]) <= 1e-12 and rel_err(s, d["sum_weight"]) <= 1e-12
    g, s = ap._aperture_psf_grid_numpy_wrap(*_common(d), d["gcf_conv_kernel"], d["gcf_weight_support"],
                                            d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert same_support(g, d["psf_grid"]) and rel_err(g, d["psf_grid"]) <= 1e-12
    assert rel_err(s, d["psf_sum_weight"]) <= 1e-12
    g, s = ap._aperture_weight_grid_numpy_wrap(*_common(d), d["gcf_weight_conv_kernel"], d["gcf_weight_support"],
                                               d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert same_support(g, d["weight_grid"]) and rel_err(g, d["weight_grid"]) <= 1e-12
    assert rel_err(s, d["weight_sum_weight"]) <= 1e-12


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("max_support,n_cf_pol", [(15, 1), (7, 1), (17, 1), (11, 2)])
def test_mosaic_vs_oracle(ap, oracle, prec, max_support, n_cf_pol):
    """Config-3 shape at reduced size: 7 pointings, CF 160x160 (oversampling 10, max support 15), variable supports.
    max_support 7 -> 8-wide register windows, 15 -> 16-wide, 17 -> the per-tap fallback kernel; n_cf_pol 2 -> the two
    polarisations use different convolution functions."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(10, 28, 6, 2, 345e9, 347e9, 300.0, 6.0, seed=31, dtype=prec)
    gcf = synth.make_mosaic_gcf(d["n_baseline"], 6, 2, n_field=7, max_support=(max_support, max_support), n_cf_pol=n_cf_pol)
    fld = synth.mosaic_field_column(28, d["n_baseline"], gcf["field_id"])
    tol = 1e-12 if prec == "f64" else 1e-5
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(256, d["cell"] * 1.25, chan_mode=mode)
        gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
        common = (d["uvw"], d["weight"], fld, gcf["cf_baseline_map"], gcf["cf_chan_map"], gcf["cf_pol_map"])
        tail = (gcf["weight_support"], gcf["phase_gradient"], d["freq_chan"], gp)
        g_ref, s_ref = oracle._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail)
        g, s = ap._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail)
        assert same_support(g, g_ref) and rel_err(g, g_ref) <= tol and rel_err(s, s_ref) <= tol
        g_ref, s_ref = oracle._aperture_psf_grid_numpy_wrap(*common, gcf["conv_kernel"], *tail)
        g, s = ap._aperture_psf_grid_numpy_wrap(*common, gcf["conv_kernel"], *tail)
        assert same_support(g, g_ref) and rel_err(g, g_ref) <= tol and rel_err(s, s_ref) <= tol
        g_ref, s_ref = oracle._aperture_weight_grid_numpy_wrap(*common, gcf["weight_conv_kernel"], *tail)
        g, s = ap._aperture_weight_grid_numpy_wrap(*common, gcf["weight_conv_kernel"], *tail)
        assert same_support(g, g_ref) and rel_err(g, g_ref) <= tol and rel_err(s, s_ref) <= tol


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("support,oversampling", [(7, 100), (5, 50), (9, 20)])
def test_degrid_vs_oracle_and_adjoint(oracle, prec, support, oversampling):
    """A7 has no reference implementation (parity unpinned): oracle restatement + adjointness with the gridder."""
    from cngi_prototype_b200 import synth, _standard_degrid, _standard_grid
    d = synth.make_vis_set(8, 30, 5, 2, 1e9, 1.1e9, 300.0, 120.0, seed=41, dtype=prec)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(oversampling, support)
    rng = np.random.default_rng(0)
    tol = 1e-12 if prec == "f64" else 1e-5
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(128, d["cell"], chan_mode=mode, support=support, oversampling=oversampling)
        n_ic = 5 if mode == "cube" else 1
        y = rng.standard_normal((n_ic, 2, 128, 128)) + 1j * rng.standard_normal((n_ic, 2, 128, 128))
        if prec == "f32":
            y = y.astype(np.complex64)
        v_ref = oracle._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp)
        vn_ref = oracle._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, normalize=True)
        for algo in ((1, 2) if support <= 7 else (1,)):      # gather kernel, register-window kernel (supports 3/5/7)
            v = _standard_degrid._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, algorithm=algo)
            assert np.array_equal(v == 0, v_ref == 0), algo   # skipped samples are exactly 0
            assert rel_err(v, v_ref) <= tol, algo
            vn = _standard_degrid._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, normalize=True,
                                                              algorithm=algo)
            assert rel_err(vn, vn_ref) <= tol, algo
            # a single-pol predict from a two-pol model grid
            v1 = _standard_degrid._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, n_pol=1, algorithm=algo)
            assert rel_err(v1[..., 0], v_ref[..., 0]) <= tol, algo
        v = _standard_degrid._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp)
        # <y, grid(x)> == <degrid(y), x> with unit weights and unflagged data
        x = np.nan_to_num(d["vis"], nan=0.5)
        ones = np.ones_like(d["weight"])
        g, _ = _standard_grid._standard_grid_numpy_wrap(x, d["uvw"], ones, d["freq_chan"], cgk, gp)
        lhs = np.vdot(y.astype(np.complex128), g.astype(np.complex128))
        rhs = np.vdot(v.astype(np.complex128), x.astype(np.complex128))
        assert abs(lhs - rhs) <= (1e-11 if prec == "f64" else 1e-4) * abs(lhs)


def test_degrid_point_source_analytic(oracle):
    """Degridding the FT of a PS-pre-corrected point-source image reproduces A*exp(-2 pi i (u l + v m)) (config 4 idea)."""
    from cngi_prototype_b200 import synth, _standard_degrid
    d = synth.config_c4(n_time=20, n_chan=4)
    n = 512
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    corr = oracle._create_prolate_spheroidal_image_2D([n, n])
    img = np.zeros((n, n))
    src = [(n // 2 + 20, n // 2 - 31, 1.0), (n // 2 - 50, n // 2 + 12, 0.6)]
    for (i, j, amp) in src:
        img[i, j] = amp / corr[i, j]
    # inverse of make_image.py:116: G = fftshift(fft2(ifftshift(img)))
    G = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(img)))
    y = np.repeat(G[None, None], 2, axis=1)
    v = _standard_degrid._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, normalize=True)
    # analytic: grid coordinate u_pix = uvw*uv_scale, pixel offset (i - n/2) <-> exp(-2 pi i u_pix*(i-n/2)/n)
    us = -(d["freq_chan"] * gp["cell_size"][0] * n) / 299792458.0
    vs = -(d["freq_chan"] * gp["cell_size"][1] * n) / 299792458.0
    up = d["uvw"][:, :, 0, None] * us[None, None, :]
    vp = d["uvw"][:, :, 1, None] * vs[None, None, :]
    model = np.zeros(up.shape, dtype=np.complex128)
    for (i, j, amp) in src:
        model += amp * np.exp(-2j * np.pi * (up * (i - n // 2) + vp * (j - n // 2)) / n)
    ok = v[..., 0] != 0
    err = np.abs(v[..., 0][ok] - model[ok]).max()
    assert ok.mean() > 0.95 and err < 1e-2   # limited by the PS kernel's aliasing rejection, not by arithmetic


def test_aperture_bulk_copy_tap_ring_matches_oracle():
    """CNGI_APERTURE_BULK=1: the W x W tap block of each sample is fetched with cp.async.bulk (+ mbarrier) into a
    shared-memory ring instead of 16 L2 loads per lane; same grid as the oracle (the knob is read once per process, hence
    the subprocess)."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from _util import rel_err, same_support
from oracle import oracle
from cngi_prototype_b200 import synth, _aperture_grid as ap
for prec, ms, ncp in (("f32", 15, 1), ("f64", 11, 2), ("f32", 7, 1)):
    d = synth.make_vis_set(10, 28, 6, 2, 345e9, 347e9, 300.0, 6.0, seed=31, dtype=prec)
    gcf = synth.make_mosaic_gcf(d["n_baseline"], 6, 2, n_field=7, max_support=(ms, ms), n_cf_pol=ncp)
    fld = synth.mosaic_field_column(28, d["n_baseline"], gcf["field_id"])
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(256, d["cell"] * 1.25, chan_mode=mode)
        gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
        common = (d["uvw"], d["weight"], fld, gcf["cf_baseline_map"], gcf["cf_chan_map"], gcf["cf_pol_map"])
        tail = (gcf["weight_support"], gcf["phase_gradient"], d["freq_chan"], gp)
        g_ref, s_ref = oracle._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail)
        g, s = ap._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail)
        tol = 1e-12 if prec == "f64" else 1e-5
        assert same_support(g, g_ref) and rel_err(g, g_ref) <= tol and rel_err(s, s_ref) <= tol, (prec, ms, mode)
print("bulk ok")
''' % (os.path.dirname(os.path.abspath(__file__)), os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CNGI_APERTURE_BULK="1"), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "bulk ok" in r.stdout, r.stdout[-3000:]
