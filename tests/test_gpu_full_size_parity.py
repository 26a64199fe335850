"""BASELINE configs 1-4 at FULL size: the CUDA path against the CPU oracle on the same inputs.

The oracle (oracle/cngi_oracle.c, the C restatement of the reference's numba loops, pinned bit for bit to the
reference-generated goldens in tests/golden/) runs multi-threaded the way the reference's dask graph does
(`n_threads=os.cpu_count()`): a whole config takes seconds on the GPU box's host cores.  Bars are north_star's:
support masks (cell indexing + flag / weight masking) bit-exact, values 1e-12 (fp64) / 1e-5 (fp32) of the peak.

  config 1: make_psf + make_image gridding, VLA-like 351 bl x 1000 t x 64 ch x 2 pol = 44.9 M samples, 1024^2, fp64,
            cube and continuum                                  (_standard_grid.py:242-371)
  config 2: Briggs(0.5) imaging weights + gridding, ALMA-like 903 bl x 500 t x 128 ch x 2 pol = 115.6 M samples,
            4096^2, fp32, continuum -- bench.py's workload      (make_imaging_weight.py:144-247, _standard_grid.py:466-518)
  config 3: aperture image / psf / weight gridding, 7-pointing mosaic, 903 bl x 200 t x 64 ch x 2 pol = 23.1 M samples,
            CF 160^2 with supports 9..15, 2048^2, fp64          (_aperture_grid.py:180-291,376-513)
  config 4: degridding predict, 351 bl x 1000 t x 64 ch x 2 pol, 4096^2, S=7, fp64 (no reference implementation:
            the oracle's restatement of the adjoint, parity unpinned -- see DESIGN.md)
"""
import os

import numpy as np
import pytest

from _util import rel_err, same_support

pytestmark = pytest.mark.gpu

N_THREADS = os.cpu_count() or 1


def _cuda(*xs):
    import torch
    return tuple(torch.as_tensor(x).cuda() for x in xs)


@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_config1_full_size_image_and_psf_vs_oracle(oracle, mode):
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg
    d = synth.config_c1()
    assert d["weight"].shape == (1000, 351, 64, 2) and d["vis"].dtype == np.complex128
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(1024, d["cell"], chan_mode=mode)
    gpp = dict(gp, do_psf=True, complex_grid=False)
    vis, uvw, w, freq = _cuda(d["vis"], d["uvw"], d["weight"], d["freq_chan"])
    g, s = sg._standard_grid_numpy_wrap(vis, uvw, w, freq, cgk, gp)
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp,
                                                    n_threads=N_THREADS)
    gh = g.cpu().numpy()
    del g
    assert same_support(gh, g_ref)
    assert rel_err(gh, g_ref) <= 1e-12 and rel_err(s.cpu().numpy(), s_ref) <= 1e-12
    del gh, g_ref
    p, ps = sg._standard_grid_psf_numpy_wrap(uvw, w, freq, cgk, gpp)
    p_ref, ps_ref = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gpp,
                                                         n_threads=N_THREADS)
    ph = p.cpu().numpy()
    assert same_support(ph, p_ref) and rel_err(ph, p_ref) <= 1e-12 and rel_err(ps.cpu().numpy(), ps_ref) <= 1e-12
    # the fused image + psf pass on the same inputs
    g2, s2, p2, ps2 = sg.standard_grid_image_psf(vis, uvw, w, freq, cgk, gp, force_fused=True)
    assert same_support(p2.cpu().numpy(), p_ref) and rel_err(p2.cpu().numpy(), p_ref) <= 1e-12
    assert rel_err(ps2.cpu().numpy(), ps_ref) <= 1e-12 and rel_err(s2.cpu().numpy(), s_ref) <= 1e-12
    torch.cuda.empty_cache()


def test_config2_full_size_weights_and_gridding_vs_oracle(oracle):
    """bench.py's step at bench.py's size.  The oracle computes in fp64 from the same fp32 samples (upcasting is exact)."""
    import torch
    from cngi_prototype_b200 import synth, _imaging_weight as iw, _standard_grid as sg
    d = synth.config_c2()
    assert d["weight"].size == 500 * 903 * 128 * 2 and d["vis"].dtype == np.complex64
    n = 4096
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    gpw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                               complex_grid=False, do_imaging_weight=True)
    parms = {"weighting": "briggs", "robust": 0.5}
    vis, uvw, w, freq = _cuda(d["vis"], d["uvw"], d["weight"], d["freq_chan"])
    rho, sw = iw.imaging_weight_grid(uvw, w, freq, gpw)
    bf = iw.calculate_briggs_parms(rho, sw, parms)
    iwt = iw._standard_imaging_weight_degrid_numpy_wrap(rho, uvw, w, bf, freq, gpw, kernel_side_layout=True)
    g, s = sg._standard_grid_numpy_wrap(vis, uvw, iwt, freq, cgk, gp)

    rho_ref, sw_ref = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gpw,
                                                           n_threads=N_THREADS)
    rho_h = rho.cpu().numpy()
    assert same_support(rho_h, rho_ref) and rel_err(rho_h, rho_ref) <= 1e-12    # density accumulates in fp64
    assert rel_err(sw.cpu().numpy(), sw_ref) <= 1e-12
    bf_ref = oracle._calculate_briggs_parms(rho_ref, sw_ref, parms)
    assert rel_err(bf.cpu().numpy(), bf_ref) <= 1e-12
    iw_ref = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_ref, (0, 1), (2, 3)), d["uvw"], d["weight"],
                                                               bf_ref, d["freq_chan"], gpw, n_threads=N_THREADS)
    iw_h = iwt.cpu().numpy()
    assert iw_h.dtype == np.float32
    assert np.array_equal(np.isnan(iw_h), np.isnan(iw_ref)) and np.array_equal(iw_h == 0, iw_ref == 0)
    fin = np.isfinite(iw_ref)
    assert np.max(np.abs(iw_h[fin] - iw_ref[fin])) / np.max(np.abs(iw_ref[fin])) <= 1e-6
    del rho_h, rho_ref, fin
    # the oracle grids the oracle's own fp64 imaging weights (the whole chain is compared, not one link)
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw_ref, d["freq_chan"], cgk, gp,
                                                    n_threads=N_THREADS)
    gh = g.cpu().numpy()
    assert same_support(gh, g_ref)
    assert rel_err(gh, g_ref) <= 1e-5 and rel_err(s.cpu().numpy(), s_ref) <= 1e-6
    del gh, g
    # the same step with the weight degrid folded into the gridder (what bench.py's pipeline and the host-array API run)
    out = torch.empty_like(w)
    g2, s2 = sg.standard_grid(vis, uvw, w, freq, cgk, gp, False, True,
                              imaging_weight_from=dict(density=rho, briggs_factors=bf, grid_parms=gpw, out=out))
    assert torch.equal(torch.nan_to_num(out, nan=-1.0), torch.nan_to_num(iwt, nan=-1.0))
    gh = g2.cpu().numpy()
    assert same_support(gh, g_ref) and rel_err(gh, g_ref) <= 1e-5 and rel_err(s2.cpu().numpy(), s_ref) <= 1e-6
    torch.cuda.empty_cache()


def test_config3_full_size_aperture_gridders_vs_oracle(oracle):
    import torch
    from cngi_prototype_b200 import synth, _aperture_grid as ap
    d = synth.config_c2(n_time=200, n_chan=64, dtype="f64")
    gcf = synth.make_mosaic_gcf(d["n_baseline"], 64, 2, n_field=7)
    field = synth.mosaic_field_column(200, d["n_baseline"], gcf["field_id"])
    n = 2048
    gp = synth.grid_parms_for(n, d["cell"] * 1.1, chan_mode="continuum")
    gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
    common = (d["uvw"], d["weight"], field, gcf["cf_baseline_map"], gcf["cf_chan_map"], gcf["cf_pol_map"])
    tail = (gcf["weight_support"], gcf["phase_gradient"], d["freq_chan"], gp)
    common_t = _cuda(*common)
    ck_t, wck_t, pg_t, freq_t, vis_t = _cuda(gcf["conv_kernel"], gcf["weight_conv_kernel"], gcf["phase_gradient"],
                                             d["freq_chan"], d["vis"])
    tail_t = (gcf["weight_support"], pg_t, freq_t, gp)
    g, s = ap._aperture_grid_numpy_wrap(vis_t, *common_t, ck_t, *tail_t)
    g_ref, s_ref = oracle._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail)
    gh = g.cpu().numpy()
    assert same_support(gh, g_ref) and rel_err(gh, g_ref) <= 1e-12 and rel_err(s.cpu().numpy(), s_ref) <= 1e-12
    g, s = ap._aperture_psf_grid_numpy_wrap(*common_t, ck_t, *tail_t)
    g_ref, s_ref = oracle._aperture_psf_grid_numpy_wrap(*common, gcf["conv_kernel"], *tail)
    gh = g.cpu().numpy()
    assert same_support(gh, g_ref) and rel_err(gh, g_ref) <= 1e-12 and rel_err(s.cpu().numpy(), s_ref) <= 1e-12
    g, s = ap._aperture_weight_grid_numpy_wrap(*common_t, wck_t, *tail_t)
    g_ref, s_ref = oracle._aperture_weight_grid_numpy_wrap(*common, gcf["weight_conv_kernel"], *tail)
    gh = g.cpu().numpy()
    assert same_support(gh, g_ref) and rel_err(gh, g_ref) <= 1e-12 and rel_err(s.cpu().numpy(), s_ref) <= 1e-12
    torch.cuda.empty_cache()


def test_config4_full_size_degrid_vs_oracle(oracle):
    import torch
    from cngi_prototype_b200 import synth, _standard_degrid as sd
    d = synth.config_c4()
    d["uvw"].reshape(-1, 3)[::997, 0] = np.nan                                   # a few rows without a uv point
    n = 4096
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    rng = np.random.default_rng(7)
    y = rng.standard_normal((1, 2, n, n)) + 1j * rng.standard_normal((1, 2, n, n))
    y_t, uvw, freq = _cuda(y, d["uvw"], d["freq_chan"])
    for normalize in (False, True):
        v = sd._standard_degrid_numpy_wrap(y_t, uvw, freq, cgk, gp, normalize=normalize).cpu().numpy()
        v_ref = oracle._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp, normalize=normalize)
        assert v.shape == (1000, 351, 64, 2)
        assert np.array_equal(v == 0, v_ref == 0)                                 # skipped samples are exactly 0
        assert rel_err(v, v_ref) <= 1e-12
    torch.cuda.empty_cache()
