"""Pins the CPU oracle (oracle/) bit-exactly to outputs of the reference's own numba code.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py from
/root/reference (ngcasa/imaging/_imaging_utils/_standard_grid.py, _aperture_grid.py,
_gridding_convolutional_kernels.py).  CPU only.
"""
import numpy as np
import pytest

from _util import load_golden


def test_ps_tables(oracle):
    d = np.load(__import__("os").path.join(__import__("_util").GOLDEN, "ps_tables.npz"))
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    assert np.array_equal(cgk, d["cgk_1D_os100_s7"])
    # known answers recorded in SURVEY.md section 8c
    assert cgk[0] == 0.9999996673648565 and cgk[100] == 0.5732453907230994
    assert cgk[299] == 2.8663376429830486e-05 and np.all(cgk[300:] == 0)
    assert np.array_equal(oracle._create_prolate_spheroidal_kernel_1D(50, 5), d["cgk_1D_os50_s5"])
    assert np.array_equal(oracle._create_prolate_spheroidal_image_2D([12, 12]), d["corr_image_12x12"])
    assert np.array_equal(oracle._create_prolate_spheroidal_image_2D([15, 13]), d["corr_image_15x13"])
    assert d["corr_image_12x12"][6, 6] == 0.9999993347298236


def test_single_sample_known_answer(oracle):
    d, gp = load_golden("std_single_sample")
    g, s = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"], gp)
    assert np.array_equal(g, d["grid"]) and np.array_equal(s, d["sum_weight"])
    assert (g != 0).sum() == 36
    assert g[0, 0, 42, 27] == 0.9071838775326958 - 0.4535919387663479j
    assert s[0, 0] == 2.67174672491271


@pytest.mark.parametrize("name", ["std_halfway_edges", "std_cube_sq", "std_cube_odd", "std_continuum_sq",
                                  "std_continuum_odd", "std_cube_s5_1pol"])
def test_standard_grid(oracle, name):
    d, gp = load_golden(name)
    g, s = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"], gp)
    assert np.array_equal(g, d["grid"])
    assert np.array_equal(s, d["sum_weight"])
    if "psf_grid" in d:
        gpp = dict(gp, do_psf=True, complex_grid=False)
        g, s = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"], gpp)
        assert np.array_equal(g, d["psf_grid"])
        assert np.array_equal(s, d["psf_sum_weight"])


@pytest.mark.parametrize("name", ["iw_cube_2pol", "iw_continuum_2pol", "iw_cube_1pol", "iw_continuum_1pol"])
def test_imaging_weights(oracle, name):
    d, gp = load_golden(name)
    rho, sw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp)
    assert np.array_equal(rho, d["density"]) and np.array_equal(sw, d["sum_weight"])
    bf = oracle._calculate_briggs_parms(rho, sw, dict(weighting="briggs", robust=0.5))
    np.testing.assert_allclose(bf, d["briggs_factors"], rtol=1e-14)
    iw = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"],
                                                           d["briggs_factors"], d["freq_chan"], gp)
    assert np.array_equal(iw, d["imaging_weight"], equal_nan=True)


@pytest.mark.parametrize("name", ["aperture_cube", "aperture_continuum"])
def test_aperture(oracle, name):
    d, gp = load_golden(name)
    common = (d["uvw"], d["weight"], d["field"], d["gcf_cf_baseline_map"], d["gcf_cf_chan_map"],
              d["gcf_cf_pol_map"])
    g, s = oracle._aperture_grid_numpy_wrap(d["vis"], *common, d["gcf_conv_kernel"], d["gcf_weight_support"],
                                            d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert np.array_equal(g, d["grid"]) and np.array_equal(s, d["sum_weight"])
    g, s = oracle._aperture_psf_grid_numpy_wrap(*common, d["gcf_conv_kernel"], d["gcf_weight_support"],
                                                d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert np.array_equal(g, d["psf_grid"]) and np.array_equal(s, d["psf_sum_weight"])
    g, s = oracle._aperture_weight_grid_numpy_wrap(*common, d["gcf_weight_conv_kernel"], d["gcf_weight_support"],
                                                   d["gcf_phase_gradient"], d["freq_chan"], gp)
    assert np.array_equal(g, d["weight_grid"]) and np.array_equal(s, d["weight_sum_weight"])


def test_degrid_is_adjoint_of_grid(oracle):
    """A7 has no reference implementation (parity unpinned): check <grid(x), y> == <x, degrid(y)>."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(6, 10, 3, 2, 1.0e9, 1.1e9, 300.0, 120.0, seed=4, flag_frac=0.0, bad_rows=False)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(64, d["cell"], chan_mode=mode)
        ones = np.ones_like(d["weight"])
        g, _ = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], ones, d["freq_chan"], cgk, gp)
        rng = np.random.default_rng(0)
        y = rng.standard_normal(g.shape) + 1j * rng.standard_normal(g.shape)
        v = oracle._standard_degrid_numpy_wrap(y, d["uvw"], d["freq_chan"], cgk, gp)
        lhs = np.vdot(y, g)          # <y, grid(x)>
        rhs = np.vdot(v, d["vis"])   # <degrid(y), x>
        assert abs(lhs - rhs) <= 1e-12 * abs(lhs)


def test_mt_driver_matches_single_thread(oracle):
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(6, 16, 8, 2, 1.0e9, 1.1e9, 300.0, 120.0, seed=8)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(64, d["cell"], chan_mode=mode)
        g1, s1 = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
        g4, s4 = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp,
                                                  n_threads=4)
        assert np.max(np.abs(g1 - g4)) <= 1e-13 * np.max(np.abs(g1))
        np.testing.assert_allclose(s1, s4, rtol=1e-13)
