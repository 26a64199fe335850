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
    # A2 -> A3 -> A4 with the threaded drivers (the CPU arm of bench.py): A4's tasks write disjoint rows -> bit-identical
    gp_iw = synth.grid_parms_for(64, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)
    rho1, sw1 = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp_iw)
    rho3, sw3 = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp_iw, n_threads=3)
    assert np.array_equal(rho1 != 0, rho3 != 0)
    np.testing.assert_allclose(rho1, rho3, rtol=1e-13)
    bf = oracle._calculate_briggs_parms(rho1, sw1, {"weighting": "briggs", "robust": 0.5})
    g_api = np.moveaxis(rho1, (0, 1), (2, 3))
    iw1 = oracle._standard_imaging_weight_degrid_numpy_wrap(g_api, d["uvw"], d["weight"], bf, d["freq_chan"], gp_iw)
    for nt in (3, 16):   # 16 > n_time = 6: one task per integration
        iwn = oracle._standard_imaging_weight_degrid_numpy_wrap(g_api, d["uvw"], d["weight"], bf, d["freq_chan"], gp_iw,
                                                                n_threads=nt)
        assert np.array_equal(iw1, iwn, equal_nan=True)


# ---- N3 direction_rotate --------------------------------------------------------
def _phase_bound(d):
    """The reference's own reproducibility limit: its BLAS dot and complex multiply may or may not fuse, so the
    phase argument y = 2 pi d f / c is only defined to a few ulp(y); |delta vis| <= |vis| * few * eps * |y|."""
    y = 2 * np.pi * np.abs(d["uvw_rot"][:, :, :3]).sum(-1).max() * np.abs(d["phase_rotation"]).max() * d["freq_chan"].max() / 299792458.0
    return 1e-15 + 8 * np.finfo(np.float64).eps * y


@pytest.mark.parametrize("name", ["direction_rotate_ctr_sp", "direction_rotate_full_dp"])
def test_direction_rotate(oracle, name):
    import os
    from _util import GOLDEN, rel_err
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    ctr, sp = bool(d["ctr"]), bool(d["sp"])
    R, P, ids = oracle.calc_rotation_mats(d["field"], d["table_ids"], d["table_dirs"], d["new_phase_center"], ctr)
    assert np.array_equal(ids, d["rot_field_id"])
    np.testing.assert_allclose(R, d["uvw_rotmat"], rtol=0, atol=1e-15)      # scipy builds them from quaternions
    np.testing.assert_allclose(P, d["phase_rotation"], rtol=0, atol=1e-18)
    ur = oracle.apply_rotation_matrix(d["uvw"], d["field"], d["uvw_rotmat"], d["rot_field_id"])
    assert rel_err(ur, d["uvw_rot"]) < 4e-16
    vr = oracle.apply_phasor(d["vis"], d["uvw_rot"], d["field"], d["freq_chan"], d["phase_rotation"],
                             d["rot_field_id"], ctr, sp)
    assert np.array_equal(np.isnan(vr), np.isnan(d["vis_rot"]))
    m = ~np.isnan(vr)
    assert np.max(np.abs(vr[m] - d["vis_rot"][m])) <= _phase_bound(d) * np.abs(d["vis"][m]).max()
    if sp:   # after the complex64 round trip the two agree bit for bit
        assert np.array_equal(vr[m], d["vis_rot"][m])


def test_direction_rotate_phase_order():
    """numpy evaluates `2.0*1j*np.pi*d*f/c` as ((2 pi d) f) (1/c): the form the oracle and the CUDA kernel use."""
    rng = np.random.default_rng(3)
    dd = rng.normal(0, 0.3, (200, 1))
    f = np.linspace(1e9, 350e9, 16)[None, :]
    y = (2.0 * 1j * np.pi * dd * f / 299792458.0)
    assert np.array_equal(y.imag, (((2.0 * np.pi) * dd) * f) * (1.0 / 299792458.0)) and not y.real.any()
    e = np.exp(y)
    assert np.array_equal(e.real, np.cos(y.imag)) and np.array_equal(e.imag, np.sin(y.imag))


# ---- N2 make_gridding_convolution_function ----------------------------------------
def _gcf_parms_from(d, tag):
    return dict(function=tag, list_dish_diameters=d["dish"], list_blockage_diameters=d["blockage"],
                unique_ant_indx=d["unique_ant_indx"], basline_ant=d["baseline_ant"], freq_chan=d["freq_chan"],
                pol=np.array([0, 1]), oversampling=d["oversampling"], max_support=d["max_support"],
                field_phase_dir=np.array([[1.0, 0.5], [1.0001, 0.5001]]), phase_center=np.array([1.0, 0.5]))


@pytest.mark.parametrize("tag", ["casa_airy", "airy"])
def test_gcf(oracle, tag):
    import os
    from _util import GOLDEN
    d = np.load(os.path.join(GOLDEN, "gcf_%s.npz" % tag))
    g = oracle.make_gridding_convolution_function(_gcf_parms_from(d, tag), dict(
        image_size=d["n_pad"], image_size_padded=d["n_pad"], cell_size=d["cell_size"]))
    for key, name in (("conv_kernel", "CONV_KERNEL"), ("weight_conv_kernel", "WEIGHT_CONV_KERNEL"),
                      ("support", "SUPPORT"), ("cf_baseline_map", "CF_BASELINE_MAP"), ("cf_chan_map", "CF_CHAN_MAP"),
                      ("pb_freq", "pb_freq"), ("pb_ant_pairs", "pb_ant_pairs")):
        assert np.array_equal(g[name], d[key]), key
    assert g["PHASE_GRADIENT"].shape == (2, 60, 60) and np.all(g["PHASE_GRADIENT"][0] == 1.0)


def test_gcf_chan_maps(oracle):
    import os
    from _util import GOLDEN
    c = np.load(os.path.join(GOLDEN, "gcf_chan_maps.npz"))
    for k in range(4):
        m, pf = oracle.create_cf_chan_map(c["f%d" % k], float(c["tol%d" % k]))
        assert np.array_equal(m, c["map%d" % k]) and np.array_equal(pf, c["pbf%d" % k])


def test_sin_projection_known_answer(oracle):
    """The only world2pix vector the reference holds: the CASA trace quoted in the comments at
    make_gridding_convolution_function.py:565-577 (crpix 120, cdelt -/+5e-05 deg)."""
    deg = np.pi / 180
    off = oracle.sin_world2pix_offset(np.array([[-179.5337374791666889 * deg, -18.863873258333338612 * deg]]),
                                      np.array([180.46846189999996568 * deg, -18.863873247222226581 * deg]),
                                      np.array([-5e-05 * deg, 5e-05 * deg]))
    np.testing.assert_allclose(off[0] + 120, [161.6249842951184803, 119.99951947142589859], rtol=0, atol=2e-9)


@pytest.mark.parametrize("tag", ["casa_airy", "airy"])
def test_make_pb_patterns(oracle, tag):
    """_airy_disk / _casa_airy_disk (_make_pb_symmetric.py:26,79), ipower 1 and 2: bit-exact."""
    import os
    from _util import GOLDEN
    d = np.load(os.path.join(GOLDEN, "pb_%s.npz" % tag))
    gp = dict(image_size=d["image_size"], image_center=d["image_center"], cell_size=d["cell_size"])
    for ipower, key in ((2, "pb"), (1, "voltage")):
        o = oracle.airy_disk(d["freq_chan"], 2, dict(list_dish_diameters=d["dish"], list_blockage_diameters=d["blockage"],
                                                     ipower=ipower), gp, casa=(tag == "casa_airy"))
        assert np.array_equal(o, d[key])
