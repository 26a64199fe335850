"""GPU parity of the standard gridder (A1/A2) against reference-generated goldens and the CPU oracle.

Bars (BASELINE.json north_star): cell indexing / flag+weight masking bit-exact (identical support masks),
values within 1e-12 (fp64) or 1e-5 (fp32) of max|reference|.  Calls go through the C ABI (ctypes).
"""
import numpy as np
import pytest

from _util import load_golden, rel_err, same_support

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
ALGOS = {"naive": 1, "track": 2, "shift": 3, "window": 4}


@pytest.fixture(scope="module")
def sg():
    import torch
    assert torch.cuda.is_available(), "these tests need a B200"
    from cngi_prototype_b200 import _standard_grid, _lib
    _lib.require_device()
    return _standard_grid


def _cast(d, prec):
    if prec == "f64":
        return d["vis"], d["weight"]
    return d["vis"].astype(np.complex64), d["weight"].astype(np.float32)


def _check(g, s, g_ref, s_ref, tol):
    assert g.shape == g_ref.shape and s.shape == s_ref.shape
    assert same_support(g, g_ref), "touched-cell mask differs (indexing / masking must be bit-exact)"
    assert rel_err(g, g_ref) <= tol, rel_err(g, g_ref)
    assert rel_err(s, s_ref) <= tol, rel_err(s, s_ref)


GOLDENS = ["std_single_sample", "std_halfway_edges", "std_cube_sq", "std_cube_odd", "std_continuum_sq",
           "std_continuum_odd", "std_cube_s5_1pol"]


@pytest.mark.parametrize("name", GOLDENS)
@pytest.mark.parametrize("algo", ["naive", "track", "window"])
@pytest.mark.parametrize("path", ["host", "device"])
def test_golden_fp64(sg, name, algo, path):
    import torch
    d, gp = load_golden(name)
    args = [d["vis"], d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"]]
    if path == "device":
        args = [torch.as_tensor(a).cuda() for a in args]
    g, s = sg._standard_grid_numpy_wrap(*args, gp, algorithm=ALGOS[algo])
    if path == "device":
        g, s = g.cpu().numpy(), s.cpu().numpy()
    _check(g, s, d["grid"], d["sum_weight"], TOL["f64"])
    if "psf_grid" in d:
        gpp = dict(gp, do_psf=True, complex_grid=False)
        g, s = sg._standard_grid_psf_numpy_wrap(*args[1:], gpp, algorithm=ALGOS[algo])
        if path == "device":
            g, s = g.cpu().numpy(), s.cpu().numpy()
        _check(g, s, d["psf_grid"], d["psf_sum_weight"], TOL["f64"])


@pytest.mark.parametrize("name", ["std_cube_sq", "std_continuum_odd", "std_halfway_edges"])
@pytest.mark.parametrize("algo", ["naive", "track", "window"])
def test_golden_fp32(sg, oracle, name, algo):
    d, gp = load_golden(name)
    vis, w = _cast(d, "f32")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(vis, d["uvw"], w, d["freq_chan"], d["cgk_1D"], gp)
    g, s = sg._standard_grid_numpy_wrap(vis, d["uvw"], w, d["freq_chan"], d["cgk_1D"], gp, algorithm=ALGOS[algo])
    assert g.dtype == np.complex64
    _check(g, s, g_ref, s_ref, TOL["f32"])


@pytest.mark.parametrize("mode", ["cube", "continuum"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_vla_like_vs_oracle(sg, oracle, mode, prec):
    """Config-1 geometry at reduced size: 27 antennas, 96 times, 16 chan, 2 pol, with flags / zero+NaN weights / NaN uvw."""
    from cngi_prototype_b200 import synth
    d = synth.config_c1(n_time=96, n_chan=16)
    vis, w = _cast(d, prec)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(512, d["cell"], chan_mode=mode)
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(vis, d["uvw"], w, d["freq_chan"], cgk, gp, n_threads=8)
    for algo in ALGOS.values():
        g, s = sg._standard_grid_numpy_wrap(vis, d["uvw"], w, d["freq_chan"], cgk, gp, algorithm=algo)
        _check(g, s, g_ref, s_ref, TOL[prec])
    gpp = dict(gp, do_psf=True, complex_grid=False)
    g_ref, s_ref = oracle._standard_grid_psf_numpy_wrap(d["uvw"], w, d["freq_chan"], cgk, gpp, n_threads=8)
    for algo in ALGOS.values():
        g, s = sg._standard_grid_psf_numpy_wrap(d["uvw"], w, d["freq_chan"], cgk, gpp, algorithm=algo)
        _check(g, s, g_ref, s_ref, TOL[prec])


def test_fused_flag_equals_nan_data(sg, oracle):
    """flag != 0 must behave exactly like DATA = NaN (cngi/vis/apply_flags.py:53)."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(8, 40, 6, 2, 1e9, 1.1e9, 300.0, 90.0, seed=5, flag_frac=0.0)
    rng = np.random.default_rng(1)
    flag = rng.random(d["vis"].shape) < 0.1
    vis_nan = d["vis"].copy()
    vis_nan[flag] = np.nan
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(128, d["cell"], chan_mode="cube")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(vis_nan, d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    for algo in ALGOS.values():
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp,
                                            flag=flag.astype(np.uint8), algorithm=algo)
        _check(g, s, g_ref, s_ref, TOL["f64"])


@pytest.mark.parametrize("support,oversampling", [(3, 20), (5, 50), (9, 100), (11, 30), (13, 40), (15, 25), (17, 10)])
@pytest.mark.parametrize("n_pol", [1, 2, 3, 4])
def test_supports_and_pol_counts(sg, oracle, support, oversampling, n_pol):
    """The reference is support-generic (_standard_grid.py:344-360): 8-wide register windows (3, 5), 16-wide ones (9, 11,
    13, 15), the per-sample fallback (17), odd pol counts, ragged channel spans."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(7, 33, 5, n_pol, 1e9, 1.2e9, 300.0, 200.0, seed=support * 10 + n_pol)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(oversampling, support)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(96, d["cell"], chan_mode=mode, support=support, oversampling=oversampling)
        g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
        _check(g, s, g_ref, s_ref, TOL["f64"])


@pytest.mark.parametrize("support", [9, 11, 13, 15])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_wide_window_kernel_supports_9_to_15(sg, oracle, support, prec):
    """16-wide register windows (csrc/standard_grid_window16_*.cu: doubled tap rows + sub-vector rotation copies): image and
    psf mode, cube and continuum, forced through ALGO_WINDOW, time-chunked accumulation, odd non-square grid, against the
    oracle; support 9 also against the track kernel it replaces as the default."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(9, 41, 19, 2, 1e9, 1.15e9, 300.0, 150.0, seed=100 + support, dtype=prec)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, support)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(96, d["cell"], chan_mode=mode, support=support, oversampling=100)
        gp["image_size_padded"] = np.array([117, 96])
        g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp, algorithm=ALGOS["window"])
        _check(g, s, g_ref, s_ref, TOL[prec])
        if support == 9:
            g_t, s_t = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp, algorithm=ALGOS["track"])
            _check(g_t, s_t, g_ref, s_ref, TOL[prec])
        gpp = dict(gp, do_psf=True, complex_grid=False)
        p_ref, ps_ref = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gpp)
        pg, ps = sg._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gpp, algorithm=ALGOS["window"])
        _check(pg, ps, p_ref, ps_ref, TOL[prec])


@pytest.mark.parametrize("algo", ["track", "shift", "window"])
@pytest.mark.parametrize("chan_group,time_segment", [(1, 0), (2, 7), (4, 16), (8, 1000), (8, 3)])
def test_invariance_to_work_decomposition(sg, oracle, chan_group, time_segment, algo):
    """Result must not depend on how tracks are cut into work items."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(9, 50, 21, 2, 1e9, 1.3e9, 300.0, 150.0, seed=77)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(160, d["cell"], chan_mode=mode)
        g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp,
                                            algorithm=ALGOS[algo], chan_group=chan_group, time_segment=time_segment)
        _check(g, s, g_ref, s_ref, TOL["f64"])


def test_many_channels_span_several_channel_windows(sg, oracle):
    """More channels than one shared-memory channel table holds (the launcher loops over 2048-channel windows)."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(3, 6, 2100, 1, 1e9, 1.5e9, 300.0, 300.0, seed=15)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(64, d["cell"], chan_mode="continuum")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    for algo in ("track", "shift", "window"):
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp, algorithm=ALGOS[algo])
        _check(g, s, g_ref, s_ref, TOL["f64"])


def test_random_order_uvw_no_coherence(sg, oracle):
    """uvw with no time coherence at all (worst case for the track kernel) must still be exact."""
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(8, 30, 4, 2, 1e9, 1.1e9, 300.0, 90.0, seed=6)
    rng = np.random.default_rng(2)
    lim = np.nanmax(np.abs(d["uvw"][..., :2]))
    d["uvw"][..., :2] = rng.uniform(-lim, lim, size=d["uvw"][..., :2].shape)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(128, d["cell"], chan_mode="continuum")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    for algo in ("track", "shift", "window"):
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp, algorithm=ALGOS[algo])
        _check(g, s, g_ref, s_ref, TOL["f64"])


def test_empty_and_all_masked(sg, oracle):
    from cngi_prototype_b200 import synth
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    d = synth.make_vis_set(4, 6, 3, 2, 1e9, 1.1e9, 300.0, 90.0, seed=9)
    gp = synth.grid_parms_for(64, d["cell"], chan_mode="cube")
    # zero time steps
    g, s = sg._standard_grid_numpy_wrap(d["vis"][:0], d["uvw"][:0], d["weight"][:0], d["freq_chan"], cgk, gp)
    assert g.shape == (3, 2, 64, 64) and not g.any() and not s.any()
    # everything flagged / zero weight / NaN uvw
    for algo in ALGOS.values():
        g, s = sg._standard_grid_numpy_wrap(d["vis"] * np.nan, d["uvw"], d["weight"], d["freq_chan"], cgk, gp,
                                            algorithm=algo)
        assert not g.any() and not s.any()
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"] * 0, d["freq_chan"], cgk, gp,
                                            algorithm=algo)
        assert not g.any() and not s.any()
        g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"] * np.nan, d["weight"], d["freq_chan"], cgk, gp,
                                            algorithm=algo)
        assert not g.any() and not s.any()


def test_accumulate_into_device_grid_matches_single_call(sg, oracle):
    """Chunked accumulation into one device-resident grid == one call over all times (graph-level reduce)."""
    import torch
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(8, 48, 8, 2, 1e9, 1.1e9, 300.0, 90.0, seed=12)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(128, d["cell"], chan_mode="continuum")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")}
    grid = sw = None
    for t0 in range(0, 48, 13):
        sl = slice(t0, t0 + 13)
        grid, sw = sg.standard_grid(T["vis"][sl], T["uvw"][sl], T["weight"][sl], T["freq_chan"], cgk, gp, False, True,
                                    grid=grid, sum_weight=sw)
    _check(grid.cpu().numpy(), sw.cpu().numpy(), g_ref, s_ref, TOL["f64"])


def test_full_size_config1_properties(sg):
    """BASELINE config 1 at full size (351 bl x 1000 t x 64 ch x 2 pol, 1024^2, S=7, fp64, cube):
    size-independent properties instead of an oracle run:
      * PSF mode: sum over each plane == sum_weight of that plane (both are sum_samples w * sum_taps conv);
      * track, shift, window and naive kernels agree to 1e-12 with identical support masks;
      * linearity: grid(2*vis) == 2*grid(vis) exactly (power-of-two scaling commutes with every rounding)."""
    import torch
    from cngi_prototype_b200 import synth
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    d = synth.config_c1()
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")}
    gp = synth.grid_parms_for(1024, d["cell"], chan_mode="cube")
    gpp = dict(gp, do_psf=True, complex_grid=False)
    g, s = sg._standard_grid_psf_numpy_wrap(T["uvw"], T["weight"], T["freq_chan"], cgk, gpp)
    tot = g.sum(dim=(2, 3))
    assert float(((tot - s).abs() / s.abs()).max()) < 1e-11
    gn, sn = sg._standard_grid_numpy_wrap(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, algorithm=1)
    for algo in ("track", "shift", "window"):
        gt, st = sg._standard_grid_numpy_wrap(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp,
                                              algorithm=ALGOS[algo])
        assert bool(((gt != 0) == (gn != 0)).all()), algo
        assert float((gt - gn).abs().max() / gn.abs().max()) < 1e-12, algo
        assert float(((st - sn).abs() / sn.abs()).max()) < 1e-12, algo
        del gt
    del gn
    g2, _ = sg._standard_grid_numpy_wrap(T["vis"] * 2, T["uvw"], T["weight"], T["freq_chan"], cgk, gp, algorithm=4,
                                         time_segment=1000)
    g1, _ = sg._standard_grid_numpy_wrap(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, algorithm=4,
                                         time_segment=1000)
    # same work decomposition, but atomics may land in a different order: equal within rounding, not bitwise
    assert float((g2 - 2 * g1).abs().max() / g1.abs().max()) < 1e-12


@pytest.mark.parametrize("pattern", [0, 1, 2])
def test_reduction_microbench_counts(pattern):
    """cngi_b200_microbench_red (the atomic-roofline probe of bench.py) issues exactly blocks*256*per_thread reductions."""
    import torch
    from cngi_prototype_b200 import _lib
    from cngi_prototype_b200._devutil import ptr, stream
    buf = torch.zeros(1 << 16, dtype=torch.complex64, device="cuda")
    _lib.check(_lib.lib().cngi_b200_microbench_red(ptr(buf), buf.numel(), pattern, 7, 5, stream()), "microbench_red")
    total = buf.sum().cpu().item()
    assert total.real == 7 * 256 * 5 and total.imag == -7 * 256 * 5


@pytest.mark.parametrize("algo", [4, 2, 1])
def test_infinite_and_huge_uvw_rows_are_skipped_not_wrapped(oracle, algo):
    """+-inf or huge finite uvw saturate the int conversion of the cell index; `centre +- half` must not wrap
    around and pass the bounds test (window, track and naive kernels, imaging-weight grid / degrid, degrid predict).  Such
    rows are skipped like any other off-grid sample; the reference (numba) has undefined behaviour there, so the expected
    result is the grid of the same data with those rows made NaN."""
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg, _imaging_weight as iw, _standard_degrid as sd
    d = synth.make_vis_set(7, 12, 4, 2, 1e9, 1.1e9, 300.0, 120.0, seed=3, dtype="f64")
    bad = d["uvw"].copy()
    bad[1, 2, 0], bad[3, 4, 1], bad[5, 6, 0], bad[7, 8, 1] = np.inf, -np.inf, 1e300, -1e300
    ref_uvw = d["uvw"].copy()
    for (t, b) in ((1, 2), (3, 4), (5, 6), (7, 8)):
        ref_uvw[t, b, :2] = np.nan
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(96, d["cell"], chan_mode="cube")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], ref_uvw, d["weight"], d["freq_chan"], cgk, gp)
    g, s = sg._standard_grid_numpy_wrap(d["vis"], bad, d["weight"], d["freq_chan"], cgk, gp, algorithm=algo)
    assert same_support(g, g_ref) and rel_err(g, g_ref) <= 1e-12 and rel_err(s, s_ref) <= 1e-12
    if algo == 4:
        gpw = dict(gp, support=1, oversampling=0, do_psf=True, complex_grid=False, do_imaging_weight=True)
        rho_ref, sw_ref = oracle._standard_grid_psf_numpy_wrap(ref_uvw, d["weight"], d["freq_chan"], np.ones(1), gpw)
        rho, sw = iw.imaging_weight_grid(bad, d["weight"], d["freq_chan"], gpw)
        assert rel_err(rho, rho_ref) <= 1e-12 and rel_err(sw, sw_ref) <= 1e-12
        bf = oracle._calculate_briggs_parms(rho_ref, sw_ref, {"weighting": "briggs", "robust": 0.5})
        w_ref = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_ref, (0, 1), (2, 3)), ref_uvw, d["weight"], bf,
                                                                 d["freq_chan"], gpw)
        w_gpu = iw._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_ref, (0, 1), (2, 3)), bad, d["weight"], bf,
                                                              d["freq_chan"], gpw)
        assert np.array_equal(np.nan_to_num(w_gpu, nan=-1), np.nan_to_num(w_ref, nan=-1))
        v_ref = oracle._standard_degrid_numpy_wrap(g_ref, ref_uvw, d["freq_chan"], cgk, gp)
        v = sd._standard_degrid_numpy_wrap(g_ref, bad, d["freq_chan"], cgk, gp)
        assert np.array_equal(v == 0, v_ref == 0) and rel_err(v, v_ref) <= 1e-12
