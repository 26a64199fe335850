"""GPU parity of the imaging-weight path (A2 density grid, A3 Briggs factors, A4 weight degrid)."""
import numpy as np
import pytest

from _util import load_golden, rel_err, same_support

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def iw():
    import torch
    assert torch.cuda.is_available()
    from cngi_prototype_b200 import _imaging_weight
    return _imaging_weight


@pytest.mark.parametrize("name", ["iw_cube_2pol", "iw_continuum_2pol", "iw_cube_1pol", "iw_continuum_1pol"])
def test_golden(iw, name):
    from cngi_prototype_b200 import _standard_grid
    d, gp = load_golden(name)
    rho, sw = _standard_grid._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp)
    assert same_support(rho, d["density"])
    assert rel_err(rho, d["density"]) <= 1e-12 and rel_err(sw, d["sum_weight"]) <= 1e-12
    bf = iw.calculate_briggs_parms(d["density"], d["sum_weight"], dict(weighting="briggs", robust=0.5))
    assert rel_err(bf, d["briggs_factors"]) <= 1e-12
    # API-side layout (u, v, chan, pol) as the reference passes it, and the kernel-side layout without a transpose
    out = iw._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(d["density"], (0, 1), (2, 3)), d["uvw"], d["weight"],
                                                        d["briggs_factors"], d["freq_chan"], gp)
    assert np.array_equal(out, d["imaging_weight"], equal_nan=True)   # same IEEE ops -> bit exact
    out2 = iw._standard_imaging_weight_degrid_numpy_wrap(d["density"], d["uvw"], d["weight"], d["briggs_factors"],
                                                         d["freq_chan"], gp, kernel_side_layout=True)
    assert np.array_equal(out2, d["imaging_weight"], equal_nan=True)
    bfu = iw.calculate_briggs_parms(d["density"], d["sum_weight"], dict(weighting="uniform"))
    assert np.all(bfu[0] == 1) and np.all(bfu[1] == 0)


@pytest.mark.parametrize("mode", ["cube", "continuum"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_alma_like_vs_oracle(iw, oracle, mode, prec):
    """Config-2 geometry at reduced size, full Briggs chain on the device (torch tensors in, torch out)."""
    import torch
    from cngi_prototype_b200 import synth
    d = synth.config_c2(n_time=40, n_chan=12, dtype=prec)
    gp = synth.grid_parms_for(512, d["cell"], chan_mode=mode, support=1, oversampling=0, do_psf=True,
                              complex_grid=False, do_imaging_weight=True)
    rho_ref, sw_ref = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp)
    parms = dict(weighting="briggs", robust=0.5)
    bf_ref = oracle._calculate_briggs_parms(rho_ref, sw_ref, parms)
    out_ref = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_ref, (0, 1), (2, 3)), d["uvw"],
                                                                d["weight"], bf_ref, d["freq_chan"], gp)
    T = {k: torch.as_tensor(d[k]).cuda() for k in ("uvw", "weight", "freq_chan")}
    rho, sw = iw.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gp)
    assert same_support(rho.cpu().numpy(), rho_ref)
    assert rel_err(rho.cpu().numpy(), rho_ref) <= 1e-12 and rel_err(sw.cpu().numpy(), sw_ref) <= 1e-12
    bf = iw.calculate_briggs_parms(rho, sw, parms)
    assert rel_err(bf.cpu().numpy(), bf_ref) <= 1e-12
    out = iw._standard_imaging_weight_degrid_numpy_wrap(rho, T["uvw"], T["weight"], bf, T["freq_chan"], gp,
                                                        kernel_side_layout=True)
    tol = 1e-12 if prec == "f64" else 1e-6
    o = out.cpu().numpy().astype(np.float64)
    m = np.isfinite(out_ref)
    assert np.array_equal(np.isnan(o), np.isnan(out_ref))
    assert np.array_equal(o == 0, out_ref == 0)
    assert np.max(np.abs(o[m] - out_ref[m])) <= tol * np.max(np.abs(out_ref[m]))


@pytest.mark.parametrize("mode,n_chan,n_pol", [("continuum", 16, 2), ("continuum", 32, 1), ("cube", 6, 2), ("continuum", 12, 2)])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fast_kernels_vs_oracle_and_first_generation(iw, oracle, monkeypatch, mode, n_chan, n_pol, prec):
    """iw_grid_fast_kernel / iw_degrid_fast_kernel (whole channel groups, identity pol map: the product path) against the
    oracle and against the general kernels they replace (CNGI_IW_GRID_OLD / CNGI_IW_DEGRID_MLP select those; n_chan = 12
    is a ragged channel group, which the general density kernel serves anyway).  Includes NaN / zero weights, NaN uvw and
    the first_pol_only + stride-0 pol planes form the pipeline uses."""
    import torch
    from cngi_prototype_b200 import synth
    d = synth.config_c2(n_time=37, n_chan=n_chan, dtype=prec)
    w = d["weight"][..., :n_pol].copy()
    w[3, 5, 2, 0] = np.nan
    w[4, 7, 1, :] = 0.0
    uvw = d["uvw"].copy()
    uvw[6, 11, 0] = np.nan
    uvw[7, 3, :] *= 40.0   # off the grid
    gp = synth.grid_parms_for(384, d["cell"], chan_mode=mode, support=1, oversampling=0, do_psf=True,
                              complex_grid=False, do_imaging_weight=True)
    rho_ref, sw_ref = oracle._standard_grid_psf_numpy_wrap(uvw, w, d["freq_chan"], np.ones(1), gp)
    parms = dict(weighting="briggs", robust=0.5)
    bf_ref = oracle._calculate_briggs_parms(rho_ref, sw_ref, parms)
    out_ref = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_ref, (0, 1), (2, 3)), uvw, w, bf_ref,
                                                                d["freq_chan"], gp)
    T = {"uvw": torch.as_tensor(uvw).cuda(), "weight": torch.as_tensor(w).cuda(), "freq_chan": torch.as_tensor(d["freq_chan"]).cuda()}
    got = {}
    for gen in ("1", "0"):
        monkeypatch.setenv("CNGI_IW_GRID_OLD", gen)
        monkeypatch.setenv("CNGI_IW_DEGRID_MLP", gen)
        rho, sw = iw.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gp)
        assert same_support(rho.cpu().numpy(), rho_ref)
        assert rel_err(rho.cpu().numpy(), rho_ref) <= 1e-12 and rel_err(sw.cpu().numpy(), sw_ref) <= 1e-12
        if n_pol == 2:   # plane 0 only + stride-0 views, as ContinuumPipeline does
            rho1, sw1 = iw.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gp, first_pol_only=True)
            assert rel_err(rho1[:, :1].cpu().numpy(), rho_ref[:, :1]) <= 1e-12 and float(rho1[:, 1:].abs().max()) == 0.0
        # the degrid generations on ONE density (the oracle's): bit-identical outputs
        rho_in = torch.as_tensor(rho_ref).cuda()
        bf_in = torch.as_tensor(bf_ref).cuda()
        out = iw._standard_imaging_weight_degrid_numpy_wrap(rho_in, T["uvw"], T["weight"], bf_in, T["freq_chan"], gp,
                                                            kernel_side_layout=True)
        got[gen] = [out.cpu().numpy()]
        if n_pol == 2:
            shared = iw._standard_imaging_weight_degrid_numpy_wrap(rho_in[:, :1].expand(-1, 2, -1, -1), T["uvw"], T["weight"],
                                                                   bf_in[:, :, :1].expand(-1, -1, 2), T["freq_chan"], gp,
                                                                   kernel_side_layout=True)
            got[gen].append(shared.cpu().numpy())
        api_side = iw._standard_imaging_weight_degrid_numpy_wrap(rho_in.permute(2, 3, 0, 1), T["uvw"], T["weight"], bf_in,
                                                                 T["freq_chan"], gp)
        got[gen].append(api_side.cpu().numpy())
        o = got[gen][0].astype(np.float64)
        m = np.isfinite(out_ref)
        assert np.array_equal(np.isnan(o), np.isnan(out_ref)) and np.array_equal(o == 0, out_ref == 0)
        assert np.max(np.abs(o[m] - out_ref[m])) <= (1e-12 if prec == "f64" else 1e-6) * np.max(np.abs(out_ref[m]))
    for a, b in zip(got["1"], got["0"]):
        assert np.array_equal(a, b, equal_nan=True)
