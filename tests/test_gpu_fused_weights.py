"""A4 folded into A1 (cngi_b200_standard_grid_weighted): natural weights in, imaging weights formed inside the gridder.

Checked against (a) the two-kernel path it replaces (cngi_b200_imaging_weight_degrid + cngi_b200_standard_grid): the
imaging weights it can write out are bit-identical, the grid agrees to the last bits (atomic order only) with identical
masks; and (b) the CPU oracle chain (_standard_grid.py:466-518 then :242-371)."""
import numpy as np
import pytest

from _util import rel_err, same_support

pytestmark = pytest.mark.gpu


def _chain(torch, d, n_uv, n_uv_iw, mode, prec, weighting, flag=None, time_chunks=1, want_out=True):
    from cngi_prototype_b200 import synth, _imaging_weight as iw, _standard_grid as sg
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n_uv, d["cell"], chan_mode=mode)
    gpw = synth.grid_parms_for(n_uv_iw, d["cell"], chan_mode=mode, support=1, oversampling=0, do_psf=True,
                               complex_grid=False, do_imaging_weight=True)
    vis, uvw, w, freq = (torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan"))
    fl = None if flag is None else torch.as_tensor(flag).cuda()
    rho, sw = iw.imaging_weight_grid(uvw, w, freq, gpw)
    bf = iw.calculate_briggs_parms(rho, sw, {"weighting": weighting, "robust": 0.5})
    iw_two = iw._standard_imaging_weight_degrid_numpy_wrap(rho, uvw, w, bf, freq, gpw, kernel_side_layout=True)
    g_two, s_two = sg.standard_grid(vis, uvw, iw_two, freq, cgk, gp, False, True, flag=fl)
    out = torch.full_like(w, -7.0) if want_out else None
    g = s = None
    n_t = w.shape[0]
    step = -(-n_t // time_chunks)
    for t0 in range(0, n_t, step):
        sl = slice(t0, min(n_t, t0 + step))
        src = dict(density=rho, briggs_factors=bf, grid_parms=gpw, out=None if out is None else out[sl])
        g, s = sg.standard_grid(vis[sl], uvw[sl], w[sl], freq, cgk, gp, False, True, flag=None if fl is None else fl[sl],
                                grid=g, sum_weight=s, imaging_weight_from=src)
    return dict(g=g, s=s, out=out, g_two=g_two, s_two=s_two, iw_two=iw_two, rho=rho, sw=sw, bf=bf, gp=gp, gpw=gpw, cgk=cgk)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("mode", ["cube", "continuum"])
@pytest.mark.parametrize("n_pol,weighting,n_uv_iw,chunks", [(2, "briggs", 160, 1), (2, "uniform", 128, 3), (1, "briggs", 160, 2)])
def test_fused_equals_two_kernel_path_and_oracle(oracle, prec, mode, n_pol, weighting, n_uv_iw, chunks):
    import torch
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(9, 37, 11, n_pol, 1e9, 1.12e9, 300.0, 90.0, seed=5 + n_pol, dtype=prec)
    # a density grid with its own size (make_imaging_weight grids on the UNPADDED image size): the same uv range mapped
    # onto fewer cells, so the two geometries differ (iw_own_scale) unless n_uv_iw == 160
    r = _chain(torch, d, 160, n_uv_iw, mode, prec, weighting, time_chunks=chunks)
    tol = 1e-12 if prec == "f64" else 2e-6
    g, g_two = r["g"].cpu().numpy(), r["g_two"].cpu().numpy()
    assert same_support(g, g_two) and rel_err(g, g_two) <= tol
    assert rel_err(r["s"].cpu().numpy(), r["s_two"].cpu().numpy()) <= 1e-12
    # the imaging weights it writes are the degrid kernel's, bit for bit (same operations in the same order)
    a, b = r["out"].cpu().numpy(), r["iw_two"].cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    # oracle chain on the same density (the density itself is compared in test_gpu_imaging_weight.py)
    rho_h, bf_h = r["rho"].cpu().numpy(), r["bf"].cpu().numpy()
    iw_ref = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho_h, (0, 1), (2, 3)), d["uvw"], d["weight"], bf_h,
                                                               d["freq_chan"], r["gpw"])
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw_ref, d["freq_chan"], r["cgk"], r["gp"])
    assert same_support(g, g_ref) and rel_err(g, g_ref) <= (1e-12 if prec == "f64" else 1e-5)
    assert rel_err(r["s"].cpu().numpy(), s_ref) <= (1e-12 if prec == "f64" else 1e-6)


def test_fused_with_flags_no_output_and_bad_samples(oracle):
    """FLAG fused as well; imaging_weight NULL (nothing written); NaN / zero natural weights, NaN uvw rows, density cells
    that are exactly zero (weights of a whole baseline zeroed -> rho == 0 there -> undivided weight)."""
    import torch
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(8, 30, 8, 2, 1e9, 1.1e9, 300.0, 120.0, seed=17, dtype="f64")
    d["weight"][:, 3] = 0.0                      # one baseline carries no weight: its cells' density stays 0
    d["weight"][5, 4, 2, 1] = np.nan
    flag = np.random.default_rng(2).random(d["vis"].shape) < 0.07
    r = _chain(torch, d, 128, 128, "continuum", "f64", "briggs", flag=flag, want_out=False)
    g, g_two = r["g"].cpu().numpy(), r["g_two"].cpu().numpy()
    assert same_support(g, g_two) and rel_err(g, g_two) <= 1e-12
    assert rel_err(r["s"].cpu().numpy(), r["s_two"].cpu().numpy()) <= 1e-12


def test_fused_rejects_what_it_cannot_do():
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg, _lib
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    d = synth.make_vis_set(5, 6, 4, 2, 1e9, 1.1e9, 300.0, 120.0, seed=1)
    gp = synth.grid_parms_for(64, d["cell"], chan_mode="cube", support=5, oversampling=50)
    vis, uvw, w, freq = (torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan"))
    src = dict(density=torch.zeros((4, 2, 64, 64), dtype=torch.float64, device="cuda"),
               briggs_factors=torch.ones((2, 4, 2), dtype=torch.float64, device="cuda"), grid_parms=gp)
    with pytest.raises(_lib.CngiError):     # support 5: the fused pass exists for support 7 (make_image.py:106-107)
        sg.standard_grid(vis, uvw, w, freq, _create_prolate_spheroidal_kernel_1D(50, 5), gp, False, True, imaging_weight_from=src)
