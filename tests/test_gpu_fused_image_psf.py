"""Fused image + psf gridding pass (cngi_b200_standard_grid_image_psf) against the reference
fixtures and the oracle: the image half must equal _standard_grid_numpy_wrap's output and the psf half
_standard_grid_psf_numpy_wrap's (/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:123,180), masks
bit-exact, values 1e-12 (fp64) / 1e-5 (fp32)."""
import numpy as np
import pytest

from _util import load_golden, rel_err, same_support

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-12, "f32": 1e-5}


def _check(got, ref, tol):
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        assert rel_err(g, r) <= tol, rel_err(g, r)
    assert same_support(got[0], ref[0]) and same_support(got[2], ref[2])


@pytest.mark.parametrize("name", ["std_single_sample", "std_halfway_edges", "std_cube_sq", "std_cube_odd",
                                  "std_continuum_sq", "std_continuum_odd"])
def test_golden_fp64(oracle, name):
    from cngi_prototype_b200 import _standard_grid as sg
    d, gp = load_golden(name)
    got = sg.standard_grid_image_psf(d["vis"], d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"], gp, force_fused=True)
    if "psf_grid" in d:
        psf, psw = d["psf_grid"], d["psf_sum_weight"]
    else:
        psf, psw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"],
                                                        dict(gp, do_psf=True, complex_grid=False))
    _check(got, (d["grid"], d["sum_weight"], psf, psw), TOL["f64"])


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("mode", ["cube", "continuum"])
@pytest.mark.parametrize("n_pol", [1, 2])
def test_vs_oracle_with_flags(oracle, prec, mode, n_pol):
    """Flagged vis with a good weight goes to the psf only; a NaN / zero weight goes to neither."""
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg
    d = synth.make_vis_set(12, 40, 9, n_pol, 1.0e9, 1.1e9, 300.0, 60.0, seed=70 + n_pol)
    rng = np.random.default_rng(5)
    flag = rng.random(d["vis"].shape) < 0.1
    gp = synth.grid_parms_for(192, d["cell"], chan_mode=mode)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    vis, w = d["vis"], d["weight"]
    if prec == "f32":
        vis, w = vis.astype(np.complex64), w.astype(np.float32)
    vis_flagged = np.where(flag, np.nan, vis.astype(np.complex128))
    ref_g, ref_s = oracle._standard_grid_numpy_wrap(vis_flagged, d["uvw"], w.astype(np.float64), d["freq_chan"], cgk, gp)
    ref_p, ref_ps = oracle._standard_grid_psf_numpy_wrap(d["uvw"], w.astype(np.float64), d["freq_chan"], cgk,
                                                         dict(gp, do_psf=True, complex_grid=False))
    dev = torch.device("cuda")
    got = sg.standard_grid_image_psf(torch.as_tensor(vis, device=dev), torch.as_tensor(d["uvw"], device=dev),
                                     torch.as_tensor(w, device=dev), d["freq_chan"], cgk, gp,
                                     flag=torch.as_tensor(flag, device=dev), force_fused=True)
    assert all(t.is_cuda for t in got)
    _check([t.cpu().numpy() for t in got], (ref_g, ref_s, ref_p, ref_ps), TOL[prec])
    # accumulate-into semantics: a second call doubles everything
    got2 = sg.standard_grid_image_psf(torch.as_tensor(vis, device=dev), torch.as_tensor(d["uvw"], device=dev),
                                      torch.as_tensor(w, device=dev), d["freq_chan"], cgk, gp,
                                      flag=torch.as_tensor(flag, device=dev), grid=got[0], sum_weight=got[1],
                                      psf_grid=got[2], psf_sum_weight=got[3])
    assert rel_err(got2[2].cpu().numpy(), 2 * ref_p) <= 2 * TOL[prec]


def test_other_support_takes_two_passes(oracle):
    from cngi_prototype_b200 import _standard_grid as sg
    d, gp = load_golden("std_cube_s5_1pol")
    got = sg.standard_grid_image_psf(d["vis"], d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"], gp)
    psf, psw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], d["cgk_1D"],
                                                    dict(gp, do_psf=True, complex_grid=False))
    _check(got, (d["grid"], d["sum_weight"], psf, psw), TOL["f64"])


def test_synthesis_imaging_chunk_fused_equals_unfused():
    import torch
    from cngi_prototype_b200 import synth, synthesis_chunk
    d = synth.make_vis_set(10, 30, 6, 2, 1.0e9, 1.1e9, 300.0, 60.0, seed=91)
    dev = torch.device("cuda")
    t = {k: torch.as_tensor(d[k], device=dev) for k in ("vis", "uvw", "weight", "freq_chan")}
    flag = torch.zeros(d["vis"].shape, dtype=torch.uint8, device=dev)
    gp = synth.grid_parms_for(160, d["cell"], chan_mode="cube")
    gp["image_size"] = np.array([128, 128])
    iw = dict(weighting="briggs", robust=0.5)
    a = synthesis_chunk.synthesis_imaging_chunk(t["vis"], t["uvw"], t["weight"], flag, t["freq_chan"], gp, iw, fused=True)
    b = synthesis_chunk.synthesis_imaging_chunk(t["vis"], t["uvw"], t["weight"], flag, t["freq_chan"], gp, iw, fused=False)
    for x, y in zip(a, b):
        assert rel_err(x.cpu().numpy(), y.cpu().numpy()) < 1e-12
