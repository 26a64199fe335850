"""BASELINE configs 2, 3, 4 and 5 at FULL size on the GPU, checked through size-independent properties (linearity, chunk
invariance, adjointness, plane sums, kernel-variant agreement).  The value-for-value comparison with the CPU oracle at
the same sizes is tests/test_gpu_full_size_parity.py.

  config 2: Briggs(0.5) imaging weights + standard gridding, ALMA-like 903 bl x 500 t x 128 ch x 2 pol = 115.6 M samples,
            4096^2, fp32, continuum -- bench.py's workload.
  config 4: degridding predict, 27 antennas (351 bl) x 1000 t x 64 ch x 2 pol = 44.9 M samples, 4096^2, S=7, fp64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALGO_NAIVE, ALGO_TRACK, ALGO_WINDOW = 1, 2, 4


def test_full_size_config2_weights_and_gridding_properties():
    import torch
    from cngi_prototype_b200 import synth, _imaging_weight as iw, _standard_grid as sg
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    d = synth.config_c2()
    assert d["weight"].size == 500 * 903 * 128 * 2 and d["vis"].dtype == np.complex64
    vis, uvw, w, freq = (torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan"))
    n = 4096
    gpw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                               complex_grid=False, do_imaging_weight=True)
    # A2: every on-grid sample adds its pol-averaged weight to its cell and to the conjugate cell, and twice to sum_weight
    rho, sw = iw.imaging_weight_grid(uvw, w, freq, gpw)
    assert rho.dtype == torch.float64 and tuple(rho.shape) == (1, 2, n, n)
    assert float(((rho.sum(dim=(2, 3)) - sw).abs() / sw).max()) < 1e-12
    assert torch.equal(rho[:, 0], rho[:, 1]) and float(rho.min()) >= 0
    # A3: Briggs factors against their definition evaluated with torch on the same density
    bf = iw.calculate_briggs_parms(rho, sw, {"weighting": "briggs", "robust": 0.5})
    f0 = (5 * 10 ** -0.5) ** 2 / ((rho * rho).sum(dim=(2, 3)) / sw)
    assert float(((bf[0] - f0).abs() / f0).max()) < 1e-12 and bool((bf[1] == 1).all())
    # A4: imaging weight = pol-averaged natural weight / (f0 rho + 1) <= natural; 0 where the uv point is NaN;
    #     per-sample, so a time slice on its own is bit-identical to the slice of the full result
    iwt = iw._standard_imaging_weight_degrid_numpy_wrap(rho, uvw, w, bf, freq, gpw, kernel_side_layout=True)
    avg = ((w[..., 0] + w[..., 1]) / 2)[..., None].expand_as(w)
    fin = torch.isfinite(iwt) & torch.isfinite(avg)
    assert bool((iwt[fin] >= 0).all()) and bool((iwt[fin] <= avg[fin] * (1 + 1e-6)).all())
    assert float(fin.float().mean()) > 0.99
    bad_uv = torch.isnan(uvw[..., 0]) | torch.isnan(uvw[..., 1])
    assert bool(bad_uv.any()) and bool((iwt[bad_uv] == 0).all())
    part = iw._standard_imaging_weight_degrid_numpy_wrap(rho, uvw[100:150], w[100:150], bf, freq, gpw, kernel_side_layout=True)
    assert torch.equal(torch.nan_to_num(part, nan=-1.0), torch.nan_to_num(iwt[100:150], nan=-1.0))
    del avg, fin, part
    # A1 fp32 continuum with those weights: product (window), track and naive kernels agree, masks identical
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    g_w, s_w = sg._standard_grid_numpy_wrap(vis, uvw, iwt, freq, cgk, gp, algorithm=ALGO_WINDOW)
    assert g_w.dtype == torch.complex64 and tuple(g_w.shape) == (1, 2, n, n)
    scale = float(g_w.abs().max())
    for algo in (ALGO_TRACK, ALGO_NAIVE):
        g_o, s_o = sg._standard_grid_numpy_wrap(vis, uvw, iwt, freq, cgk, gp, algorithm=algo)
        assert bool(((g_o != 0) == (g_w != 0)).all()), algo
        assert float((g_o - g_w).abs().max()) / scale < 1e-5, algo
        assert float(((s_o - s_w).abs() / s_w.abs()).max()) < 1e-6, algo
        del g_o
    # time-chunked accumulation into one device grid == one call
    g_c = s_c = None
    for t0 in range(0, 500, 125):
        g_c, s_c = sg.standard_grid(vis[t0:t0 + 125], uvw[t0:t0 + 125], iwt[t0:t0 + 125], freq, cgk, gp, False, True,
                                    grid=g_c, sum_weight=s_c)
    assert bool(((g_c != 0) == (g_w != 0)).all()) and float((g_c - g_w).abs().max()) / scale < 1e-5
    assert float(((s_c - s_w).abs() / s_w.abs()).max()) < 1e-6
    # linearity: power-of-two scaling commutes with every fp32 rounding (only the atomic order differs)
    g_2, _ = sg._standard_grid_numpy_wrap(vis * 2, uvw, iwt, freq, cgk, gp, algorithm=ALGO_WINDOW)
    assert float((g_2 - 2 * g_w).abs().max()) / scale < 1e-5
    del g_2, g_c
    # psf mode: the plane sum equals sum_weight (both are sum_samples w * sum_taps conv)
    g_p, s_p = sg._standard_grid_psf_numpy_wrap(uvw, iwt, freq, cgk, dict(gp, do_psf=True, complex_grid=False))
    assert g_p.dtype == torch.float32
    assert float(((g_p.double().sum(dim=(2, 3)) - s_p).abs() / s_p).max()) < 1e-5


def test_full_size_config4_degrid_is_the_adjoint_of_the_gridder():
    """<grid(x), y> == <x, degrid(y)> at 4096^2 for 44.9 M fp64 samples (no reference implementation exists for the
    predict, SURVEY section 8a A7: adjointness with the parity-checked gridder is the full-size check), window and
    gather degrid kernels agree, samples the gridder would skip come back as exact zeros."""
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg, _standard_degrid as sd
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    d = synth.config_c4()
    d["uvw"].reshape(-1, 3)[::997, 0] = np.nan                                   # a few rows without a uv point
    uvw, freq = torch.as_tensor(d["uvw"]).cuda(), torch.as_tensor(d["freq_chan"]).cuda()
    n_t, n_b, n_c, n_p = 1000, 351, 64, 2
    n = 4096
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    gen = torch.Generator(device="cuda").manual_seed(7)
    y = torch.randn((1, n_p, n, n, 2), dtype=torch.float64, device="cuda", generator=gen)
    y = torch.view_as_complex(y)
    v = sd._standard_degrid_numpy_wrap(y, uvw, freq, cgk, gp)                      # auto = window kernel
    assert tuple(v.shape) == (n_t, n_b, n_c, n_p) and v.dtype == torch.complex128
    v1 = sd._standard_degrid_numpy_wrap(y, uvw, freq, cgk, gp, algorithm=1)        # gather kernel
    assert float((v - v1).abs().max() / v.abs().max()) < 1e-12
    del v1
    bad_uv = torch.isnan(uvw[..., 0]) | torch.isnan(uvw[..., 1])
    assert bool(bad_uv.any()) and bool((v[bad_uv] == 0).all())
    x = torch.view_as_complex(torch.randn((n_t, n_b, n_c, n_p, 2), dtype=torch.float64, device="cuda", generator=gen))
    ones = torch.ones((n_t, n_b, n_c, n_p), dtype=torch.float64, device="cuda")
    g, _ = sg._standard_grid_numpy_wrap(x, uvw, ones, freq, cgk, gp)
    lhs = torch.vdot(y.reshape(-1), g.reshape(-1))
    rhs = torch.vdot(v.reshape(-1), x.reshape(-1))
    assert float((lhs - rhs).abs() / lhs.abs()) < 1e-11


def test_config3_size_aperture_gridders_properties():
    """BASELINE config 3 at the size DESIGN.md's rows quote (903 bl x 200 t x 64 ch x 2 pol = 23.1 M samples, 7-pointing
    mosaic, CF 160^2 with supports 9..15, 2048^2, continuum):
      * psf mode: Re(sum of a grid plane) == sum_weight of that plane (grid += conv * w, sum_weight += w * Re(sum conv),
        _aperture_grid.py:496-511);
      * fp32 agrees with fp64 to 1e-5 with identical support masks (index math is fp64 in both);
      * time-chunked accumulation into one device grid == one call;
      * the weight gridder touches at most max_support^2 cells per plane, centred on the grid (:276-287), and its
        plane sum also equals sum_weight."""
    import torch
    from cngi_prototype_b200 import synth, _aperture_grid as ap
    d32 = synth.config_c2(n_time=200, n_chan=64, dtype="f32")
    gcf = synth.make_mosaic_gcf(d32["n_baseline"], 64, 2, n_field=7)
    field = synth.mosaic_field_column(d32["uvw"].shape[0], d32["n_baseline"], gcf["field_id"])
    n = 2048
    gp = synth.grid_parms_for(n, d32["cell"] * 1.1, chan_mode="continuum")
    gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
    uvw, freq, fld = (torch.as_tensor(x).cuda() for x in (d32["uvw"], d32["freq_chan"], field))
    G = {k: torch.as_tensor(v).cuda() for k, v in gcf.items()}
    maps = (G["cf_baseline_map"], G["cf_chan_map"], G["cf_pol_map"])
    w32, v32 = torch.as_tensor(d32["weight"]).cuda(), torch.as_tensor(d32["vis"]).cuda()
    w64, v64 = w32.double(), v32.to(torch.complex128)

    def image(v, w, **kw):
        return ap._aperture_grid_numpy_wrap(v, uvw, w, fld, *maps, G["conv_kernel"], gcf["weight_support"],
                                            G["phase_gradient"], freq, gp, **kw)

    def psf(w):
        return ap._aperture_psf_grid_numpy_wrap(uvw, w, fld, *maps, G["conv_kernel"], gcf["weight_support"],
                                                G["phase_gradient"], freq, gp)

    p64, ps64 = psf(w64)
    assert p64.dtype == torch.complex128 and tuple(p64.shape) == (1, 2, n, n)
    assert float(((p64.sum(dim=(2, 3)).real - ps64).abs() / ps64).max()) < 1e-11
    p32, ps32 = psf(w32)
    assert float(((p32.to(torch.complex128).sum(dim=(2, 3)).real - ps32).abs() / ps32).max()) < 1e-5
    del p64, p32
    g64, s64 = image(v64, w64)
    g32, s32 = image(v32, w32)
    scale = float(g64.abs().max())
    assert bool(((g32 != 0) == (g64 != 0)).all())
    assert float((g32.to(torch.complex128) - g64).abs().max()) / scale < 1e-5
    assert float(((s32 - s64).abs() / s64.abs()).max()) < 1e-6
    del g32
    gc = sc = None
    for t0 in range(0, 200, 50):
        sl = slice(t0, t0 + 50)
        gc, sc = ap._aperture_grid_numpy_wrap(v64[sl], uvw[sl], w64[sl], fld[sl], *maps, G["conv_kernel"],
                                              gcf["weight_support"], G["phase_gradient"], freq, gp, grid=gc, sum_weight=sc)
    assert bool(((gc != 0) == (g64 != 0)).all()) and float((gc - g64).abs().max()) / scale < 1e-12
    assert float(((sc - s64).abs() / s64.abs()).max()) < 1e-12
    del gc, g64
    gw, sw = ap._aperture_weight_grid_numpy_wrap(uvw, w64, fld, *maps, G["weight_conv_kernel"], gcf["weight_support"],
                                                 G["phase_gradient"], freq, gp)
    ms = int(np.max(gcf["weight_support"]))
    nz = (gw != 0).nonzero()
    assert 0 < nz.shape[0] <= 2 * ms * ms
    assert int((nz[:, 2] - n // 2).abs().max()) <= ms // 2 and int((nz[:, 3] - n // 2).abs().max()) <= ms // 2
    assert float(((gw.sum(dim=(2, 3)).real - sw).abs() / sw.abs()).max()) < 1e-11


def test_config5_grid_size_cube_chunked_through_one_buffer():
    """BASELINE config 5's grid (8192^2 image, padded to 9830 = 2 * 5 * 983 per side, fp32 cube) on a few channels:
      * the cube driver that walks `chan_chunk` planes at a time through one grid buffer gives, channel for channel, the
        image of that channel gridded on its own (chunk invariance -- what sharding the cube by channel relies on);
      * PSF: the centre pixel of every plane is 1 / (PS correcting image at the centre): the plane sum of the uv-grid
        equals sum_weight, so ifft -> / sum_weight is exactly 1 there (make_psf.py:117-130), and it is the plane's maximum."""
    import torch
    from cngi_prototype_b200 import synth, imaging
    from cngi_prototype_b200._gridding_convolutional_kernels import correcting_function_1D
    d = synth.config_c1(n_time=200, n_chan=6, dtype="f32")
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD
    ds = {"DATA": torch.as_tensor(d["vis"]).cuda(), "UVW": torch.as_tensor(d["uvw"]).cuda(),
          "WEIGHT": torch.as_tensor(d["weight"]).cuda(), "chan": d["freq_chan"]}
    gp = {"image_size": [8192, 8192], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.2, "chan_mode": "cube"}
    psf = imaging.make_psf(ds, gp, weight_key="WEIGHT", chan_chunk=4)
    P = psf["PSF"]
    assert tuple(P.shape) == (8192, 8192, 6, 2) and P.dtype == torch.float32
    cu, cv = correcting_function_1D(np.array([9830, 9830]), np.array([8192, 8192]))
    centre = 1.0 / (cu[4096] * cv[4096])
    assert float((P[4096, 4096] - centre).abs().max()) < 1e-5
    assert float(P.amax(dim=(0, 1)).sub(P[4096, 4096]).abs().max()) == 0.0
    img = imaging.make_image(ds, gp, weight_key="WEIGHT", chan_chunk=4)
    one = {k: (v[5:6] if k == "chan" else (v[:, :, 5:6] if getattr(v, "ndim", 0) == 4 else v)) for k, v in ds.items()}
    img5 = imaging.make_image(one, gp, weight_key="WEIGHT")
    a, b = img["IMAGE"][:, :, 5], img5["IMAGE"][:, :, 0]
    assert float((a - b).abs().max() / b.abs().max()) < 1e-5
    assert float(((img["SUM_WEIGHT"][5] - img5["SUM_WEIGHT"][0]).abs() / img5["SUM_WEIGHT"][0]).max()) < 1e-6
