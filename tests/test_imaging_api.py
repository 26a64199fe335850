"""API layer (cngi_prototype_b200/imaging.py): parameter handling on CPU, full chains against the oracle on the GPU."""
import numpy as np
import pytest

from _util import rel_err


def test_check_grid_parms_mirrors_reference_conventions():
    """_check_imaging_parms.py:22-41: arcsec -> rad, x negated, padded size = int(fft_padding * size), defaults."""
    from cngi_prototype_b200.imaging import _check_grid_parms, _check_imaging_weights_parms
    gp = {"image_size": [200, 400], "cell_size": [0.08, 0.08]}
    assert _check_grid_parms(gp)
    assert gp["chan_mode"] == "cube" and gp["fft_padding"] == 1.2
    assert list(gp["image_size_padded"]) == [240, 480]
    assert list(gp["image_center"]) == [100, 200]
    rad = 0.08 * np.pi / (3600 * 180)
    assert gp["cell_size"][0] == -rad and gp["cell_size"][1] == rad
    assert list(_check_grid_parms_padded([4096, 4096], 1.2)) == [4915, 4915]     # 5 * 983: odd
    assert not _check_grid_parms({"image_size": [10, 10], "cell_size": [1, 1], "fft_padding": 0.5})
    assert not _check_grid_parms({"image_size": [10, 10], "cell_size": [1, 1], "chan_mode": "mfs"})
    iw = {}
    assert _check_imaging_weights_parms(iw) and iw["weighting"] == "natural"
    iw = {"weighting": "briggs"}
    assert _check_imaging_weights_parms(iw) and iw["robust"] == 0.5
    assert not _check_imaging_weights_parms({"weighting": "briggs", "robust": 3})


def _check_grid_parms_padded(size, pad):
    from cngi_prototype_b200.imaging import _check_grid_parms
    gp = {"image_size": size, "cell_size": [1, 1], "fft_padding": pad}
    assert _check_grid_parms(gp)
    return gp["image_size_padded"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_weight_psf_image_chain_vs_oracle(oracle, mode):
    """make_imaging_weight(briggs) -> make_psf -> make_image on a dict dataset, padded (odd padded size), time-chunked,
    against the same chain through the oracle."""
    from cngi_prototype_b200 import synth, imaging
    d = synth.config_c1(n_time=36, n_chan=6)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD * 1.3    # padded grid covers the same uv range
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    gp = {"image_size": [150, 135], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.3, "chan_mode": mode}
    iwp = {"weighting": "briggs", "robust": 0.5}
    ds2 = imaging.make_imaging_weight(ds, iwp, gp, time_chunk=10)
    psf = imaging.make_psf(ds2, gp, time_chunk=7)
    img = imaging.make_image(ds2, gp, time_chunk=11)
    # oracle chain
    g = dict(gp)
    assert imaging._check_grid_parms(g)
    assert list(g["image_size_padded"]) == [195, 175]
    gw = dict(g, image_size_padded=g["image_size"], oversampling=0, support=1, do_psf=True, complex_grid=False,
              do_imaging_weight=True)
    rho, sw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gw)
    bf = oracle._calculate_briggs_parms(rho, sw, iwp)
    iw = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"], bf,
                                                           d["freq_chan"], gw)
    m = np.isfinite(iw)
    assert np.array_equal(np.isnan(ds2["IMAGING_WEIGHT"]), np.isnan(iw))
    assert np.max(np.abs(ds2["IMAGING_WEIGHT"][m] - iw[m])) <= 1e-12 * np.max(np.abs(iw[m]))
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gi = dict(g, oversampling=100, support=7, do_psf=False, complex_grid=True, do_imaging_weight=False)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(g["image_size_padded"]), g["image_size"])
    gg, ss = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw, d["freq_chan"], cgk, gi)
    ref_img = oracle.correct_image(oracle.grid_to_uncorrected_image(gg, g["image_size"]), ss, corr)
    gpp, sp = oracle._standard_grid_psf_numpy_wrap(d["uvw"], iw, d["freq_chan"], cgk, dict(gi, do_psf=True, complex_grid=False))
    ref_psf = oracle.correct_image(oracle.grid_to_uncorrected_image(gpp, g["image_size"]), sp, corr)
    assert img["IMAGE"].shape == ref_img.shape == (150, 135, 6 if mode == "cube" else 1, 2)
    assert rel_err(img["IMAGE"], ref_img) <= 1e-11 and rel_err(img["SUM_WEIGHT"], ss) <= 1e-12
    assert rel_err(psf["PSF"], ref_psf) <= 1e-11 and rel_err(psf["PSF_SUM_WEIGHT"], sp) <= 1e-12
    # the PSF peaks at the image centre with value ~1 after normalisation
    c = psf["PSF"][75, 67]
    assert np.all(np.abs(c - 1.0) < 1e-3)
    # natural weighting aliases WEIGHT, make_grid returns the API-side layout
    assert imaging.make_imaging_weight(ds, {"weighting": "natural"}, gp)["IMAGING_WEIGHT"] is ds["WEIGHT"]
    G = imaging.make_grid(ds2, gp)
    assert G["GRID"].shape == (195, 175, 6 if mode == "cube" else 1, 2)
    assert rel_err(np.moveaxis(G["GRID"], (2, 3), (0, 1)), gg) <= 1e-12


@pytest.mark.gpu
def test_channel_chunked_cube_equals_unchunked(oracle):
    """make_image / make_psf with chan_chunk (bounded-memory cube imaging) and distributed.cube_imaging (the per-rank
    driver of the channel-sharded cube, world size 1 here) give the unchunked cube."""
    import torch
    from cngi_prototype_b200 import synth, imaging, distributed as D
    d = synth.config_c1(n_time=30, n_chan=7)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD * 1.25
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    gp = {"image_size": [96, 90], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.25, "chan_mode": "cube"}
    ref = imaging.make_image(ds, gp, weight_key="WEIGHT")
    for chunk in (1, 3, 7):
        out = imaging.make_image(ds, gp, weight_key="WEIGHT", chan_chunk=chunk, time_chunk=13)
        assert out["IMAGE"].shape == ref["IMAGE"].shape == (96, 90, 7, 2)
        assert rel_err(out["IMAGE"], ref["IMAGE"]) <= 1e-12 and rel_err(out["SUM_WEIGHT"], ref["SUM_WEIGHT"]) <= 1e-12
    psf = imaging.make_psf(ds, gp, weight_key="WEIGHT")
    psf3 = imaging.make_psf(ds, gp, weight_key="WEIGHT", chan_chunk=3)
    assert rel_err(psf3["PSF"], psf["PSF"]) <= 1e-12
    # the distributed driver with the CUDA operators
    g = dict(gp)
    assert imaging._check_grid_parms(g)
    g.update(oversampling=100, support=7, do_psf=False, complex_grid=True, do_imaging_weight=False)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    T = {"vis": torch.as_tensor(d["vis"]).cuda(), "uvw": torch.as_tensor(d["uvw"]).cuda(),
         "weight": torch.as_tensor(d["weight"]).cuda(), "freq_chan": torch.as_tensor(d["freq_chan"]).cuda()}
    ops = D.cuda_ops()
    zeros = ops.zeros
    ops.zeros = lambda shape, cplx: zeros(shape, cplx, "f64")
    img, sw, (clo, chi) = D.cube_imaging(ops, T, g, cgk, chan_chunk=2)
    assert (clo, chi) == (0, 7)
    assert rel_err(img.cpu().numpy(), ref["IMAGE"]) <= 1e-12 and rel_err(sw.cpu().numpy(), ref["SUM_WEIGHT"]) <= 1e-12
    # image + psf cubes from one fused pass per chunk (ops.grid_image_psf -> cngi_b200_standard_grid_image_psf)
    img, sw, pimg, psw, _ = D.cube_imaging(ops, T, g, cgk, chan_chunk=3, with_psf=True)
    assert rel_err(img.cpu().numpy(), ref["IMAGE"]) <= 1e-12 and rel_err(pimg.cpu().numpy(), psf["PSF"]) <= 1e-12
    assert rel_err(psw.cpu().numpy(), psf["PSF_SUM_WEIGHT"]) <= 1e-12
    # gridding of chunk j + 1 under the transform of chunk j (two grid buffers, side stream): same cubes
    for kw in ({}, {"with_psf": True}):
        res = D.cube_imaging(ops, T, g, cgk, chan_chunk=2, overlap=True, **kw)
        torch.cuda.synchronize()
        assert rel_err(res[0].cpu().numpy(), ref["IMAGE"]) <= 1e-12 and rel_err(res[1].cpu().numpy(), ref["SUM_WEIGHT"]) <= 1e-12
        if kw:
            assert rel_err(res[2].cpu().numpy(), psf["PSF"]) <= 1e-12


@pytest.mark.gpu
def test_continuum_pipeline_side_stream_and_fused_weights_match_the_plain_pipeline():
    """ContinuumPipeline with the weight chain on a concurrent high-priority stream, and with the weight degrid folded
    into the gridder (fuse_weights, the default), == the single-stream pipeline that materialises the imaging weights."""
    import torch
    from types import SimpleNamespace
    from cngi_prototype_b200 import synth, distributed as D
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    d = synth.config_c2(n_time=24, n_chan=16, dtype="f32")
    n = 256
    T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")}
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    gp_iw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)
    cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, 7)).cuda()

    def make_bufs():
        return SimpleNamespace(density=torch.empty((1, 2, n, n), dtype=torch.float64, device="cuda"),
                               dsw=torch.empty((1, 2), dtype=torch.float64, device="cuda"),
                               grid=torch.empty((1, 2, n, n), dtype=torch.complex64, device="cuda"),
                               gsw=torch.empty((1, 2), dtype=torch.float64, device="cuda"))
    res = []
    for side, fuse in ((None, False), (torch.cuda.Stream(priority=-1), False), (None, True)):
        pipe = D.ContinuumPipeline(D.cuda_ops(), gp, gp_iw, dict(weighting="briggs", robust=0.5), cgk, make_bufs,
                                   side_stream=side, fuse_weights=fuse)
        assert pipe.fuse_weights == fuse
        for _ in range(4):
            pipe.step(T)
        iw = pipe.flush()
        torch.cuda.synchronize()
        res.append((pipe.last.grid.cpu().numpy().copy(), pipe.last.gsw.cpu().numpy().copy(),
                    None if fuse else iw.cpu().numpy().copy()))
    assert rel_err(res[1][0], res[0][0]) <= 1e-5 and rel_err(res[1][1], res[0][1]) <= 1e-12
    assert rel_err(res[2][0], res[0][0]) <= 1e-5 and rel_err(res[2][1], res[0][1]) <= 1e-12
    assert np.array_equal(res[2][0] != 0, res[0][0] != 0)
    a, b = res[1][2], res[0][2]   # fp64 atomics land in a different order: equal to rounding, not bitwise
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(b)
    assert np.max(np.abs(a[m] - b[m])) <= 1e-6 * np.max(np.abs(b[m]))


@pytest.mark.gpu
def test_flag_policy_image_and_psf_use_the_same_samples(oracle):
    """FLAG set on FINITE data, apply_flags not run.  Default (reference behaviour): FLAG is not read by either
    function.  apply_flags=True: the flagged samples leave the image AND the psf, so SUM_WEIGHT == PSF_SUM_WEIGHT."""
    from cngi_prototype_b200 import synth, imaging
    d = synth.make_vis_set(9, 24, 5, 2, 1e9, 1.1e9, 300.0, 120.0, seed=77, flag_frac=0.0, bad_rows=False)
    flag = np.random.default_rng(3).random(d["vis"].shape) < 0.1
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD * 1.25
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "FLAG": flag, "chan": d["freq_chan"]}
    gp = {"image_size": [120, 120], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.25, "chan_mode": "cube"}
    img, psf = imaging.make_image(ds, gp, weight_key="WEIGHT"), imaging.make_psf(ds, gp, weight_key="WEIGHT")
    assert rel_err(img["SUM_WEIGHT"], psf["PSF_SUM_WEIGHT"]) < 1e-13           # nothing flagged: same sample set
    imgf = imaging.make_image(ds, gp, weight_key="WEIGHT", apply_flags=True)
    psff = imaging.make_psf(ds, gp, weight_key="WEIGHT", apply_flags=True, time_chunk=7)
    assert rel_err(imgf["SUM_WEIGHT"], psff["PSF_SUM_WEIGHT"]) < 1e-13          # flagged: still the same sample set
    assert np.all(imgf["SUM_WEIGHT"] < img["SUM_WEIGHT"])
    # == running apply_flags first (DATA and WEIGHT NaN where flagged), through the oracle
    g = dict(gp)
    assert imaging._check_grid_parms(g)
    g.update(oversampling=100, support=7, do_imaging_weight=False)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    _, s_ref = oracle._standard_grid_numpy_wrap(oracle.apply_flags_variable(d["vis"], flag), d["uvw"],
                                                oracle.apply_flags_variable(d["weight"], flag), d["freq_chan"], cgk,
                                                dict(g, do_psf=False, complex_grid=True))
    assert rel_err(imgf["SUM_WEIGHT"], s_ref) < 1e-13
    gr = imaging.make_grid(ds, gp, weight_key="WEIGHT", apply_flags=True)
    assert rel_err(gr["SUM_WEIGHT"], s_ref) < 1e-13


@pytest.mark.gpu
def test_host_datasets_lazy_weights_and_pinned_inputs(oracle):
    """numpy datasets: make_imaging_weight returns IMAGING_WEIGHT as a LazyDeviceArray (the reference returns a lazy dask
    variable), make_grid consumes the device copy; page-locked and pageable inputs, explicit and default time chunks,
    and the all-device path give the same numbers."""
    import torch
    from cngi_prototype_b200 import synth, imaging
    from cngi_prototype_b200._lazy import LazyDeviceArray
    d = synth.config_c1(n_time=40, n_chan=6)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD
    gp = {"image_size": [160, 160], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.2, "chan_mode": "continuum"}
    iwp = {"weighting": "briggs", "robust": 0.5}
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    pinned = {k: torch.as_tensor(v).pin_memory().numpy() for k, v in ds.items()}
    dev = {k: torch.as_tensor(v).cuda() for k, v in ds.items()}
    ref_w = imaging.make_imaging_weight(dev, iwp, gp)
    ref_g = imaging.make_grid(ref_w, gp)
    for src, tc in ((ds, 0), (pinned, 0), (pinned, 7)):
        w = imaging.make_imaging_weight(src, iwp, gp, time_chunk=tc)
        assert isinstance(w["IMAGING_WEIGHT"], LazyDeviceArray) and w["IMAGING_WEIGHT"].shape == d["weight"].shape
        g = imaging.make_grid(w, gp, time_chunk=tc)
        assert isinstance(g["GRID"], np.ndarray) and g["GRID"].shape == (192, 192, 1, 2)
        assert rel_err(g["GRID"], ref_g["GRID"].cpu().numpy()) < 1e-13
        assert rel_err(g["SUM_WEIGHT"], ref_g["SUM_WEIGHT"].cpu().numpy()) < 1e-13
        iw = np.asarray(w["IMAGING_WEIGHT"])                    # materialises once, reads like numpy
        iw_ref = ref_w["IMAGING_WEIGHT"].cpu().numpy()           # (the density sums in atomic order: last-bit differences)
        assert np.array_equal(np.isnan(iw), np.isnan(iw_ref)) and rel_err(np.nan_to_num(iw), np.nan_to_num(iw_ref)) < 1e-13
        assert w["IMAGING_WEIGHT"][3, 2].shape == (6, 2)
        img = imaging.make_image(w, gp, time_chunk=tc)
        psf = imaging.make_psf(w, gp, time_chunk=tc)
        assert isinstance(img["IMAGE"], np.ndarray) and img["IMAGE"].shape == (160, 160, 1, 2)
        assert abs(psf["PSF"][80, 80, 0, 0] - 1.0) < 1e-3
    # a weight array the caller made by hand (not lazy) takes the upload path
    w2 = dict(ds, IMAGING_WEIGHT=iw)
    g2 = imaging.make_grid(w2, gp)
    assert rel_err(g2["GRID"], ref_g["GRID"].cpu().numpy()) < 1e-13


@pytest.mark.gpu
def test_chunk_operators_and_api_from_several_threads(oracle):
    """dask runs the reference's chunk functions from a thread pool (they are nogil, _standard_grid.py:242): the operators
    and the host-array API must give the same result when called from several threads at once."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from cngi_prototype_b200 import synth, imaging, _standard_grid as sg
    d = synth.config_c1(n_time=30, n_chan=6)
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(160, d["cell"], chan_mode="cube")
    g_ref, s_ref = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    chunks = [slice(t, t + 5) for t in range(0, 30, 5)]

    def chunk_task(sl):
        torch.cuda.set_device(0)
        return sg._standard_grid_numpy_wrap(d["vis"][sl], d["uvw"][sl], d["weight"][sl], d["freq_chan"], cgk, gp)

    with ThreadPoolExecutor(6) as pool:
        for _ in range(3):
            parts = list(pool.map(chunk_task, chunks))
            g = sum(p[0].astype(np.complex128) for p in parts)
            s = sum(p[1] for p in parts)
            assert rel_err(g, g_ref) < 1e-12 and rel_err(s, s_ref) < 1e-12 and np.array_equal(g != 0, g_ref != 0)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD
    agp = {"image_size": [160, 160], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.0, "chan_mode": "cube"}
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}

    def api_task(_):
        torch.cuda.set_device(0)
        w = imaging.make_imaging_weight(ds, {"weighting": "briggs", "robust": 0.5}, agp)
        return imaging.make_grid(w, agp, time_chunk=4)

    one = api_task(0)
    with ThreadPoolExecutor(4) as pool:
        for r in pool.map(api_task, range(8)):
            assert rel_err(r["GRID"], one["GRID"]) < 1e-12 and rel_err(r["SUM_WEIGHT"], one["SUM_WEIGHT"]) < 1e-13


@pytest.mark.gpu
def test_host_cube_with_deferred_weights_chunked_by_channel(oracle):
    """numpy cube dataset: make_imaging_weight (deferred weights) -> make_image / make_psf walked `chan_chunk` image
    channels at a time == the unchunked result; the weights are only materialised when the psf (or the caller) needs them."""
    from cngi_prototype_b200 import synth, imaging
    d = synth.config_c1(n_time=30, n_chan=7)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD * 1.2
    gp = {"image_size": [128, 128], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.2, "chan_mode": "cube"}
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    w = imaging.make_imaging_weight(ds, {"weighting": "uniform"}, gp)
    assert not w["IMAGING_WEIGHT"].computed
    full = imaging.make_image(w, gp)
    assert not w["IMAGING_WEIGHT"].computed                      # formed inside the gridder, never written
    part = imaging.make_image(w, gp, chan_chunk=3)
    assert w["IMAGING_WEIGHT"].computed                          # the channel slices needed the array
    assert full["IMAGE"].shape == part["IMAGE"].shape == (128, 128, 7, 2)
    assert rel_err(part["IMAGE"], full["IMAGE"]) < 1e-12 and rel_err(part["SUM_WEIGHT"], full["SUM_WEIGHT"]) < 1e-13
    p1, p2 = imaging.make_psf(w, gp), imaging.make_psf(w, gp, chan_chunk=2)
    assert rel_err(p2["PSF"], p1["PSF"]) < 1e-12
