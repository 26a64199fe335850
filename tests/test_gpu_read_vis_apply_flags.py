"""N4 on the GPU: cngi_b200_apply_flags against the oracle's where().astype() (bit-exact), the zarr -> pinned -> device
pipeline, and make_image / make_psf streamed from a vis.zarr store against the oracle on the same samples.

Reference: cngi/vis/apply_flags.py:53, cngi/dio/read_vis.py:21, synthesis_imaging_cube.py:180.
"""
import numpy as np
import pytest

from _util import rel_err, same_support

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 1000, 4099, 1 << 20])
def test_apply_flags_bit_exact_all_paths(oracle, dtype, n):
    """Out of place, in place, in place on a mis-aligned flag pointer (head / tail elements), flagged counter."""
    import torch
    from cngi_prototype_b200.apply_flags import apply_flags_chunk
    rng = np.random.default_rng(n + 7)
    x = rng.standard_normal(n + 3).astype(dtype)
    if np.issubdtype(dtype, np.complexfloating):
        x = (x + 1j * rng.standard_normal(n + 3)).astype(dtype)
    flag = rng.random(n + 3) < 0.25
    want = oracle.apply_flags_variable(x[:n], flag[:n])
    out = apply_flags_chunk(x[:n], flag[:n])
    assert out.is_cuda and np.array_equal(_bits(out.cpu().numpy()), _bits(want))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    xd = torch.as_tensor(x).cuda()
    fd = torch.as_tensor(flag).cuda()
    for off in (0, 3):                                   # off = 3: the flag bytes start 3 past a 16-byte boundary
        d = xd[off:off + n].clone()
        r = apply_flags_chunk(d, fd[off:off + n], inplace=True, n_flagged=cnt)
        assert r.data_ptr() == d.data_ptr()
        assert np.array_equal(_bits(d.cpu().numpy()), _bits(oracle.apply_flags_variable(x[off:off + n], flag[off:off + n])))
    assert int(cnt.item()) == int(flag[:n].sum() + flag[3:3 + n].sum())
    # out of place on views that are not 16-byte aligned: the element-wise kernel
    r = apply_flags_chunk(xd[1:1 + n], fd[1:1 + n])
    assert np.array_equal(_bits(r.cpu().numpy()), _bits(oracle.apply_flags_variable(x[1:1 + n], flag[1:1 + n])))
    assert np.array_equal(_bits(xd.cpu().numpy()), _bits(x))          # inputs untouched


def test_apply_flags_all_and_none_flagged_and_errors(oracle):
    import torch
    from cngi_prototype_b200.apply_flags import apply_flags_chunk
    from cngi_prototype_b200 import _lib
    x = np.random.default_rng(0).standard_normal((5, 4, 3, 2)) * (1 + 1j)
    assert np.isnan(apply_flags_chunk(x, np.ones(x.shape, bool)).cpu().numpy()).all()
    assert np.array_equal(apply_flags_chunk(x, np.zeros(x.shape, np.uint8)).cpu().numpy(), x)
    assert np.array_equal(_bits(apply_flags_chunk(x, (np.arange(x.size).reshape(x.shape) % 3) * 7).cpu().numpy()),
                          _bits(oracle.apply_flags_variable(x, np.arange(x.size).reshape(x.shape) % 3)))   # any non-zero flags
    with pytest.raises(TypeError):
        apply_flags_chunk(np.arange(4), np.zeros(4, bool))
    with pytest.raises(ValueError):
        apply_flags_chunk(np.zeros(4), np.zeros(5, bool))
    d = torch.zeros(64, dtype=torch.float64, device="cuda")
    f = torch.zeros(64, dtype=torch.uint8, device="cuda")
    L = _lib.lib()
    assert L.cngi_b200_apply_flags(d.data_ptr(), d.data_ptr() + 8, f.data_ptr(), 32, _lib.ELEM_F64, None, None) == 1
    assert b"overlap" in L.cngi_b200_last_error()
    assert L.cngi_b200_apply_flags(d.data_ptr(), d.data_ptr(), f.data_ptr(), 32, 9, None, None) == 3
    assert L.cngi_b200_apply_flags(None, None, None, 0, _lib.ELEM_F64, None, None) == 0


def _store(tmp_path, n_time=30, n_chan=6, compressor="default"):
    from cngi_prototype_b200 import synth, read_vis as rv
    d = synth.config_c1(n_time=n_time, n_chan=n_chan)      # DATA already holds some NaNs; FLAG adds 5 %
    rng = np.random.default_rng(11)
    flag = rng.random(d["vis"].shape) < 0.05
    xds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "FLAG": flag, "chan": d["freq_chan"]}
    comp = rv.DEFAULT_COMPRESSOR if compressor == "default" else compressor
    store = rv.write_vis(str(tmp_path / "c1.vis.zarr"), xds, chunks={"time": 7, "chan": 4}, compressor=comp)
    return store, d, flag


@pytest.mark.parametrize("native", [True, False], ids=["native", "python"])
@pytest.mark.parametrize("depth", [2, 3])
def test_device_chunk_pipeline_delivers_every_block(tmp_path, depth, native):
    """Blocks arrive in order with the right contents although pinned / device buffer sets are recycled while the
    previous block's kernels may still be running; early exit does not hang the reader."""
    import torch
    from cngi_prototype_b200 import read_vis as rv
    store, d, flag = _store(tmp_path)
    xds = rv.read_vis(store, partition="xds0").xds0
    seen, acc = [], []
    for sl, blk in xds.iter_device_chunks(["DATA", "UVW", "FLAG"], time_chunk=4, depth=depth, workers=4, native=native):
        assert all(t.is_cuda for t in blk.values()) and blk["FLAG"].dtype == torch.uint8
        seen.append(sl)
        acc.append({k: t.clone() for k, t in blk.items()})      # queued on the current stream before the set is reused
    torch.cuda.synchronize()
    assert seen == xds.time_blocks(4) and len(seen) == 8
    for k, ref in (("DATA", d["vis"]), ("UVW", d["uvw"]), ("FLAG", flag.astype(np.uint8))):
        got = torch.cat([a[k] for a in acc]).cpu().numpy()
        assert np.array_equal(_bits(got), _bits(ref)), k
    it = xds.iter_device_chunks(["UVW"], time_chunk=4, depth=depth)
    next(it)
    it.close()                                           # finally-block joins the reader thread


@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_make_image_streamed_from_zarr_matches_oracle(oracle, tmp_path, mode):
    """read_vis -> make_psf / make_image straight from the store (FLAG fused in the gridder) == oracle on the samples with
    apply_flags applied; also == the in-memory path fed by apply_flags (NaN data, no FLAG variable)."""
    from cngi_prototype_b200 import imaging, read_vis as rv, apply_flags as af
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    store, d, flag = _store(tmp_path)
    mxds = rv.read_vis(store, partition="xds0")
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD
    gp = {"image_size": [144, 160], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.25, "chan_mode": mode}
    img = imaging.make_image(mxds.xds0, gp, weight_key="WEIGHT", apply_flags=True)
    psf = imaging.make_psf(mxds.xds0, gp, weight_key="WEIGHT", time_chunk=5, apply_flags=True)
    psf_noflag = imaging.make_psf(mxds.xds0, gp, weight_key="WEIGHT", time_chunk=5)   # the reference never reads FLAG
    # oracle chain on the flagged samples
    g = dict(gp)
    assert imaging._check_grid_parms(g)
    g.update(oversampling=100, support=7)
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    vis_f = oracle.apply_flags_variable(d["vis"], flag)
    grid, sw = oracle._standard_grid_numpy_wrap(vis_f, d["uvw"], d["weight"], d["freq_chan"], cgk,
                                                dict(g, do_psf=False, complex_grid=True, do_imaging_weight=False))
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(g["image_size_padded"]), g["image_size"])
    ref = oracle.correct_image(oracle.grid_to_uncorrected_image(grid, g["image_size"]), sw, corr)
    assert rel_err(img["IMAGE"], ref) < 1e-12 and rel_err(img["SUM_WEIGHT"], sw) < 1e-13
    pgrid, psw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk,
                                                      dict(g, do_psf=True, complex_grid=False, do_imaging_weight=False))
    pref = oracle.correct_image(oracle.grid_to_uncorrected_image(pgrid, g["image_size"]), psw, corr)
    assert rel_err(psf_noflag["PSF"], pref) < 1e-12 and rel_err(psf_noflag["PSF_SUM_WEIGHT"], psw) < 1e-13
    # apply_flags=True == cngi.vis.apply_flags first: the weights are NaN where FLAG is set, for the psf too
    w_f = oracle.apply_flags_variable(d["weight"], flag)
    pgrid, psw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], w_f, d["freq_chan"], cgk,
                                                      dict(g, do_psf=True, complex_grid=False, do_imaging_weight=False))
    pref = oracle.correct_image(oracle.grid_to_uncorrected_image(pgrid, g["image_size"]), psw, corr)
    assert rel_err(psf["PSF"], pref) < 1e-12 and rel_err(psf["PSF_SUM_WEIGHT"], psw) < 1e-13
    # the grid itself: masks exact
    gr = imaging.make_grid(mxds.xds0, gp, weight_key="WEIGHT", apply_flags=True)
    assert same_support(np.moveaxis(gr["GRID"], (2, 3), (0, 1)), grid)
    # apply_flags first, then the in-memory path without a FLAG variable
    fx = af.apply_flags(mxds, "xds0").attrs["xds0"]
    assert set(fx) == {"DATA", "UVW", "WEIGHT", "FLAG", "chan"}
    assert fx["DATA"].is_cuda and fx["WEIGHT"].is_cuda and not hasattr(fx["UVW"], "is_cuda")   # UVW has other dims
    assert np.array_equal(_bits(fx["DATA"].cpu().numpy()), _bits(vis_f))
    assert np.array_equal(_bits(fx["WEIGHT"].cpu().numpy()), _bits(oracle.apply_flags_variable(d["weight"], flag)))
    ds = {"DATA": fx["DATA"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    img2 = imaging.make_image(ds, gp, weight_key="WEIGHT", time_chunk=9)
    assert rel_err(img2["IMAGE"].cpu().numpy(), ref) < 1e-12
