"""BASELINE config 3 chain on the GPU (direction_rotate -> GCF -> make_mosaic_pb -> make_image_with_gcf /
make_psf_with_gcf) against the same chain assembled from the oracle's restatements of
/root/reference/ngcasa/imaging/{direction_rotate,make_gridding_convolution_function,make_mosaic_pb,
make_image_with_gcf,make_psf_with_gcf}.py and _imaging_utils/{_aperture_grid,_normalize}.py.
fp64: 1e-9 relative on images (FFT of a 240^2 grid after three chained stages), masks bit-exact."""
import numpy as np
import pytest

from _util import rel_err

pytestmark = pytest.mark.gpu


def _problem(seed=3, n_time=12, n_chan=4, n_pol=2):
    rng = np.random.default_rng(seed)
    n_ant = 9
    types = np.array([0, 0, 0, 1, 0, 1, 0, 0, 1])
    a1, a2 = np.triu_indices(n_ant, 1)
    n_b = len(a1)
    uvw = rng.normal(0.0, 110.0, (n_time, n_b, 3))
    uvw[:, :, 2] *= 0.05
    uvw[2, 5, 0] = np.nan
    vis = rng.standard_normal((n_time, n_b, n_chan, n_pol)) + 1j * rng.standard_normal((n_time, n_b, n_chan, n_pol))
    w = rng.uniform(0.5, 1.5, (n_time, n_b, n_chan, n_pol))
    w[rng.random(w.shape) < 0.01] = 0.0
    flag = (rng.random(w.shape) < 0.03)
    ids = np.array([0, 1, 2])
    dirs = np.array([[1.0, 0.5], [1.00003, 0.50002], [0.99996, 0.49999]])
    field = np.repeat(ids[np.arange(n_time) % 3][:, None], n_b, 1).astype(np.int64)
    field[1, 3] = -2147483648
    field[n_time - 1, 0] = -2147483648
    vis_ds = dict(DATA=vis, UVW=uvw, WEIGHT=w, FLAG=flag, FIELD_ID=field, chan=np.linspace(100.0e9, 101.1e9, n_chan))
    field_ds = dict(field_id=ids, PHASE_DIR=dirs)
    rotation_parms = dict(new_phase_center=[1.0, 0.5], common_tangent_reprojection=True, single_precision=False)
    gcf_parms = dict(function="casa_airy", list_dish_diameters=np.array([10.7, 6.25]),
                     list_blockage_diameters=np.array([0.75, 0.75]), unique_ant_indx=types,
                     basline_ant=np.stack([a1, a2], 1), pol=np.arange(n_pol), oversampling=[5, 5], max_support=[11, 11])
    return vis_ds, field_ds, rotation_parms, gcf_parms


def _oracle_chain(O, vis_ds, field_ds, rotation_parms, gcf_parms, grid_parms, norm_parms, apply_flags=False):
    gp = dict(grid_parms)
    gp["image_size"] = np.array(gp["image_size"])
    gp["image_size_padded"] = (gp["fft_padding"] * gp["image_size"]).astype(int)
    gp["cell_size"] = np.array(gp["cell_size"]) * np.pi / (3600 * 180) * np.array([-1.0, 1.0])
    ctr = rotation_parms["common_tangent_reprojection"]
    R, P, rid = O.calc_rotation_mats(vis_ds["FIELD_ID"], field_ds["field_id"], field_ds["PHASE_DIR"],
                                     rotation_parms["new_phase_center"], ctr)
    uvw = O.apply_rotation_matrix(vis_ds["UVW"], vis_ds["FIELD_ID"], R, rid)
    vis = O.apply_phasor(vis_ds["DATA"], uvw, vis_ds["FIELD_ID"], vis_ds["chan"], P, rid, ctr,
                         rotation_parms["single_precision"])
    weight = vis_ds["WEIGHT"]
    if apply_flags:       # cngi/vis/apply_flags.py:53 NaNs every variable with FLAG's dims: DATA and WEIGHT together
        vis = np.where(vis_ds["FLAG"], np.nan, vis)
        weight = np.where(vis_ds["FLAG"], np.nan, weight)
    g = O.make_gridding_convolution_function(dict(gcf_parms, freq_chan=vis_ds["chan"], field_phase_dir=field_ds["PHASE_DIR"],
                                                  phase_center=np.array(rotation_parms["new_phase_center"])), gp)
    agp = dict(gp, oversampling=g["oversampling"], field_id=np.asarray(field_ds["field_id"]), do_psf=False)
    common = (uvw, weight, vis_ds["FIELD_ID"], g["CF_BASELINE_MAP"], g["CF_CHAN_MAP"], g["CF_POL_MAP"])
    tail = (g["SUPPORT"], g["PHASE_GRADIENT"], vis_ds["chan"])
    wg, wsw = O._aperture_weight_grid_numpy_wrap(*common, g["WEIGHT_CONV_KERNEL"], *tail, agp)
    sw1 = wsw.copy()
    sw1[sw1 == 0] = 1
    weight_pb = O.grid_to_uncorrected_image(wg, gp["image_size"]) / sw1[None, None]
    pb = np.sqrt(np.abs(weight_pb))
    ig, isw = O._aperture_grid_numpy_wrap(vis, *common, g["CONV_KERNEL"], *tail, agp)
    pg, psw = O._aperture_psf_grid_numpy_wrap(*common, g["CONV_KERNEL"], *tail, dict(agp, do_psf=True))
    norm = {"flat_noise": pb, "flat_sky": weight_pb, "none": np.ones_like(pb)}[norm_parms["norm_type"]]
    out = {}
    for name, grid, sw in (("IMAGE", ig, isw), ("PSF", pg, psw)):
        img = O.normalize_image(O.grid_to_uncorrected_image(grid, gp["image_size"]), sw, norm, g["oversampling"])
        if norm_parms["pb_limit"] > 0:
            img[pb < norm_parms["pb_limit"]] = 0.0
        if norm_parms["single_precision"]:
            img = img.astype(np.float32).astype(np.float64)
        out[name] = img
    c = gp["image_size"] // 2
    out["PSF"] = out["PSF"] / out["PSF"][c[0], c[1], :, :]
    out.update(PB=pb, WEIGHT_PB=weight_pb, WEIGHT_PB_SUM_WEIGHT=wsw, SUM_WEIGHT=isw, PSF_SUM_WEIGHT=psw)
    return out, g


@pytest.mark.parametrize("chan_mode,norm_type,single,apply_flags", [("cube", "flat_sky", False, True),
                                                                     ("continuum", "flat_noise", False, False),
                                                                     ("cube", "none", True, True)])
def test_mosaic_chain(oracle, chan_mode, norm_type, single, apply_flags):
    from cngi_prototype_b200 import mosaic
    vis_ds, field_ds, rotation_parms, gcf_parms = _problem()
    grid_parms = dict(image_size=[200, 200], cell_size=[0.55, 0.55], fft_padding=1.2, chan_mode=chan_mode)
    norm_parms = dict(norm_type=norm_type, pb_limit=0.2, single_precision=single)
    img, gcf, _ = mosaic.mosaic_imaging(vis_ds, field_ds, rotation_parms, gcf_parms, grid_parms, norm_parms, time_chunk=5,
                                        apply_flags=apply_flags)
    ref, g = _oracle_chain(oracle, vis_ds, field_ds, rotation_parms, gcf_parms, grid_parms, norm_parms, apply_flags)
    assert np.array_equal(gcf["SUPPORT"].cpu().numpy(), g["SUPPORT"])
    for k in ("WEIGHT_PB_SUM_WEIGHT", "SUM_WEIGHT", "PSF_SUM_WEIGHT"):
        assert rel_err(img[k], ref[k]) < 1e-11, k
    for k in ("WEIGHT_PB", "PB"):
        assert img[k].shape == ref[k].shape and rel_err(img[k], ref[k]) < 1e-9, k
    tol = 3e-7 if single else 1e-9
    for k in ("IMAGE", "PSF"):
        assert img[k].shape == ref[k].shape
        assert np.mean((img[k] == 0) != (ref[k] == 0)) < 1e-4, k      # pb_limit mask (a pixel within 1e-12 of the limit may flip)
        same = (img[k] == 0) == (ref[k] == 0)
        assert rel_err(np.where(same, img[k], 0), np.where(same, ref[k], 0)) < tol, k
    c = 100
    np.testing.assert_allclose(img["PSF"][c, c], 1.0, rtol=1e-6)


def test_api_functions_separately_torch(oracle):
    """make_mosaic_pb / make_image_with_gcf on CUDA tensors with a ready-made gcf_dataset and img_dataset."""
    import torch
    from cngi_prototype_b200 import mosaic, direction_rotate as dr, make_gridding_convolution_function as mg
    vis_ds, field_ds, rotation_parms, gcf_parms = _problem(seed=9, n_time=6, n_chan=2)
    grid_parms = dict(image_size=[200, 200], cell_size=[0.55, 0.55], fft_padding=1.2, chan_mode="cube")
    norm_parms = dict(norm_type="flat_sky", pb_limit=0.0, single_precision=False)
    ref, _ = _oracle_chain(oracle, vis_ds, field_ds, rotation_parms, gcf_parms, grid_parms, norm_parms)
    dev = torch.device("cuda")
    tv = {k: torch.as_tensor(v, device=dev) for k, v in vis_ds.items()}
    rot = dr.direction_rotate(tv, field_ds, rotation_parms)
    tv["DATA"], tv["UVW"] = rot["DATA_ROT"], rot["UVW_ROT"]
    n_pad = np.array([240, 240])
    cell = np.array([-0.55, 0.55]) * np.pi / (3600 * 180)
    gcf = mg.make_gridding_convolution_function(
        dict(gcf_parms, freq_chan=vis_ds["chan"], field_phase_dir=field_ds["PHASE_DIR"], field_id=field_ds["field_id"],
             phase_center=np.array([1.0, 0.5])), dict(image_size=np.array([200, 200]), image_size_padded=n_pad, cell_size=cell))
    pbs = mosaic.make_mosaic_pb(tv, gcf, grid_parms)
    assert pbs["PB"].is_cuda and rel_err(pbs["PB"].cpu().numpy(), ref["PB"]) < 1e-9
    img = mosaic.make_image_with_gcf(tv, gcf, pbs, grid_parms, norm_parms)
    assert img["IMAGE"].is_cuda and rel_err(img["IMAGE"].cpu().numpy(), ref["IMAGE"]) < 1e-9
