"""The reference's xds signatures (cngi_prototype_b200/xds.py) against a duck-typed Dataset stand-in.

xarray is not importable here, so `MiniDataset` below provides the handful of members the adapters use -- attrs,
item access / assignment, coords, copy(deep=True) -- exactly the members an xarray.Dataset offers under the same names.
The sel_parms / data_groups resolution (cngi/_utils/_check_parms.py:122-223) is checked on the CPU; the compute on the GPU."""
import copy

import numpy as np
import pytest

from _util import rel_err


class MiniDataset:
    def __init__(self, variables=None, coords=None, attrs=None):
        self.vars, self.coords, self.attrs = dict(variables or {}), dict(coords or {}), dict(attrs or {})

    def __getitem__(self, name):
        return self.vars[name]

    def __setitem__(self, name, value):
        self.vars[name] = value

    def __contains__(self, name):
        return name in self.vars

    def copy(self, deep=True):
        return MiniDataset(dict(self.vars), dict(self.coords), copy.deepcopy(self.attrs))   # arrays shared, metadata copied


def _mxds(d, extra_group=False):
    from cngi_prototype_b200.xds import Variable
    dims = ("time", "baseline", "chan", "pol")
    groups = {"1": {"id": "1", "data": "DATA", "uvw": "UVW", "weight": "DATA_WEIGHT"}}
    v = {"DATA": Variable(d["vis"], dims), "UVW": Variable(d["uvw"], ("time", "baseline", "uvw_index")),
         "DATA_WEIGHT": Variable(d["weight"], dims)}
    if extra_group:
        v["CORRECTED_DATA"] = Variable(d["vis"] * 2, dims)
        groups["2"] = {"id": "2", "data": "CORRECTED_DATA", "uvw": "UVW", "weight": "DATA_WEIGHT"}
    xds0 = MiniDataset(v, {"chan": Variable(d["freq_chan"], ("chan",))}, {"data_groups": [groups]})
    return MiniDataset(attrs={"xds0": xds0})


def test_sel_parms_resolution_mirrors_reference_rules():
    from cngi_prototype_b200 import synth
    from cngi_prototype_b200.xds import _check_sel_parms
    d = synth.config_c1(n_time=3, n_chan=2)
    xds = _mxds(d, extra_group=True).attrs["xds0"]
    s = {"xds": "xds0"}
    _check_sel_parms(xds, s)                                     # defaults: first group in, a new id out
    assert s["data_group_in"] == {"id": "1", "data": "DATA", "uvw": "UVW", "weight": "DATA_WEIGHT"}
    assert s["data_group_out"]["id"] == "3"
    s = {"xds": "xds0", "data_group_in_id": 2}
    _check_sel_parms(xds, s, new_or_modified_data_variables={"imaging_weight": "IMAGING_WEIGHT"}, append_to_in_id=True)
    assert s["data_group_in"]["data"] == "CORRECTED_DATA" and s["data_group_out"]["id"] == "2"
    assert s["data_group_out"]["imaging_weight"] == "IMAGING_WEIGHT" and s["data_group_out"]["weight"] == "DATA_WEIGHT"
    s = {"xds": "xds0", "imaging_weight": "IW_BRIGGS", "data_group_out_id": 7}
    _check_sel_parms(xds, s, new_or_modified_data_variables={"imaging_weight": "IMAGING_WEIGHT"}, append_to_in_id=True)
    assert s["data_group_out"]["imaging_weight"] == "IW_BRIGGS" and s["data_group_out"]["id"] == "7"
    with pytest.raises(AssertionError):                          # unknown group id
        _check_sel_parms(xds, {"xds": "xds0", "data_group_in_id": 9})
    with pytest.raises(AssertionError):                          # a variable of the group is missing from the dataset
        _check_sel_parms(xds, {"xds": "xds0", "data_group_in": {"weight": "NO_SUCH_WEIGHT"}})
    # an empty image dataset gets a data group; a name another group already points at cannot be overwritten
    img = MiniDataset()
    si = {}
    _check_sel_parms(img, si, new_or_modified_data_variables={"sum_weight": "SUM_WEIGHT", "image": "IMAGE"}, append_to_in_id=True)
    assert img.attrs["data_groups"] == [{"0": {"id": "0"}}] and si["data_group_out"] == {"id": "0", "sum_weight": "SUM_WEIGHT", "image": "IMAGE"}
    xds.attrs["data_groups"][0]["1"]["imaging_weight"] = "IMAGING_WEIGHT"
    xds["IMAGING_WEIGHT"] = xds["DATA_WEIGHT"]
    with pytest.raises(AssertionError):
        _check_sel_parms(xds, {"xds": "xds0", "data_group_in_id": 2}, new_or_modified_data_variables={"imaging_weight": "IMAGING_WEIGHT"},
                         append_to_in_id=True)


def test_natural_weighting_reuses_the_weight_variable_without_a_gpu():
    from cngi_prototype_b200 import synth, xds as X
    d = synth.config_c1(n_time=3, n_chan=2)
    mx = _mxds(d)
    out = X.make_imaging_weight(mx, {"weighting": "natural"}, {"image_size": [64, 64], "cell_size": [1.0, 1.0]}, {"xds": "xds0"})
    assert out is not mx and "imaging_weight" not in mx.attrs["xds0"].attrs["data_groups"][0]["1"]          # input untouched
    assert out.attrs["xds0"].attrs["data_groups"][0]["1"]["imaging_weight"] == "DATA_WEIGHT"
    with pytest.raises(AssertionError):
        X.make_imaging_weight(mx, {"weighting": "natural"}, {}, {})                                       # 'xds' is required


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_xds_chain_matches_the_mapping_api_and_oracle(oracle, mode):
    from cngi_prototype_b200 import synth, imaging, xds as X
    d = synth.config_c1(n_time=30, n_chan=5)
    cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD * 1.25
    gp = {"image_size": [128, 120], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.25, "chan_mode": mode}
    iwp = {"weighting": "briggs", "robust": 0.5}
    mx = _mxds(d)
    mx2 = X.make_imaging_weight(mx, iwp, gp, {"xds": "xds0"})
    x2 = mx2.attrs["xds0"]
    assert "IMAGING_WEIGHT" in x2 and "IMAGING_WEIGHT" not in mx.attrs["xds0"]
    assert x2["IMAGING_WEIGHT"].dims == ("time", "baseline", "chan", "pol")
    assert x2.attrs["data_groups"][0]["1"]["imaging_weight"] == "IMAGING_WEIGHT"
    img0 = MiniDataset()
    img1 = X.make_psf(mx2, img0, gp, {"xds": "xds0"}, {})
    img2 = X.make_image(mx2, img1, gp, {"xds": "xds0"}, {"image": "DIRTY"})
    img3 = X.make_grid(mx2, img2, gp, {"xds": "xds0"}, {})
    assert img0.vars == {} and "DIRTY" not in img1                                                         # inputs untouched
    grp = img3.attrs["data_groups"][0]["0"]
    assert grp["psf"] == "PSF" and grp["image"] == "DIRTY" and grp["grid"] == "GRID"
    n_ic = 5 if mode == "cube" else 1
    assert img3["DIRTY"].dims == ("l", "m", "time", "chan", "pol") and img3["DIRTY"].shape == (128, 120, 1, n_ic, 2)
    assert img3["GRID"].dims == ("u", "v", "time", "chan", "pol") and img3["GRID"].shape == (160, 150, 1, n_ic, 2)
    assert img3["SUM_WEIGHT"].dims == ("time", "chan", "pol") and img3["PSF_SUM_WEIGHT"].shape == (1, n_ic, 2)
    # the same chain through the mapping API
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]}
    ref_w = imaging.make_imaging_weight(ds, iwp, gp)
    ref_i, ref_p = imaging.make_image(ref_w, gp), imaging.make_psf(ref_w, gp)
    assert rel_err(img3["DIRTY"].values[:, :, 0], ref_i["IMAGE"]) < 1e-12
    assert rel_err(img3["PSF"].values[:, :, 0], ref_p["PSF"]) < 1e-12
    assert rel_err(img3["SUM_WEIGHT"].values[0], ref_i["SUM_WEIGHT"]) < 1e-13
    # and the oracle, for the image
    g = dict(gp)
    assert imaging._check_grid_parms(g)
    gw = dict(g, image_size_padded=g["image_size"], oversampling=0, support=1, do_psf=True, complex_grid=False, do_imaging_weight=True)
    rho, sw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gw)
    bf = oracle._calculate_briggs_parms(rho, sw, iwp)
    iw = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"], bf, d["freq_chan"], gw)
    gi = dict(g, oversampling=100, support=7, do_psf=False, complex_grid=True, do_imaging_weight=False)
    gg, ss = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw, d["freq_chan"], oracle._create_prolate_spheroidal_kernel_1D(100, 7), gi)
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(g["image_size_padded"]), g["image_size"])
    ref = oracle.correct_image(oracle.grid_to_uncorrected_image(gg, g["image_size"]), ss, corr)
    assert rel_err(img3["DIRTY"].values[:, :, 0], ref) < 1e-11
    # a second data group (e.g. corrected data) is selected by id
    mx3 = _mxds(d, extra_group=True)
    img4 = X.make_image(mx3, MiniDataset(), gp, {"xds": "xds0", "data_group_in_id": 2}, {})
    img5 = X.make_image(mx3, MiniDataset(), gp, {"xds": "xds0"}, {})
    assert rel_err(img4["IMAGE"].values, 2 * img5["IMAGE"].values) < 1e-12
