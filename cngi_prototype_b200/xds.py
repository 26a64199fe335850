"""The reference's `xds` API signatures on top of the B200 operators.

    make_imaging_weight(vis_mxds, imaging_weights_parms, grid_parms, sel_parms)      ngcasa/imaging/make_imaging_weight.py:20
    make_grid (vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms)          ngcasa/imaging/make_grid.py:27
    make_psf  (vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms)          ngcasa/imaging/make_psf.py:27
    make_image(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms)          ngcasa/imaging/make_image.py:27

Same argument order, the same `sel_parms` / `data_groups` resolution (cngi/_utils/_check_parms.py:122-223: `xds`,
`data_group_in_id`, `data_group_out_id`, per-variable names, defaults from the first data group, the "modified variable is
referenced by another group" check) and the same outputs (variable names, dims `['l','m','time','chan','pol']` /
`['time','chan','pol']`, `data_groups` bookkeeping, inputs never mutated).  The compute goes through
`cngi_prototype_b200.imaging` -- CUDA only, no CPU fallback.

xarray and dask are not importable in this image, so the adapters are duck-typed instead of import-guarded: they need
  mxds.attrs[name] -> xds,  mxds.copy(deep=True);
  xds.attrs['data_groups'], xds[name] with `.data` (numpy / torch / anything with `.compute()`), `.dims`;
  xds.coords['chan'].values;  xds[name] = variable;  xds.copy(deep=True)
which an xarray.Dataset provides unchanged.  New variables are `xarray.DataArray`s when xarray imports, else `Variable`
below (data + dims).  What is NOT reproduced: the sky-image coordinates that make_image takes from
cngi.image.make_empty_sky_image (:147) and the Gaussian beam fit make_psf ends with (:159) -- image analysis, out of scope.
"""
import copy

import numpy as np

from . import imaging

try:  # pragma: no cover - absent in this image
    import xarray as _xr
except Exception:
    _xr = None


class Variable:
    """Minimal data variable (data + named dims) used when xarray is not importable."""

    def __init__(self, data, dims):
        self.data, self.dims = data, tuple(dims)
        assert len(self.dims) == len(data.shape), (self.dims, tuple(data.shape))

    @property
    def values(self):
        return np.asarray(self.data)

    @property
    def shape(self):
        return tuple(self.data.shape)

    def __deepcopy__(self, memo):
        return Variable(self.data, self.dims)   # arrays are never written in place: share them


def _new_variable(data, dims):
    if _xr is not None:  # pragma: no cover
        return _xr.DataArray(np.asarray(data), dims=list(dims))
    return Variable(data, dims)


def _values(var):
    """The array behind a data variable: dask arrays are computed, everything else is passed through."""
    x = getattr(var, "data", var)
    if hasattr(x, "compute"):
        x = x.compute()
    return x


# -------------------------------------------------------------------------------------------------------------------
#  sel_parms / data_groups resolution (cngi/_utils/_check_parms.py:122-223)
# -------------------------------------------------------------------------------------------------------------------
def _has_variable(xds, name):
    try:
        xds[name]
        return True
    except Exception:
        print("######### ERROR Data array ", name, "can not be found in dataset.")
        return False


def _check_sel_parms(xds, sel_parms, new_or_modified_data_variables=None, required_data_variables=None, append_to_in_id=False):
    """Fills sel_parms['data_group_in'] / ['data_group_out'] in place.

    data_group_in  = the group named by data_group_in_id (default: the first group of xds.attrs['data_groups'][0]), with
                     individual names overridable through sel_parms['data_group_in'];
    data_group_out = data_group_in + the variables this function creates (defaults in new_or_modified_data_variables,
                     overridable by sel_parms[<key>]), under id data_group_out_id | the input id (append_to_in_id) | max id + 1.
    """
    new = dict(new_or_modified_data_variables or {})
    required = dict(required_data_variables or {})
    if "data_groups" not in xds.attrs:          # an empty (image) dataset
        gid = str(sel_parms["data_group_out_id"]) if "data_group_out_id" in sel_parms else "0"
        xds.attrs["data_groups"] = [{gid: {"id": gid}}]
        append_to_in_id = True
    groups = xds.attrs["data_groups"][0]
    ids = [int(k) for k in groups]
    user_in = dict(sel_parms.get("data_group_in", {}))
    user_out = dict(sel_parms.get("data_group_out", {}))
    if "data_group_in_id" in sel_parms:
        gid = str(sel_parms["data_group_in_id"])
        assert int(gid) in ids, "######### ERROR: data_group_in id does not exist in " + str(sel_parms.get("xds", "xds"))
        base_in = copy.deepcopy(groups[gid])
    else:
        base_in = copy.deepcopy(list(groups.values())[0])
    group_in = {**required, **base_in, **{k: v for k, v in user_in.items() if k != "id"}}
    if "data_group_out_id" in sel_parms:
        out_id = str(sel_parms["data_group_out_id"])
    elif append_to_in_id:
        out_id = str(group_in["id"])
    else:
        out_id = str(max(ids) + 1)
    named = {k: sel_parms[k] for k in new if k in sel_parms}                       # e.g. sel_parms['image'] = 'MY_IMAGE'
    named.update({k: v for k, v in user_out.items() if k in new})
    group_out = {**group_in, "id": out_id, **new, **named}
    sel_parms["data_group_in"], sel_parms["data_group_out"] = group_in, group_out
    ok = all(_has_variable(xds, v) for k, v in group_in.items() if k not in ("id", "properties") and not isinstance(v, dict))
    assert ok, "######### ERROR: sel_parms checking failed"
    for key in new:   # a variable this function (re)writes must not be what another data group points at
        for gid, grp in groups.items():
            assert gid == out_id or grp.get(key) != group_out[key], \
                "Data variables, that are modified by the function, can not be replaced if they are referenced in another data_group"
    return True


def _register_group(xds, group_out):
    xds.attrs["data_groups"][0] = {**xds.attrs["data_groups"][0], group_out["id"]: group_out}


def _vis_mapping(vis_xds, group_in, need_data=True):
    """The plain mapping cngi_prototype_b200.imaging works on, from the variables a data group names."""
    ds = {"UVW": _values(vis_xds[group_in["uvw"]]), "WEIGHT": _values(vis_xds[group_in["weight"]]),
          "chan": np.asarray(vis_xds.coords["chan"].values, dtype=np.float64)}
    if need_data:
        ds["DATA"] = _values(vis_xds[group_in["data"]])
    if "imaging_weight" in group_in:
        ds["IMAGING_WEIGHT"] = _values(vis_xds[group_in["imaging_weight"]])
    return ds


def _select_vis(vis_mxds, vis_sel_parms):
    _mxds = vis_mxds.copy(deep=True)
    _sel = copy.deepcopy(vis_sel_parms)
    assert "xds" in _sel, "######### ERROR: xds must be specified in sel_parms"   # xds names are not fixed: no default
    return _mxds, _mxds.attrs[_sel["xds"]], _sel


# -------------------------------------------------------------------------------------------------------------------
#  API functions
# -------------------------------------------------------------------------------------------------------------------
def make_imaging_weight(vis_mxds, imaging_weights_parms, grid_parms, sel_parms):
    """Returns a copy of vis_mxds whose selected xds carries the imaging weights (make_imaging_weight.py:20-113).

    sel_parms: 'xds' (required), 'data_group_in_id', 'data_group_out_id' (default: the input group), 'imaging_weight'
    (variable name, default 'IMAGING_WEIGHT').  Natural weighting reuses the weight variable (no new array)."""
    print("######################### Start make_imaging_weights #########################")
    _mxds, _vis_xds, _sel = _select_vis(vis_mxds, sel_parms)
    _iwp = copy.deepcopy(imaging_weights_parms)
    _check_sel_parms(_vis_xds, _sel, new_or_modified_data_variables={"imaging_weight": "IMAGING_WEIGHT"}, append_to_in_id=True)
    assert int(_vis_xds[_sel["data_group_in"]["weight"]].shape[-1]) <= 2, "Full polarization is not supported."   # (:90)
    assert imaging._check_imaging_weights_parms(_iwp), "######### ERROR: imaging_weights_parms checking failed"
    g_in, g_out = _sel["data_group_in"], _sel["data_group_out"]
    if _iwp["weighting"] == "natural":
        g_out["imaging_weight"] = g_in["weight"]
        _register_group(_vis_xds, g_out)
        print("Since weighting is natural input weight will be reused as imaging weight.")
        return _mxds
    ds = _vis_mapping(_vis_xds, {k: v for k, v in g_in.items() if k != "imaging_weight"}, need_data=False)
    out = imaging.make_imaging_weight(ds, _iwp, grid_parms)
    _vis_xds[g_out["imaging_weight"]] = _new_variable(out["IMAGING_WEIGHT"], _vis_xds[g_in["data"]].dims)
    _register_group(_vis_xds, g_out)
    print("######################### Created graph for make_imaging_weight #########################")
    return _mxds


def _imaging_call(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms, new_vars, fn, keys):
    _mxds, _vis_xds, _vsel = _select_vis(vis_mxds, vis_sel_parms)
    _img_xds = img_xds.copy(deep=True)
    _isel = copy.deepcopy(img_sel_parms)
    _gp = copy.deepcopy(grid_parms)
    assert imaging._check_grid_parms(copy.deepcopy(_gp)), "######### ERROR: grid_parms checking failed"
    _check_sel_parms(_vis_xds, _vsel)
    _check_sel_parms(_img_xds, _isel, new_or_modified_data_variables=new_vars, append_to_in_id=True)
    g_in, g_out = _vsel["data_group_in"], _isel["data_group_out"]
    ds = _vis_mapping(_vis_xds, g_in)
    res = fn(ds, _gp, weight_key="IMAGING_WEIGHT" if "imaging_weight" in g_in else "WEIGHT")
    plane, sw = res[keys[0]], res[keys[1]]
    _img_xds[g_out["sum_weight"]] = _new_variable(sw[None], ["time", "chan", "pol"])
    _img_xds[g_out[keys[2]]] = _new_variable(plane[:, :, None], [keys[3], keys[4], "time", "chan", "pol"])
    _register_group(_img_xds, g_out)
    return _img_xds


def make_grid(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms):
    """GRID ['u','v','time','chan','pol'] complex and SUM_WEIGHT ['time','chan','pol'] in a copy of img_xds (make_grid.py:27-141)."""
    print("######################### Start make_grid #########################")
    return _imaging_call(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms,
                         {"sum_weight": "SUM_WEIGHT", "grid": "GRID"}, imaging.make_grid, ("GRID", "SUM_WEIGHT", "grid", "u", "v"))


def make_image(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms):
    """IMAGE ['l','m','time','chan','pol'] and SUM_WEIGHT ['time','chan','pol'] in a copy of img_xds (make_image.py:27-156)."""
    print("######################### Start make_image #########################")
    return _imaging_call(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms,
                         {"sum_weight": "SUM_WEIGHT", "image": "IMAGE"}, imaging.make_image, ("IMAGE", "SUM_WEIGHT", "image", "l", "m"))


def make_psf(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms):
    """PSF ['l','m','time','chan','pol'] and PSF_SUM_WEIGHT in a copy of img_xds (make_psf.py:27-160).  The reference also
    registers and computes `psf_fit` (fit_gaussian, :159): image analysis, out of scope -- the key is not added."""
    print("######################### Start make_psf #########################")
    return _imaging_call(vis_mxds, img_xds, grid_parms, vis_sel_parms, img_sel_parms,
                         {"sum_weight": "PSF_SUM_WEIGHT", "psf": "PSF"}, imaging.make_psf,
                         ("PSF", "PSF_SUM_WEIGHT", "psf", "l", "m"))
