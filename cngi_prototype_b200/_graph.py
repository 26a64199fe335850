"""Graph-level boundary (SURVEY.md section 8b "secondary boundary"): the reference's three graph builders with their
own names and arguments, on mapping datasets.

  _graph_standard_grid(vis_dataset, cgk_1D, grid_parms, sel_parms)          _standard_grid.py:23-106
  _graph_standard_degrid(vis_dataset, grid, briggs_factors, cgk_1D, grid_parms, sel_parms)      :381-440
  _graph_aperture_grid(vis_dataset, gcf_dataset, grid_parms, sel_parms)     _aperture_grid.py:25-142

The reference walks the dask chunks of the imaging-weight variable (time x baseline x chan, pol forced to one chunk,
:35-36), wraps one per-chunk operator call per chunk in dask.delayed and sums the partial grids with a pairwise tree
(_tree_sum_list :109-120).  Here the same chunk walk calls the same per-chunk operators, but every chunk ACCUMULATES into
one device-resident grid (the reductions are atomic, so no tree and no whole-grid copies between workers).  A "dataset" is
a mapping of arrays (numpy or torch CUDA); variable names come from sel_parms['data_group_in'] exactly as in the
reference ({'data', 'uvw', 'imaging_weight'}), `chan` holds the frequencies, and the optional entry
vis_dataset['chunks'] = {'time': n, 'baseline': n, 'chan': n} plays the role of the dask chunk sizes.
Returns follow the reference: [grid (n_u, n_v, n_chan, n_pol), sum_weight (n_chan, n_pol)] after its moveaxis.
"""
import numpy as np

from ._devutil import torch, is_torch, device_of
from ._standard_grid import standard_grid
from ._imaging_weight import imaging_weight_grid, _standard_imaging_weight_degrid_numpy_wrap
from ._aperture_grid import _aperture


def _t(x, dev, dtype=None):
    t = x if is_torch(x) else torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=dev, dtype=dtype) if dtype is not None else t.to(device=dev)


def _chunk_slices(n, size):
    size = int(size) if size else n
    return [slice(i, min(n, i + size)) for i in range(0, n, max(size, 1))] or [slice(0, 0)]


def _walk(vis_dataset, shape):
    ch = vis_dataset.get("chunks", {}) if hasattr(vis_dataset, "get") else {}
    for st in _chunk_slices(shape[0], ch.get("time", 0)):
        for sb in _chunk_slices(shape[1], ch.get("baseline", 0)):
            for sc in _chunk_slices(shape[2], ch.get("chan", 0)):
                yield st, sb, sc


def _out(t, like_torch):
    return t if like_torch else t.cpu().numpy()


def _graph_standard_grid(vis_dataset, cgk_1D, grid_parms, sel_parms):
    """grid_parms keys as the per-chunk wrappers read them (chan_mode, image_size_padded, cell_size, oversampling,
    support, complex_grid, do_psf, do_imaging_weight).  do_psf selects the real weight-only gridder (or, with
    do_imaging_weight, the density grid) like the reference's branch at :63-75."""
    names = sel_parms["data_group_in"]
    w_in = vis_dataset[names["imaging_weight"]]
    like_torch = is_torch(w_in)
    dev = device_of(w_in, vis_dataset[names["uvw"]])
    w, uvw = _t(w_in, dev), _t(vis_dataset[names["uvw"]], dev, torch.float64)
    freq = _t(vis_dataset["chan"], dev, torch.float64)
    do_psf = bool(grid_parms["do_psf"])
    vis = None if do_psf else _t(vis_dataset[names["data"]], dev)
    cube = grid_parms["chan_mode"] == "cube"
    n_u, n_v = (int(x) for x in grid_parms["image_size_padded"])
    n_chan, n_pol = int(w.shape[2]), int(w.shape[3])
    rdt = torch.float32 if w.dtype == torch.float32 else torch.float64
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    density = do_psf and bool(grid_parms.get("do_imaging_weight", False))
    gdt = torch.float64 if density else (rdt if do_psf else cdt)
    n_ic = n_chan if cube else 1
    grid = torch.zeros((n_ic, n_pol, n_u, n_v), dtype=gdt, device=dev)
    sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev)
    for st, sb, sc in _walk(vis_dataset, w.shape):
        g, s = (grid[sc], sum_weight[sc]) if cube else (grid, sum_weight)     # a chan chunk owns its image planes (:83-86)
        wc = w[st, sb, sc].contiguous()
        if wc.numel() == 0:
            continue
        uc = uvw[st, sb].contiguous()
        if density:
            imaging_weight_grid(uc, wc, freq[sc], grid_parms, grid=g, sum_weight=s)
        else:
            standard_grid(None if do_psf else vis[st, sb, sc].contiguous(), uc, wc, freq[sc], cgk_1D, grid_parms, do_psf,
                          not do_psf, grid=g, sum_weight=s)
    return [_out(grid.permute(2, 3, 0, 1), like_torch), _out(sum_weight, like_torch)]


def _graph_standard_degrid(vis_dataset, grid, briggs_factors, cgk_1D, grid_parms, sel_parms):
    """Imaging-weight degrid graph (the only branch the reference implements, :416-428): `grid` is the density
    (n_u, n_v, n_chan, n_pol) as _graph_standard_grid returns it, briggs_factors (2, n_chan, n_pol).
    Returns the imaging weights (n_time, n_baseline, n_chan, n_pol)."""
    assert grid_parms.get("do_imaging_weight", False), "Degridding of visibilities and psf still needs to be implemented"
    names = sel_parms["data_group_in"]
    w_in = vis_dataset[names["imaging_weight"]]
    like_torch = is_torch(w_in)
    dev = device_of(w_in, vis_dataset[names["uvw"]])
    w, uvw = _t(w_in, dev), _t(vis_dataset[names["uvw"]], dev, torch.float64)
    freq = _t(vis_dataset["chan"], dev, torch.float64)
    rho = _t(grid, dev, torch.float64)
    bf = _t(briggs_factors, dev, torch.float64)
    cube = grid_parms["chan_mode"] == "cube"
    out = torch.zeros(tuple(w.shape), dtype=w.dtype, device=dev)
    for st, sb, sc in _walk(vis_dataset, w.shape):
        wc = w[st, sb, sc].contiguous()
        if wc.numel() == 0:
            continue
        a_sc = sc if cube else slice(0, 1)
        out[st, sb, sc] = _standard_imaging_weight_degrid_numpy_wrap(rho[:, :, a_sc].contiguous(), uvw[st, sb].contiguous(),
                                                                     wc, bf[:, a_sc].contiguous(), freq[sc], grid_parms)
    return _out(out, like_torch)


def _graph_aperture_grid(vis_dataset, gcf_dataset, grid_parms, sel_parms):
    """grid_parms['grid_weights'] -> A6 with WEIGHT_CONV_KERNEL; else A5 (psf when grid_parms['do_psf']) with CONV_KERNEL
    (_aperture_grid.py:67-112).  vis_dataset['FIELD_ID'] (n_time, n_baseline); gcf_dataset as
    make_gridding_convolution_function returns it.  The grid is always complex (:62)."""
    names = sel_parms["data_group_in"]
    w_in = vis_dataset[names["imaging_weight"]]
    like_torch = is_torch(w_in)
    dev = device_of(w_in, vis_dataset[names["uvw"]])
    w, uvw = _t(w_in, dev), _t(vis_dataset[names["uvw"]], dev, torch.float64)
    freq = _t(vis_dataset["chan"], dev, torch.float64)
    field = _t(vis_dataset["FIELD_ID"], dev, torch.int64).reshape(w.shape[0], w.shape[1])
    weights_mode = bool(grid_parms.get("grid_weights", False))
    do_psf = bool(grid_parms["do_psf"]) and not weights_mode
    vis = None if (weights_mode or do_psf) else _t(vis_dataset[names["data"]], dev)

    def host(key):
        x = gcf_dataset[key]
        return x.cpu().numpy() if is_torch(x) else np.asarray(x)

    gp = dict(grid_parms)
    gp["field_id"] = host("field_id").astype(np.int64)              # :59
    gp.setdefault("oversampling", host("oversampling"))
    gp["complex_grid"] = True
    cf_bl, cf_ch, cf_pol, support = host("CF_BASELINE_MAP"), host("CF_CHAN_MAP"), host("CF_POL_MAP"), host("SUPPORT")
    kernel = gcf_dataset["WEIGHT_CONV_KERNEL" if weights_mode else "CONV_KERNEL"]
    entry = "cngi_b200_aperture_weight_grid" if weights_mode else "cngi_b200_aperture_grid"
    cube = gp["chan_mode"] == "cube"
    n_u, n_v = (int(x) for x in gp["image_size_padded"])
    n_chan, n_pol = int(w.shape[2]), int(w.shape[3])
    cdt = torch.complex64 if w.dtype == torch.float32 else torch.complex128
    n_ic = n_chan if cube else 1
    grid = torch.zeros((n_ic, n_pol, n_u, n_v), dtype=cdt, device=dev)
    sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev)
    for st, sb, sc in _walk(vis_dataset, w.shape):
        wc = w[st, sb, sc].contiguous()
        if wc.numel() == 0:
            continue
        g, s = (grid[sc], sum_weight[sc]) if cube else (grid, sum_weight)
        _aperture(entry, None if vis is None else vis[st, sb, sc].contiguous(), uvw[st, sb].contiguous(), wc,
                  field[st, sb].contiguous(), cf_bl[sb], cf_ch[sc], cf_pol, kernel, support, gcf_dataset["PHASE_GRADIENT"],
                  freq[sc], gp, do_psf, grid=g, sum_weight=s)
    return [_out(grid.permute(2, 3, 0, 1), like_torch), _out(sum_weight, like_torch)]
