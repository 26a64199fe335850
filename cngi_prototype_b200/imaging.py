"""API layer: make_imaging_weight / make_grid / make_psf / make_image on plain-mapping visibility datasets.

Mirrors the flat, stateless API functions of /root/reference/ngcasa/imaging (make_imaging_weight.py:20,
make_grid.py, make_psf.py:27, make_image.py:27) and their parameter handling
(_imaging_utils/_check_imaging_parms.py:22-41,123-146).  xarray and dask do not exist in this environment, so a
"vis dataset" here is any mapping with

    DATA (n_time, n_baseline, n_chan, n_pol) complex, UVW (n_time, n_baseline, 3), WEIGHT (like DATA, real),
    optional FLAG (bool/uint8, like DATA), optional IMAGING_WEIGHT, and chan (n_chan,) frequencies in Hz,

with numpy arrays or torch CUDA tensors as values (an xarray Dataset's ``.values`` drop in unchanged).  Where the
reference builds a lazy dask graph of per-chunk tasks plus a tree-sum (_standard_grid.py:58-98), these functions loop
over time chunks and accumulate into ONE device-resident grid -- the chunk operators are the same ones the dask graph
would call (cngi_prototype_b200._standard_grid etc.).  Results are returned as a dict of arrays of the input kind.
"""
import copy
import numbers

import numpy as np

from ._devutil import torch, is_torch, device_of, small_constant
from ._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D, correcting_function_1D
from ._standard_grid import standard_grid
from ._imaging_weight import (imaging_weight_grid, calculate_briggs_parms,
                              _standard_imaging_weight_degrid_numpy_wrap)
from ._fft import grid_to_image
from ._lazy import LazyDeviceArray, ChunkFeeder, streams, is_host

ARCSEC_TO_RAD = np.pi / (3600 * 180)


def _check_grid_parms(grid_parms):
    """Defaults / validation / unit conversion of grid_parms, in place (_check_imaging_parms.py:22-41).

    image_size [nx, ny] ints; cell_size [x, y] arcsec -> radians with x negated; fft_padding in [1, 10] default 1.2;
    chan_mode 'cube' (default) | 'continuum'; derives image_size_padded = int(fft_padding * image_size).
    """
    ok = True
    def bad(msg):
        nonlocal ok
        print("######### Parameter checking error: ", msg)
        ok = False
    if "image_size" not in grid_parms or len(grid_parms["image_size"]) != 2:
        bad("image_size must be a list of two ints")
    if "cell_size" not in grid_parms or len(grid_parms["cell_size"]) != 2 or \
            not all(isinstance(x, numbers.Number) for x in grid_parms["cell_size"]):
        bad("cell_size must be a list of two numbers (arcsec)")
    grid_parms.setdefault("fft_padding", 1.2)
    if not (isinstance(grid_parms["fft_padding"], numbers.Number) and 1 <= grid_parms["fft_padding"] <= 10):
        bad("fft_padding must be a number in [1, 10]")
    grid_parms.setdefault("chan_mode", "cube")
    if grid_parms["chan_mode"] not in ("cube", "continuum"):
        bad("chan_mode must be 'cube' or 'continuum'")
    if ok:
        grid_parms["image_size"] = np.array(grid_parms["image_size"]).astype(int)
        grid_parms.setdefault("image_center", grid_parms["image_size"] // 2)
        grid_parms["image_size_padded"] = (grid_parms["fft_padding"] * grid_parms["image_size"]).astype(int)
        grid_parms["image_center"] = np.array(grid_parms["image_center"])
        grid_parms["cell_size"] = ARCSEC_TO_RAD * np.array(grid_parms["cell_size"], dtype=np.float64)
        grid_parms["cell_size"][0] = -grid_parms["cell_size"][0]
    return ok


def _check_imaging_weights_parms(parms):
    """_check_imaging_parms.py:123-146: weighting natural (default) | uniform | briggs | briggs_abs, robust in [-2, 2]."""
    parms.setdefault("weighting", "natural")
    if parms["weighting"] not in ("natural", "uniform", "briggs", "briggs_abs"):
        print("######### Parameter checking error: weighting must be natural, uniform, briggs or briggs_abs")
        return False
    if parms["weighting"] in ("briggs", "briggs_abs"):
        parms.setdefault("robust", 0.5)
        if not (isinstance(parms["robust"], numbers.Number) and -2 <= parms["robust"] <= 2):
            print("######### Parameter checking error: robust must be a number in [-2, 2]")
            return False
    if parms["weighting"] == "briggs_abs":
        parms.setdefault("briggs_abs_noise", 1.0)
    return True


def _time_chunks(n_time, time_chunk):
    step = int(time_chunk) if time_chunk else n_time
    return [slice(t, min(n_time, t + step)) for t in range(0, n_time, max(step, 1))]


def _dev(ds, key, device, dtype=None):
    x = ds[key]
    if isinstance(x, LazyDeviceArray):
        x = x.device_tensor()
    if not is_torch(x) and dtype is not None and np.asarray(x).nbytes <= 256 * 1024:
        return small_constant(x, dtype, device)     # e.g. the channel frequencies: no synchronous copy per call
    t = x if is_torch(x) else torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device=device)


def _out(t, like_torch):
    return t if like_torch else t.cpu().numpy()


class _HostCall:
    """Context of an API call on HOST arrays: kernels go to the calling thread's compute stream (so that calls made
    from several threads overlap), results come back through page-locked memory.  For device inputs it is a no-op and
    everything stays on the caller's current stream."""

    def __init__(self, dev, host):
        self.dev, self.host = dev, host
        self._ctx = None

    def __enter__(self):
        if self.host:
            self._dctx = torch.cuda.device(self.dev)
            self._dctx.__enter__()
            self._ctx = torch.cuda.stream(streams(self.dev).compute)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
            self._dctx.__exit__(*exc)
        return False

    def lazy(self, base, dims=None, sources=None, deferred=None):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        return LazyDeviceArray(base, ev, dims, sources, deferred)

    def result(self, base, dims=None):
        """Device tensor -> what the caller gets: the tensor (device inputs) or a numpy array (host inputs; one flat
        D2H into page-locked memory, permuted as a view like the reference's moveaxis)."""
        if not self.host:
            return base if dims is None else base.permute(*dims)
        return self.lazy(base, dims).numpy()


def make_imaging_weight(vis_dataset, imaging_weights_parms, grid_parms, time_chunk=0, _density_hook=None):
    """Adds IMAGING_WEIGHT to a copy of the dataset (natural: aliases WEIGHT; uniform / briggs: density grid on the
    UNPADDED image size, Briggs factors, degrid -- make_imaging_weight.py:95-104,144-247).

    Device (torch CUDA) variables: IMAGING_WEIGHT is a device tensor.  Host (numpy) variables: UVW and WEIGHT are fed to
    the GPU in time chunks (copy of chunk k+1 under the density kernel of chunk k) and IMAGING_WEIGHT comes back as a
    LazyDeviceArray -- the counterpart of the lazy dask variable the reference returns (:95-104): it reads like a numpy
    array, and a following make_grid / make_psf / make_image uses the device copy without moving it.
    _density_hook(density, sum_weight): called on the device accumulators before the Briggs factors are taken -- where
    distributed.make_imaging_weight sums the density over the ranks that hold the other time shards."""
    _iw = copy.deepcopy(imaging_weights_parms)
    _gp = copy.deepcopy(grid_parms)
    assert _check_imaging_weights_parms(_iw), "######### ERROR: imaging_weights_parms checking failed"
    out = dict(vis_dataset)
    if _iw["weighting"] == "natural":
        out["IMAGING_WEIGHT"] = vis_dataset["WEIGHT"]
        return out
    assert _check_grid_parms(_gp), "######### ERROR: grid_parms checking failed"
    _gp["image_size_padded"] = _gp["image_size"]          # no padding: no FFT follows (:153)
    _gp.update(oversampling=0, support=1, do_psf=True, complex_grid=False, do_imaging_weight=True)
    host = is_host(vis_dataset["WEIGHT"])
    dev = device_of(vis_dataset["WEIGHT"], vis_dataset["UVW"])
    with _HostCall(dev, host) as call:
        freq = _dev(vis_dataset, "chan", dev, torch.float64)
        feeder = ChunkFeeder({"UVW": vis_dataset["UVW"], "WEIGHT": vis_dataset["WEIGHT"]},
                             {"UVW": torch.float64, "WEIGHT": None}, dev, time_chunk)
        n_pol = int(vis_dataset["WEIGHT"].shape[3])
        density = sw = None
        for _, blk in feeder:                                # the reference's chunk loop + tree sum, on one device grid
            density, sw = imaging_weight_grid(blk["UVW"], blk["WEIGHT"], freq, _gp, grid=density, sum_weight=sw,
                                              first_pol_only=n_pol >= 2)
        uvw, w = feeder.full("UVW"), feeder.full("WEIGHT")
        if _density_hook is not None:
            _density_hook(density[:, :1] if n_pol >= 2 else density, sw[:, :1] if n_pol >= 2 else sw)
        if n_pol >= 2:   # only pol plane 0 was gridded (all planes are identical, _standard_grid.py:328-330): the other
            rho = density[:, :1].expand(-1, n_pol, -1, -1)   # planes are stride-0 views of it, no replication pass
            bf = calculate_briggs_parms(density[:, :1], sw[:, :1], _iw).expand(-1, -1, n_pol)
        else:
            rho, bf = density, calculate_briggs_parms(density, sw, _iw)
        def degrid():
            return _standard_imaging_weight_degrid_numpy_wrap(rho, uvw, w, bf, freq, _gp, kernel_side_layout=True)

        if host:
            # Deferred: the weight degrid runs only if somebody reads IMAGING_WEIGHT (or make_psf needs it); make_grid /
            # make_image form the weights inside the gridder from what is stored here (cngi_b200_standard_grid_weighted)
            out["IMAGING_WEIGHT"] = call.lazy(None, sources=feeder.sources({"UVW": vis_dataset["UVW"],
                                                                            "WEIGHT": vis_dataset["WEIGHT"]}),
                                              deferred=dict(make=degrid, shape=tuple(w.shape), dtype=w.dtype, device=dev,
                                                            natural=w, density=rho, briggs_factors=bf.contiguous(),
                                                            grid_parms=_gp))
        else:
            out["IMAGING_WEIGHT"] = degrid()
    return out


def _flagged_weights(w, flag):
    """apply_flags semantics for the weights (cngi/vis/apply_flags.py:53 NaNs every variable with FLAG's dims)."""
    return torch.where(flag != 0, torch.full((), float("nan"), dtype=w.dtype, device=w.device), w)


def _grid(vis_dataset, grid_parms, do_psf, time_chunk, weight_key, apply_flags=False):
    """Returns (grid, sum_weight, checked grid_parms) on the device, on the current stream."""
    _gp = copy.deepcopy(grid_parms)
    assert _check_grid_parms(_gp), "######### ERROR: grid_parms checking failed"
    _gp["oversampling"], _gp["support"] = 100, 7          # make_image.py:106-107
    _gp["complex_grid"], _gp["do_psf"], _gp["do_imaging_weight"] = (not do_psf), do_psf, False
    cgk_1D = _create_prolate_spheroidal_kernel_1D(_gp["oversampling"], _gp["support"])
    wkey = weight_key if weight_key in vis_dataset else "WEIGHT"
    use_flag = bool(apply_flags) and "FLAG" in vis_dataset
    wsrc = vis_dataset[wkey]
    dev = wsrc.device if isinstance(wsrc, LazyDeviceArray) else device_of(wsrc, vis_dataset["UVW"])
    if hasattr(vis_dataset, "iter_device_chunks"):
        # zarr-backed dataset (read_vis.VisDataset): time blocks are decoded into pinned buffers, copied on a copy stream
        # and gridded as they arrive -- the reference's per-chunk dask tasks as a three-stage pipeline on one GPU
        names = [wkey, "UVW"] + ([] if do_psf else ["DATA"]) + (["FLAG"] if use_flag else [])
        freq = _dev(vis_dataset, "chan", dev, torch.float64)
        grid = sw = None
        for _, blk in vis_dataset.iter_device_chunks(names, time_chunk, dev):
            w = _flagged_weights(blk[wkey], blk["FLAG"]) if (use_flag and do_psf) else blk[wkey]
            grid, sw = standard_grid(None if do_psf else blk["DATA"], blk["UVW"], w, freq, cgk_1D, _gp, do_psf,
                                     not do_psf, flag=blk.get("FLAG") if not do_psf else None, grid=grid, sum_weight=sw)
        return grid, sw, _gp
    freq = _dev(vis_dataset, "chan", dev, torch.float64)
    arrays, dtypes, known = {"UVW": vis_dataset["UVW"], wkey: wsrc}, {"UVW": torch.float64, wkey: None}, {}
    fused = None
    if isinstance(wsrc, LazyDeviceArray):   # weights made by make_imaging_weight on this dataset: already on the device,
        if not do_psf and not wsrc.computed and wsrc.deferred is not None:   # together with the UVW they were computed from
            # not materialised yet: grid with the NATURAL weights and let the gridder form the imaging weights (A4 in A1)
            fused = {k: wsrc.deferred[k] for k in ("density", "briggs_factors", "grid_parms")}
            torch.cuda.current_stream(dev).wait_event(wsrc.ready)
            known[wkey] = wsrc.deferred["natural"]
        else:
            known[wkey] = wsrc.device_tensor()
        known["UVW"] = wsrc.source("UVW", vis_dataset["UVW"])
    if not do_psf:
        arrays["DATA"], dtypes["DATA"] = vis_dataset["DATA"], None
    if use_flag:
        arrays["FLAG"], dtypes["FLAG"] = vis_dataset["FLAG"], torch.uint8
    feeder = ChunkFeeder(arrays, dtypes, dev, time_chunk, known=known)
    grid = sw = None
    for _, blk in feeder:
        w = _flagged_weights(blk[wkey], blk["FLAG"]) if (use_flag and do_psf) else blk[wkey]
        grid, sw = standard_grid(None if do_psf else blk["DATA"], blk["UVW"], w, freq, cgk_1D, _gp, do_psf, not do_psf,
                                 flag=blk["FLAG"] if (use_flag and not do_psf) else None, grid=grid, sum_weight=sw,
                                 imaging_weight_from=fused)
    return grid, sw, _gp


def _host_call_for(vis_dataset, weight_key, data_key):
    wkey = weight_key if weight_key in vis_dataset else "WEIGHT"
    probe = vis_dataset[data_key] if data_key in vis_dataset else vis_dataset["UVW"]
    wsrc = vis_dataset[wkey]
    dev = wsrc.device if isinstance(wsrc, LazyDeviceArray) else device_of(wsrc, vis_dataset["UVW"], probe)
    host = is_host(probe) and not hasattr(vis_dataset, "iter_device_chunks")
    return _HostCall(dev, host)


def make_grid(vis_dataset, grid_parms, time_chunk=0, weight_key="IMAGING_WEIGHT", apply_flags=False, lazy=False,
              _grid_hook=None):
    """GRID (u, v, chan, pol) complex and SUM_WEIGHT (chan, pol): make_grid.py:112-137 (stops after gridding).

    Host (numpy) DATA is fed to the GPU in time chunks, each chunk's copy running under the gridding kernel of the chunk
    before; the result returns through page-locked memory as numpy arrays (GRID as the permuted view of the kernel-side
    array, like the reference's moveaxis).  apply_flags: see make_image.
    lazy=True (host datasets): GRID and SUM_WEIGHT come back as LazyDeviceArrays -- the call returns as soon as the work
    is queued, and the transfer to the host happens (on its own stream) when the caller first reads the values, the way
    the reference's lazy result is moved by `.compute()`; a caller that walks many datasets can so read the result of
    one while the next is being fed.  _grid_hook(grid, sum_weight): see distributed.make_grid."""
    like_torch = is_torch(vis_dataset["DATA"])
    with _host_call_for(vis_dataset, weight_key, "DATA") as call:
        grid, sw, _ = _grid(vis_dataset, grid_parms, False, time_chunk, weight_key, apply_flags)
        if _grid_hook is not None:
            _grid_hook(grid, sw)
        if call.host and lazy:
            return {"GRID": call.lazy(grid, (2, 3, 0, 1)), "SUM_WEIGHT": call.lazy(sw)}
        if call.host:
            return {"GRID": call.result(grid, (2, 3, 0, 1)), "SUM_WEIGHT": call.result(sw)}
    return {"GRID": _out(grid.permute(2, 3, 0, 1), like_torch), "SUM_WEIGHT": _out(sw, like_torch)}


def _image(vis_dataset, grid_parms, do_psf, time_chunk, weight_key, chan_chunk=0, apply_flags=False):
    """grid -> ifft -> crop -> / sum_weight / PS image.  chan_chunk > 0 (cube mode only): the image channels are
    processed `chan_chunk` at a time, so that the padded uv-grids of one chunk, not of the whole cube, are resident
    (the per-channel independence synthesis_imaging_cube.py:105-124 exploits with dask chunks)."""
    n_chan = int(vis_dataset["chan"].shape[0])
    if chan_chunk and grid_parms.get("chan_mode", "cube") == "cube" and chan_chunk < n_chan:
        images, sws = [], []
        for c0 in range(0, n_chan, int(chan_chunk)):
            sl = slice(c0, min(n_chan, c0 + int(chan_chunk)))
            sub = {k: (v[sl] if k == "chan" else
                       ((v.device_tensor() if isinstance(v, LazyDeviceArray) else v)[:, :, sl] if getattr(v, "ndim", 0) == 4 else v))
                   for k, v in vis_dataset.items()}
            img, sw = _image(sub, grid_parms, do_psf, time_chunk, weight_key, apply_flags=apply_flags)
            images.append(img)
            sws.append(sw)
        return torch.cat(images, dim=2), torch.cat(sws, dim=0)
    grid, sw, gp = _grid(vis_dataset, grid_parms, do_psf, time_chunk, weight_key, apply_flags)
    cu, cv = correcting_function_1D(gp["image_size_padded"], gp["image_size"])
    return grid_to_image(grid, gp["image_size"], sum_weight=sw, corr_u=cu, corr_v=cv), sw


def make_psf(vis_dataset, grid_parms, time_chunk=0, weight_key="IMAGING_WEIGHT", chan_chunk=0, apply_flags=False):
    """PSF (l, m, chan, pol) and PSF_SUM_WEIGHT (chan, pol): real PS gridding of the weights, inverse FFT, crop,
    / sum_weight / PS correcting image (make_psf.py:105-130).  The Gaussian beam fit (fit_gaussian) is out of scope.
    apply_flags: see make_image."""
    like_torch = is_torch(vis_dataset["UVW"])
    with _host_call_for(vis_dataset, weight_key, "UVW") as call:
        img, sw = _image(vis_dataset, grid_parms, True, time_chunk, weight_key, chan_chunk, apply_flags)
        if call.host:
            return {"PSF": call.result(img.contiguous()), "PSF_SUM_WEIGHT": call.result(sw)}
    return {"PSF": _out(img, like_torch), "PSF_SUM_WEIGHT": _out(sw, like_torch)}


def make_image(vis_dataset, grid_parms, time_chunk=0, weight_key="IMAGING_WEIGHT", chan_chunk=0, apply_flags=False):
    """IMAGE (l, m, chan, pol) and SUM_WEIGHT (chan, pol): complex PS gridding of DATA * weight, inverse FFT, crop,
    / sum_weight / PS correcting image (make_image.py:106-130).

    FLAG: like the reference, make_image / make_psf / make_grid do NOT read FLAG -- the user runs cngi.vis.apply_flags
    first (cngi/vis/apply_flags.py:53), which NaNs DATA and the weights together, and NaN samples are skipped
    (_standard_grid.py:340).  apply_flags=True fuses that step: FLAG != 0 drops the sample from the image AND from the
    psf (the weight is treated as NaN there), so IMAGE / SUM_WEIGHT and PSF / PSF_SUM_WEIGHT use the same sample set,
    exactly as if apply_flags had been run."""
    like_torch = is_torch(vis_dataset["DATA"])
    with _host_call_for(vis_dataset, weight_key, "DATA") as call:
        img, sw = _image(vis_dataset, grid_parms, False, time_chunk, weight_key, chan_chunk, apply_flags)
        if call.host:
            return {"IMAGE": call.result(img.contiguous()), "SUM_WEIGHT": call.result(sw)}
    return {"IMAGE": _out(img, like_torch), "SUM_WEIGHT": _out(sw, like_torch)}


def predict_modelvis_image(img_dataset, vis_dataset, grid_parms, model_key="MODEL", time_chunk=0):
    """MODEL_DATA (n_time, n_baseline, n_chan, n_pol) from a model image (l, m, chan, pol), Jy/pixel -- BASELINE config 4.

    The reference's predict_modelvis_image.py:20-40 is a stub whose comments list the steps (fourier_transform,
    _degrid); this is the inverse of make_image: divide by the PS correcting image, zero pad to image_size_padded,
    forward FFT, degrid with the same prolate-spheroidal taps (support 7, oversampling 100) normalised by the tap sum.
    A model with one channel is a continuum model (chan_mode 'continuum'), otherwise one image channel per data channel.
    Returns a copy of vis_dataset with MODEL_DATA added (the variable self_cal reads, calibration/self_cal.py:107-109)."""
    from ._standard_degrid import _standard_degrid_numpy_wrap
    from ._fft import image_to_grid
    _gp = copy.deepcopy(grid_parms)
    model = img_dataset[model_key]
    n_model_chan = int(model.shape[2])
    _gp.setdefault("chan_mode", "continuum" if n_model_chan == 1 else "cube")
    assert _check_grid_parms(_gp), "######### ERROR: grid_parms checking failed"
    n_chan = int(vis_dataset["chan"].shape[0])
    assert n_model_chan == (1 if _gp["chan_mode"] == "continuum" else n_chan), \
        "######### ERROR: model image channels do not match chan_mode"
    _gp["oversampling"], _gp["support"] = 100, 7
    cgk_1D = _create_prolate_spheroidal_kernel_1D(_gp["oversampling"], _gp["support"])
    like_torch = is_torch(vis_dataset["UVW"])
    dev = device_of(model, vis_dataset["UVW"])
    uvw, freq = _dev(vis_dataset, "UVW", dev, torch.float64), _dev(vis_dataset, "chan", dev, torch.float64)
    cu, cv = correcting_function_1D(_gp["image_size_padded"], _gp["image_size"])
    model_t = model if is_torch(model) else torch.as_tensor(np.ascontiguousarray(model))
    grid = image_to_grid(model_t.to(dev), _gp["image_size_padded"], corr_u=cu, corr_v=cv)
    parts = [_standard_degrid_numpy_wrap(grid, uvw[sl], freq, cgk_1D, _gp, normalize=True)
             for sl in _time_chunks(uvw.shape[0], time_chunk)]
    out = dict(vis_dataset)
    out["MODEL_DATA"] = _out(parts[0] if len(parts) == 1 else torch.cat(parts, dim=0), like_torch)
    return out
