"""Per-chunk standard-gridder operators with the reference's names, arguments and return values.

Mirrors /root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:
  _standard_grid_numpy_wrap                    :123
  _standard_grid_psf_numpy_wrap                :180
  _standard_imaging_weight_degrid_numpy_wrap   :443
and calculate_briggs_parms (ngcasa/imaging/make_imaging_weight.py:198-213).

Inputs may be numpy arrays (host; the call goes through the C ABI's host entry point, which streams
them to the GPU and copies the result back) or torch CUDA tensors (device resident; nothing leaves the
GPU and torch tensors are returned).  Either way the arithmetic is done by libcngi_b200.so -- there is
no CPU fallback.  dtype selects the precision: float32/complex64 -> CNGI_F32, otherwise CNGI_F64; cell
indices and masks are fp64-exact in both.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, CHAN_CUBE, CHAN_CONTINUUM, ALGO_AUTO
from ._devutil import to_device

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _chan_mode(grid_parms):
    mode = grid_parms["chan_mode"]
    if mode == "cube":
        return CHAN_CUBE
    if mode == "continuum":
        return CHAN_CONTINUUM
    raise ValueError("chan_mode must be 'cube' or 'continuum', got %r" % (mode,))


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_c(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)


def _precision_of(weight):
    dt = weight.dtype
    if _is_torch(weight):
        return F32 if dt == torch.float32 else F64
    return F32 if dt == np.float32 else F64


def _dtypes(precision, on_torch):
    if on_torch:
        return (torch.float32, torch.complex64) if precision == F32 else (torch.float64, torch.complex128)
    return (np.float32, np.complex64) if precision == F32 else (np.float64, np.complex128)


def _host_out(shape, dtype):
    """Output buffer for the host path; pinned when possible so the D2H copy runs at full PCIe rate."""
    if torch is not None and torch.cuda.is_available():
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64,
               np.complex128: torch.complex128}[dtype]
        return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    return np.empty(shape, dtype=dtype)


def _fill_common(a, shape4, n_ic, n_uv, grid_parms):
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = shape4
    a.n_imag_chan, a.n_imag_pol = n_ic, shape4[3]
    a.n_u, a.n_v = int(n_uv[0]), int(n_uv[1])
    cell = grid_parms["cell_size"]
    a.delta_lm[0], a.delta_lm[1] = float(cell[0]), float(cell[1])
    a.chan_mode = _chan_mode(grid_parms)


def standard_grid(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, do_psf, complex_grid, flag=None,
                  algorithm=ALGO_AUTO, chan_group=0, time_segment=0, grid=None, sum_weight=None, time_chunk=0,
                  imaging_weight_from=None):
    """Shared body of the two reference wrappers (adds optional fused flags and accumulate-into buffers).

    Device path (torch CUDA tensors): `grid`/`sum_weight`, if given, are accumulated into (the device
    resident accumulator the graph level uses); otherwise fresh zeroed tensors are returned.

    imaging_weight_from (device path, image mode, support 7): `weight` holds the NATURAL weights and the imaging weights
    are formed inside the gridder (cngi_b200_standard_grid_weighted: _standard_imaging_weight_degrid_jit,
    _standard_grid.py:466-518, folded into phase 1) from a dict with
        density (n_imag_chan, n_pol, n_u', n_v') float64, kernel-side, any strides (e.g. pol planes expanded with stride 0),
        briggs_factors (2, n_imag_chan, n_pol), grid_parms (the density grid's: image_size_padded, cell_size),
        out (optional tensor like `weight` that receives the imaging weights).
    """
    L = _lib.lib()
    on_torch = _is_torch(weight)
    precision = _precision_of(weight)
    rdt, cdt = _dtypes(precision, on_torch)
    shape4 = tuple(int(s) for s in weight.shape)
    n_chan, n_pol = shape4[2], shape4[3]
    n_ic = n_chan if grid_parms["chan_mode"] == "cube" else 1
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    gdt = cdt if complex_grid else rdt

    a = _lib.StdGridArgs()
    _fill_common(a, shape4, n_ic, n_uv, grid_parms)
    a.support, a.oversampling = int(grid_parms["support"]), int(grid_parms["oversampling"])
    a.precision, a.do_psf, a.complex_grid = precision, int(bool(do_psf)), int(bool(complex_grid))
    a.algorithm, a.chan_group, a.time_segment = int(algorithm), int(chan_group), int(time_segment)

    if on_torch:
        _lib.require_device()
        dev = weight.device
        keep = []  # keep converted tensors alive until the launch is queued

        def dev_t(x, dt):
            t = to_device(x, dt, dev)
            keep.append(t)
            return t

        w = dev_t(weight, rdt)
        v = None if do_psf else dev_t(vis_data, cdt)
        f = None if (flag is None or do_psf) else dev_t(flag, torch.uint8)
        uvw_t = dev_t(uvw, torch.float64)
        freq_t = dev_t(freq_chan, torch.float64)
        cgk_t = dev_t(cgk_1D, torch.float64)
        if grid is None:
            grid = torch.zeros((n_ic, n_pol, int(n_uv[0]), int(n_uv[1])), dtype=gdt, device=dev)
        if sum_weight is None:
            sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev)
        assert grid.is_contiguous() and grid.dtype == gdt and sum_weight.dtype == torch.float64
        a.vis, a.weight, a.flag, a.uvw = _ptr(v), _ptr(w), _ptr(f), _ptr(uvw_t)
        a.freq_chan, a.cgk_1D, a.grid, a.sum_weight = _ptr(freq_t), _ptr(cgk_t), _ptr(grid), _ptr(sum_weight)
        if imaging_weight_from is not None:
            assert not do_psf and complex_grid, "imaging_weight_from: image mode only"
            src = imaging_weight_from
            rho = src["density"]
            n_uv_iw = np.asarray(src["grid_parms"]["image_size_padded"]).astype(np.int64)
            assert rho.dtype == torch.float64 and tuple(rho.shape) == (n_ic, n_pol, int(n_uv_iw[0]), int(n_uv_iw[1])), \
                tuple(rho.shape)
            bf = src["briggs_factors"]
            assert tuple(bf.shape) == (2, n_ic, n_pol), tuple(bf.shape)
            f = _lib.IwFusedArgs()
            # pol planes that are stride-0 views of one plane (how the pipeline / make_imaging_weight hand them in after
            # gridding pol plane 0 only) are identical by construction: the kernel gathers and divides once per sample
            f.pol_shared = int(n_pol >= 2 and _is_torch(bf) and bf.stride(2) == 0 and rho.stride(1) == 0) \
                if src.get("pol_shared") is None else int(bool(src["pol_shared"]))
            bf = dev_t(bf, torch.float64)
            f.density, f.briggs_factors = _ptr(rho), _ptr(bf)
            st = rho.stride()
            for i, v in enumerate((st[2], st[3], st[0], st[1])):
                f.density_stride[i] = int(v)
            out = src.get("out")
            if out is not None:
                assert out.is_contiguous() and out.dtype == rdt and tuple(out.shape) == shape4
            f.imaging_weight = _ptr(out)
            f.n_u, f.n_v = int(n_uv_iw[0]), int(n_uv_iw[1])
            cell_iw = src["grid_parms"]["cell_size"]
            f.delta_lm[0], f.delta_lm[1] = float(cell_iw[0]), float(cell_iw[1])
            with torch.cuda.device(dev):
                _lib.check(L.cngi_b200_standard_grid_weighted(C.byref(a), C.byref(f), _stream()),
                           "cngi_b200_standard_grid_weighted")
            return grid, sum_weight
        with torch.cuda.device(dev):
            _lib.check(L.cngi_b200_standard_grid(C.byref(a), _stream()), "cngi_b200_standard_grid")
        return grid, sum_weight

    # host path
    assert imaging_weight_from is None, "imaging_weight_from needs device-resident inputs"
    w = _np_c(weight, rdt)
    v = None if do_psf else _np_c(vis_data, cdt)
    f = None if (flag is None or do_psf) else _np_c(flag, np.uint8)
    uvw_h = _np_c(uvw, np.float64)
    freq_h = _np_c(freq_chan, np.float64)
    cgk_h = _np_c(cgk_1D, np.float64)
    grid = _host_out((n_ic, n_pol, int(n_uv[0]), int(n_uv[1])), gdt)
    sum_weight = _host_out((n_ic, n_pol), np.float64)
    a.vis, a.weight, a.flag, a.uvw = _ptr(v), _ptr(w), _ptr(f), _ptr(uvw_h)
    a.freq_chan, a.cgk_1D, a.grid, a.sum_weight = _ptr(freq_h), _ptr(cgk_h), _ptr(grid), _ptr(sum_weight)
    _lib.check(L.cngi_b200_standard_grid_host(C.byref(a), int(time_chunk)), "cngi_b200_standard_grid_host")
    return grid, sum_weight


def _standard_grid_numpy_wrap(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, **kw):
    """Grids visibilities * weights.  Returns (grid (n_imag_chan, n_pol, n_u, n_v), sum_weight (n_imag_chan, n_pol)).

    Reads the same grid_parms keys as the reference: chan_mode, image_size_padded, cell_size, oversampling,
    support, complex_grid, do_psf.
    """
    return standard_grid(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, grid_parms["do_psf"],
                         grid_parms["complex_grid"], **kw)


def _standard_grid_psf_numpy_wrap(uvw, weight, freq_chan, cgk_1D, grid_parms, **kw):
    """Grids weights only onto a REAL grid (PSF), or -- when grid_parms['do_imaging_weight'] -- the
    imaging-weight density grid (support 1, conjugate cell, pol-averaged weight; make_imaging_weight.py:153-161)."""
    if grid_parms.get("do_imaging_weight", False):
        from ._imaging_weight import imaging_weight_grid
        return imaging_weight_grid(uvw, weight, freq_chan, grid_parms, **kw)
    return standard_grid(None, uvw, weight, freq_chan, cgk_1D, grid_parms, grid_parms["do_psf"], False, **kw)


def standard_grid_image_psf(vis_data, uvw, weight, freq_chan, cgk_1D, grid_parms, flag=None, grid=None,
                            sum_weight=None, psf_grid=None, psf_sum_weight=None, force_fused=False):
    """Image grid AND psf grid of the same samples in ONE pass (cngi_b200_standard_grid_image_psf): what
    synthesis_imaging_cube.py:195-211 computes with _make_psf followed by _make_image on the same uvw and weights.
    Returns (grid complex, sum_weight, psf_grid real, psf_sum_weight), kernel-side layouts, accumulated into the
    buffers given.  Device path only (numpy inputs are uploaded); support must be 7 -- otherwise the two single
    passes are issued."""
    L = _lib.lib()
    _lib.require_device()
    like_torch = _is_torch(weight)
    dev = weight.device if like_torch else torch.device("cuda", torch.cuda.current_device())
    keep = []

    def dev_t(x, dt):
        t = to_device(x, dt, dev)
        keep.append(t)
        return t

    precision = _precision_of(weight)
    rdt, cdt = _dtypes(precision, True)
    w = dev_t(weight, rdt)
    shape4 = tuple(int(s) for s in w.shape)
    n_chan, n_pol = shape4[2], shape4[3]
    n_ic = n_chan if grid_parms["chan_mode"] == "cube" else 1
    n_uv = np.asarray(grid_parms["image_size_padded"]).astype(np.int64)
    gshape = (n_ic, n_pol, int(n_uv[0]), int(n_uv[1]))
    # measured on the B200 (tools/probe_fused.py, C2 at 4096^2): the fused pass saves 29 % (fp32 continuum), 16 % (fp32
    # cube), 12 % (fp64 continuum) and nothing for fp64 cubes (234 registers: two blocks per SM), which take two passes
    two_passes = int(grid_parms["support"]) != 7 or (precision == F64 and grid_parms["chan_mode"] == "cube" and not force_fused)
    if two_passes:
        gp = dict(grid_parms)
        psf_grid, psf_sum_weight = standard_grid(None, uvw, w, freq_chan, cgk_1D, gp, True, False, grid=psf_grid,
                                                 sum_weight=psf_sum_weight)
        grid, sum_weight = standard_grid(vis_data, uvw, w, freq_chan, cgk_1D, gp, False, True, flag=flag, grid=grid,
                                         sum_weight=sum_weight)
    else:
        a = _lib.StdGridArgs()
        _fill_common(a, shape4, n_ic, n_uv, grid_parms)
        a.support, a.oversampling = int(grid_parms["support"]), int(grid_parms["oversampling"])
        a.precision, a.do_psf, a.complex_grid = precision, 0, 1
        grid = torch.zeros(gshape, dtype=cdt, device=dev) if grid is None else grid
        psf_grid = torch.zeros(gshape, dtype=rdt, device=dev) if psf_grid is None else psf_grid
        sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev) if sum_weight is None else sum_weight
        psf_sum_weight = torch.zeros((n_ic, n_pol), dtype=torch.float64, device=dev) if psf_sum_weight is None \
            else psf_sum_weight
        assert grid.is_contiguous() and grid.dtype == cdt and psf_grid.is_contiguous() and psf_grid.dtype == rdt
        f = None if flag is None else dev_t(flag, torch.uint8)
        a.vis, a.weight, a.flag, a.uvw = _ptr(dev_t(vis_data, cdt)), _ptr(w), _ptr(f), _ptr(dev_t(uvw, torch.float64))
        a.freq_chan, a.cgk_1D = _ptr(dev_t(freq_chan, torch.float64)), _ptr(dev_t(cgk_1D, torch.float64))
        a.grid, a.sum_weight = _ptr(grid), _ptr(sum_weight)
        with torch.cuda.device(dev):
            _lib.check(L.cngi_b200_standard_grid_image_psf(C.byref(a), _ptr(psf_grid), _ptr(psf_sum_weight), _stream()),
                       "cngi_b200_standard_grid_image_psf")
    outs = (grid, sum_weight, psf_grid, psf_sum_weight)
    return outs if like_torch else tuple(t.cpu().numpy() for t in outs)
