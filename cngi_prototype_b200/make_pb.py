"""Primary-beam images on the GPU: make_pb and its two pattern functions.

Mirrors /root/reference/ngcasa/imaging/make_pb.py:60-118 (API function) and
_imaging_utils/_make_pb_symmetric.py:26-76 (_airy_disk), :79-132 (_casa_airy_disk) -- the functions the reference maps
over channel chunks (make_pb.py:107) and calls per chunk in synthesis_imaging_cube.py:262-284.  The pattern arithmetic
is the device function the convolution-function producer uses (csrc/gcf.cu); here it fills
pb[l, m, chan, pol, dish_type] = voltage^ipower directly.
"""
import copy
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import torch, device_of, ptr, stream
from .imaging import _check_grid_parms


def _pattern(function, freq_chan, pol, pb_parms, grid_parms, device=None):
    L = _lib.lib()
    dev = device if device is not None else device_of()
    freq = np.ascontiguousarray(np.asarray(freq_chan, dtype=np.float64).reshape(-1))
    dish = np.ascontiguousarray(pb_parms["list_dish_diameters"], dtype=np.float64)
    block = np.ascontiguousarray(pb_parms["list_blockage_diameters"], dtype=np.float64)
    size = np.asarray(grid_parms["image_size"]).astype(np.int64)
    centre = np.asarray(grid_parms.get("image_center", size // 2)).astype(np.int64)
    cell = np.asarray(grid_parms["cell_size"], dtype=np.float64)
    n_pol = len(pol)
    pb = torch.empty((int(size[0]), int(size[1]), len(freq), n_pol, len(dish)), dtype=torch.float64, device=dev)
    a = _lib.PbArgs()
    a.image_size[0], a.image_size[1] = int(size[0]), int(size[1])
    a.image_center[0], a.image_center[1] = int(centre[0]), int(centre[1])
    a.cell_size[0], a.cell_size[1] = float(cell[0]), float(cell[1])
    a.function, a.ipower = (1 if function == "casa_airy" else 0), int(pb_parms["ipower"])
    a.n_chan, a.freq_chan_host, a.n_pol = len(freq), freq.ctypes.data, n_pol
    a.n_dish, a.dish_diameter_host, a.blockage_diameter_host = len(dish), dish.ctypes.data, block.ctypes.data
    a.pb = ptr(pb)
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_make_pb(C.byref(a), stream()), "cngi_b200_make_pb")
        torch.cuda.current_stream().synchronize()     # freq / dish host arrays are read by the async copy
    return pb


def _airy_disk(freq_chan, pol, pb_parms, grid_parms, device=None):
    """(l, m, chan, pol, dish) CUDA tensor; same arguments as the reference's chunk function."""
    return _pattern("airy", freq_chan, pol, pb_parms, grid_parms, device)


def _casa_airy_disk(freq_chan, pol, pb_parms, grid_parms, device=None):
    return _pattern("casa_airy", freq_chan, pol, pb_parms, grid_parms, device)


def make_pb(img_dataset, pb_parms, grid_parms, sel_parms=None):
    """img_dataset: mapping with 'chan' (Hz) and 'pol'.  pb_parms: function 'casa_airy' (default) | 'airy',
    list_dish_diameters, list_blockage_diameters (metres).  Adds sel_parms['data_group_out']['pb'] (default 'PB'),
    (l, m, chan, pol, dish_type) -- the reference also inserts a length-1 time axis (make_pb.py:112) -- to a copy."""
    _pb = copy.deepcopy(pb_parms)
    _gp = copy.deepcopy(grid_parms)
    _pb.setdefault("function", "casa_airy")
    ok = isinstance(_pb.get("list_dish_diameters"), (list, tuple, np.ndarray)) and \
        isinstance(_pb.get("list_blockage_diameters"), (list, tuple, np.ndarray)) and \
        len(_pb["list_dish_diameters"]) == len(_pb["list_blockage_diameters"])
    assert ok, "######### ERROR: user_imaging_weights_parms checking failed"
    assert _check_grid_parms(_gp), "######### ERROR: grid_parms checking failed"
    assert _pb["function"] in ("airy", "casa_airy"), "Only the airy function has been implemented"
    _pb["ipower"] = 2
    name = ((sel_parms or {}).get("data_group_out") or {}).get("pb", "PB")
    out = dict(img_dataset)
    out[name] = _pattern(_pb["function"], img_dataset["chan"], img_dataset["pol"], _pb, _gp)
    out["dish_type"] = np.arange(len(_pb["list_dish_diameters"]))
    return out
