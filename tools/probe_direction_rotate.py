"""Device-timing probe of cngi_b200_direction_rotate on the C2/C3 sample shape (development tool).

Calls the C ABI directly on preallocated buffers (the Python mirror adds ~1 ms of host work per call: small uploads and
the status read-back), CUDA events on the launching stream.  Prints one JSON object: ms and achieved GB/s
(algorithmic bytes = vis in + vis out + uvw in + uvw out) per precision."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import direction_rotate as dr, _lib  # noqa: E402
from cngi_prototype_b200._devutil import ptr, stream  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

n_t, n_b, n_c, n_p = 500, 903, 128, 2
g = torch.Generator(device="cuda").manual_seed(1)
uvw = torch.randn((n_t, n_b, 3), dtype=torch.float64, device="cuda", generator=g) * 300
uvw_rot = torch.empty_like(uvw)
ids = np.arange(7)
dirs = np.stack([1.0 + 4e-4 * np.cos(ids), 0.5 + 4e-4 * np.sin(ids)], 1)
field = torch.as_tensor(np.repeat((np.arange(n_t) % 7)[:, None], n_b, 1).astype(np.int64)).cuda()
freq = torch.as_tensor(np.linspace(345e9, 347e9, n_c)).cuda()
R, P, rid = dr.calc_rotation_mats(field, ids, dirs, dict(new_phase_center=[1.0, 0.5]))
Rt, Pt, ridt = torch.as_tensor(R).cuda(), torch.as_tensor(P).cuda(), torch.as_tensor(rid).cuda()
status = torch.zeros(1, dtype=torch.int32, device="cuda")
L = _lib.lib()
out = {}
for name, cdt, prec in (("f32", torch.complex64, _lib.F32), ("f64", torch.complex128, _lib.F64)):
    vis = torch.randn((n_t, n_b, n_c, n_p), dtype=cdt, device="cuda")
    vis_rot = torch.empty_like(vis)
    a = _lib.DirectionRotateArgs()
    a.n_time, a.n_baseline, a.n_chan, a.n_pol = n_t, n_b, n_c, n_p
    a.vis, a.vis_rot, a.uvw, a.uvw_rot, a.field, a.freq_chan = ptr(vis), ptr(vis_rot), ptr(uvw), ptr(uvw_rot), ptr(field), ptr(freq)
    a.uvw_rotmat, a.phase_rotation, a.rot_field_id, a.n_field, a.status = ptr(Rt), ptr(Pt), ptr(ridt), 7, ptr(status)
    a.common_tangent_reprojection, a.single_precision, a.precision = 1, 0, prec
    ms, best = timeit(lambda: _lib.check(L.cngi_b200_direction_rotate(C.byref(a), stream()), "direction_rotate"), n=9, warm=3)
    nbytes = 2 * vis.numel() * vis.element_size() + 2 * uvw.numel() * 8
    out[name] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "samples": vis.numel(), "Gvis/s": vis.numel() / ms / 1e6}
    del vis, vis_rot
assert int(status.item()) == 0
print(json.dumps(out))
