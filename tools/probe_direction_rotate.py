"""Device-timing probe of cngi_b200_direction_rotate on the C2/C3 sample shape (development tool).

Prints one JSON object: ms and achieved GB/s (algorithmic bytes = vis in + vis out + uvw in + uvw out) per precision."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import direction_rotate as dr  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

n_t, n_b, n_c, n_p = 500, 903, 128, 2
g = torch.Generator(device="cuda").manual_seed(1)
uvw = torch.randn((n_t, n_b, 3), dtype=torch.float64, device="cuda", generator=g) * 300
ids = np.arange(7)
dirs = np.stack([1.0 + 4e-4 * np.cos(ids), 0.5 + 4e-4 * np.sin(ids)], 1)
field = torch.as_tensor(np.repeat((np.arange(n_t) % 7)[:, None], n_b, 1).astype(np.int64)).cuda()
freq = np.linspace(345e9, 347e9, n_c)
R, P, rid = dr.calc_rotation_mats(field, ids, dirs, dict(new_phase_center=[1.0, 0.5]))
out = {}
for name, cdt in (("f32", torch.complex64), ("f64", torch.complex128)):
    vis = torch.randn((n_t, n_b, n_c, n_p), dtype=cdt, device="cuda")
    ms, _ = timeit(lambda: dr.rotate_chunk(vis, uvw, field, freq, R, P, rid, True, False))
    nbytes = 2 * vis.numel() * vis.element_size() + 2 * uvw.numel() * 8
    out[name] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "samples": vis.numel(), "Gvis/s": vis.numel() / ms / 1e6}
    del vis
print(json.dumps(out))
