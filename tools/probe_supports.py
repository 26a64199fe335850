"""Standard gridder on C2 (fp32, continuum, 4096^2) for supports 7 .. 15: window kernels vs the kernels they replace."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid as sg  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402
from probe_std_grid import timeit  # noqa: E402

d = synth.config_c2(dtype="f32")
vis, uvw, w, freq = (torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan"))
grid = torch.zeros((1, 2, 4096, 4096), dtype=torch.complex64, device="cuda")
gsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
out = {}
for S in (7, 9, 11, 13, 15):
    cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, S)).cuda()
    gp = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum", support=S)
    row = {}
    for name, algo in (("window", 4), ("track", 2), ("naive", 1)):
        if (name == "track" and S > 9) or (name == "naive" and S not in (11,)):
            continue
        row[name] = round(timeit(lambda: sg.standard_grid(vis, uvw, w, freq, cgk, gp, False, True, grid=grid, sum_weight=gsw,
                                                          algorithm=algo), n=3, warm=1)[0], 3)
    out["S=%d" % S] = row
print(json.dumps({"C2 fp32 continuum 4096^2, 115.6 M samples, ms per pass": out}))
