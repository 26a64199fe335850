"""Device timing of the degrid predict kernel (development tool)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_degrid
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
from tools.probe_std_grid import timeit
d = synth.config_c2(dtype="f32")
cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
gp = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum")
uvw, freq = torch.as_tensor(d["uvw"]).cuda(), torch.as_tensor(d["freq_chan"]).cuda()
for dt in (torch.complex64, torch.complex128):
    model = torch.randn((1, 2, 4096, 4096), dtype=dt, device="cuda")
    for algo in (2, 1):   # register-window kernel, gather kernel
        med, best = timeit(lambda: _standard_degrid._standard_degrid_numpy_wrap(model, uvw, freq, cgk, gp, normalize=True,
                                                                                algorithm=algo))
        print(json.dumps(dict(dtype=str(dt), algo=algo, ms=round(med, 3), gvis_s=round(d["weight"].size / med / 1e6, 2))))
