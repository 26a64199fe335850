"""One chunk of the config-5 cube through the gridder (8 channels x 2 pol, 9830^2 padded grid, fp32) -- the launch ncu
captures for the reduction sectors per sample of cube-mode gridding.  Prints samples per launch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid as sg  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402

n_time, n_chan = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 8
rng = np.random.default_rng(4321)
bl = synth.baseline_vectors(synth.antenna_layout(43, 300.0, rng))
ha = synth.EARTH_RATE * 6.0 * (np.arange(n_time) - n_time // 2)
uvw = torch.as_tensor(synth.uvw_tracks(bl, ha, np.deg2rad(-23.0))).cuda()
freq = torch.as_tensor(np.linspace(345.0e9, 347.0e9, 1024)[:n_chan]).cuda()
cell = 1.0 / (2.0 * np.max(np.linalg.norm(bl, axis=1)) * 347.0e9 / synth.C_LIGHT * 1.15)
gp = synth.grid_parms_for(9830, cell, chan_mode="cube")
shape = (n_time, len(bl), n_chan, 2)
vis = torch.view_as_complex(torch.randn(shape + (2,), dtype=torch.float32, device="cuda"))
w = torch.rand(shape, dtype=torch.float32, device="cuda") + 0.5
cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, 7)).cuda()
grid = torch.zeros((n_chan, 2, 9830, 9830), dtype=torch.complex64, device="cuda")
gsw = torch.zeros((n_chan, 2), dtype=torch.float64, device="cuda")
for _ in range(2):
    sg.standard_grid(vis, uvw, w, freq, cgk, gp, False, True, grid=grid, sum_weight=gsw)
torch.cuda.synchronize()
print("samples_per_launch", int(np.prod(shape)))
