"""Device timing of make_gridding_convolution_function at the config-3 size (n_pad 2048, os 10, max_support 15,
one dish type, 7 fields) and, with --cpu, the oracle (numpy.fft + scipy jn: the reference's own arithmetic) beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import make_gridding_convolution_function as mg  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

n_ant = 43
a1, a2 = np.triu_indices(n_ant, 1)
k = np.arange(7)
out = {}
for name, n_pad, types, dishes, blocks in (("c3_one_dish", 2048, np.zeros(n_ant, dtype=int), [10.7], [0.75]),
                                           ("c3_two_dishes", 2048, (np.arange(n_ant) % 4 == 0).astype(int), [10.7, 6.25], [0.75, 0.75])):
    cell = np.array([-0.02, 0.02]) * np.pi / (180 * 3600)
    gp = dict(function="casa_airy", list_dish_diameters=np.array(dishes), list_blockage_diameters=np.array(blocks),
              unique_ant_indx=types, basline_ant=np.stack([a1, a2], 1), freq_chan=np.linspace(345e9, 347e9, 128),
              pol=np.array([0, 1]), field_phase_dir=np.stack([1.0 + 3e-5 * np.cos(k), 0.5 + 3e-5 * np.sin(k)], 1),
              phase_center=np.array([1.0, 0.5]), oversampling=[10, 10], max_support=[15, 15])
    grid_parms = dict(image_size=np.array([n_pad, n_pad]), image_size_padded=np.array([n_pad, n_pad]), cell_size=cell)
    g = mg.make_gridding_convolution_function(gp, grid_parms)
    ms, best = timeit(lambda: mg.make_gridding_convolution_function(gp, grid_parms), n=5, warm=1)
    n_items = g["CONV_KERNEL"].shape[0] * g["CONV_KERNEL"].shape[1]
    out[name] = {"ms": ms, "items": int(n_items), "ms_per_item": ms / n_items,
                 "support": sorted(set(g["SUPPORT"].cpu().numpy().reshape(-1).tolist()))}
    if "--cpu" in sys.argv:
        from oracle import oracle as O
        t0 = time.time()
        o = O.make_gridding_convolution_function(gp, grid_parms)
        out[name]["cpu_oracle_s"] = time.time() - t0
        out[name]["rel_err_conv_kernel"] = float(np.abs(g["CONV_KERNEL"].cpu().numpy() - o["CONV_KERNEL"]).max() / np.abs(o["CONV_KERNEL"]).max())
        out[name]["support_equal"] = bool(np.array_equal(g["SUPPORT"].cpu().numpy(), o["SUPPORT"]))
print(json.dumps(out))
